"""In-tree build of everything the package needs.

    python horayzon_b200/_build.py           # or __graft_entry__.build()

1. ``csrc/*.cu`` -> ``libhorayzon_b200.so`` with nvcc for sm_100a
   (``-gencode arch=compute_100a,code=sm_100a -lineinfo``).
2. ``horizon.pyx``, ``shadow.pyx``, ``topo_param.pyx``, ``transform.pyx``, ``direction.pyx`` -> C (Cython) -> extension
   modules next to this file, linked against the library with rpath $ORIGIN.
The build products stay in-tree (git-ignored) so that they travel to the GPU box.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PYX = ("horizon", "shadow", "topo_param", "transform", "direction")
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_cuda(verbose=False):
    env = dict(os.environ)
    out = subprocess.run(["make", "-C", os.path.join(HERE, "csrc"), "-j8"], env=env,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if out.returncode != 0 or verbose:
        sys.stdout.write(out.stdout)
    if out.returncode != 0:
        raise RuntimeError("nvcc build of libhorayzon_b200.so failed")
    return os.path.join(HERE, "libhorayzon_b200.so")


def build_cython(verbose=False):
    import numpy as np
    from Cython.Build import cythonize  # noqa: F401  (presence check)
    ext_suffix = sysconfig.get_config_var("EXT_SUFFIX")
    py_inc = sysconfig.get_paths()["include"]
    lib = os.path.join(HERE, "libhorayzon_b200.so")
    header = os.path.join(ROOT, "include", "horayzon_b200.h")
    for name in PYX:
        pyx = os.path.join(HERE, name + ".pyx")
        c_file = os.path.join(HERE, "_gen_" + name + ".c")
        target = os.path.join(HERE, name + ext_suffix)
        if not _stale(target, [pyx, header, lib]):
            continue
        subprocess.check_call([sys.executable, "-m", "cython", "-3", pyx, "-o", c_file,
                               "--module-name", "horayzon_b200." + name])
        cmd = [GCC, "-O2", "-fPIC", "-shared", "-fno-strict-aliasing",
               "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
               "-I", os.path.join(ROOT, "include"), "-I", np.get_include(), "-I", py_inc,
               c_file, "-o", target, "-L", HERE, "-lhorayzon_b200", "-Wl,-rpath,$ORIGIN"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        os.remove(c_file)


def build_all(verbose=False):
    build_cuda(verbose)
    build_cython(verbose)


if __name__ == "__main__":
    build_all(verbose=True)
