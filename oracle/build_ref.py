"""Build the part of the UNMODIFIED reference that compiles here into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  The reference's pure-Cython modules
(``topo_param.pyx``, ``transform.pyx``, ``direction.pyx``) are compiled from the
sources WHERE THEY LIE under /root/reference/horayzon with the reference's own
flags (``-O3 -ffast-math``, libs m/pthread: reference ``setup.py:24,39,50-64``).
Nothing is copied into the repository; generated C files and the extension
modules go to ``oracle/_ref/horayzon/`` (git-ignored, but shipped to the GPU
box).  ``horizon_comp.cpp`` / ``shadow_comp.cpp`` need Intel Embree 4 + oneTBB,
which are not installed and not installable offline: they are NOT built.

No ``__init__.py`` is written: the reference's own ``__init__`` imports the
Embree-linked modules, so ``horayzon`` is used as a namespace package:

    sys.path.insert(0, "oracle/_ref"); from horayzon import topo_param
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/horayzon"
OUT = os.path.join(HERE, "_ref", "horayzon")
MODULES = ("topo_param", "transform", "direction")
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"


def build(force=False):
    import numpy as np
    if not os.path.isdir(REF_SRC):
        raise FileNotFoundError(REF_SRC + " is not available (GPU box?): use the prebuilt oracle/_ref")
    os.makedirs(OUT, exist_ok=True)
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    for m in MODULES:
        pyx = os.path.join(REF_SRC, m + ".pyx")
        target = os.path.join(OUT, m + ext)
        if not force and os.path.exists(target) and os.path.getmtime(target) >= os.path.getmtime(pyx):
            continue
        c_file = os.path.join(OUT, m + ".c")
        subprocess.check_call([sys.executable, "-m", "cython", "-3", pyx, "-o", c_file,
                               "--module-name", "horayzon." + m], stderr=subprocess.DEVNULL)
        subprocess.check_call([GCC, "-O3", "-ffast-math", "-fPIC", "-shared", "-w",
                               "-I", np.get_include(), "-I", sysconfig.get_paths()["include"],
                               c_file, "-o", target, "-lm", "-lpthread"])
        os.remove(c_file)
    return OUT


def load():
    """Import the compiled reference modules (or raise ImportError)."""
    root = os.path.join(HERE, "_ref")
    if root not in sys.path:
        sys.path.insert(0, root)
    from horayzon import topo_param, transform, direction  # noqa: F401
    return topo_param, transform, direction


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
