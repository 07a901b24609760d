"""CPU tests of the boundary: the C-ABI library loads and exports every symbol
include/horayzon_b200.h declares, the wrappers keep the reference's signatures
and validation messages, and compute calls fail loudly without a GPU."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest

import horayzon_b200 as hb
from horayzon_b200 import resident, synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "horayzon_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hzb_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    names = _declared_functions()
    assert len(names) >= 25
    L = ctypes.CDLL(os.path.join(ROOT, "horayzon_b200", "libhorayzon_b200.so"))
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    L.hzb_version.restype = ctypes.c_char_p
    assert b"sm_100a" in L.hzb_version()


def test_library_is_built_for_sm_100a_only(built):
    out = os.popen("cuobjdump --list-elf %s 2>/dev/null" % os.path.join(ROOT, "horayzon_b200", "libhorayzon_b200.so")).read()
    if not out:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_signatures_mirror_the_reference():
    """Parameter names/defaults of horizon.pyx:29-49, :218-232 and shadow.pyx:27-38."""
    doc = hb.horizon.horizon_gridded.__doc__
    assert "horizon_gridded" in doc or "Horizon" in doc
    c = syn.make_config("cfg1", n=48)
    # keyword names must be accepted exactly as the reference spells them
    with pytest.raises(ValueError, match="invalid input argument for ray_algorithm"):
        hb.horizon.horizon_gridded(vert_grid=c["vert_grid"], dem_dim_0=48, dem_dim_1=48, vec_norm=c["vec_norm"],
                                   vec_north=c["vec_north"], offset_0=16, offset_1=16, dist_search=5.0, azim_num=8,
                                   hori_acc=0.25, ray_algorithm="nope", geom_type="grid",
                                   vert_simp=np.zeros(4, np.float32), num_vert_simp=1,
                                   tri_ind_simp=np.zeros(4, np.int32), num_tri_simp=1, elev_ang_low_lim=-15.0,
                                   mask=None, hori_fill=0.0, ray_org_elev=0.01)
    t = hb.shadow.Terrain()
    assert all(hasattr(t, m) for m in ("initialise", "shadow", "sw_dir_cor"))


def _hg_args(c):
    return [c["vert_grid"], c["dem_dim_0"], c["dem_dim_1"], c["vec_norm"], c["vec_north"],
            c["offset_0"], c["offset_1"], c["dist_search"]]


def test_horizon_gridded_validation_messages():
    """Same exceptions and wording as horizon.pyx:109-156."""
    c = syn.make_config("cfg1", n=48)
    a = _hg_args(c)
    with pytest.raises(ValueError, match="inconsistency between input arguments vert_grid"):
        hb.horizon.horizon_gridded(a[0][:100], *a[1:])
    with pytest.raises(ValueError, match="dem_dim_1, offset_0, offset_1 and vec_norm"):
        hb.horizon.horizon_gridded(a[0], a[1], a[2], a[3], a[4], 40, 16, 5.0)
    with pytest.raises(ValueError, match="vec_norm and/or vec_north"):
        hb.horizon.horizon_gridded(a[0], a[1], a[2], a[3], a[4][:, :-1], 16, 16, 5.0)
    with pytest.raises(ValueError, match="invalid input argument for geom_type"):
        hb.horizon.horizon_gridded(*a, geom_type="mesh")
    with pytest.raises(ValueError, match="limit of hori_acc"):
        hb.horizon.horizon_gridded(*a, hori_acc=10.5)
    with pytest.raises(ValueError, match="shape of mask is inconsistent"):
        hb.horizon.horizon_gridded(*a, mask=np.ones((3, 3), np.uint8))
    with pytest.raises(TypeError, match="minimal allowed value for 'ray_org_elev'"):
        hb.horizon.horizon_gridded(*a, ray_org_elev=0.001)
    with pytest.raises(ValueError, match="triangle indices of simplified outer domain"):
        hb.horizon.horizon_gridded(*a, vert_simp=np.zeros(9, np.float32), num_vert_simp=3,
                                   tri_ind_simp=np.array([0, 1, 3], np.int32), num_tri_simp=1)
    with pytest.raises(ValueError, match="32'767"):
        hb.horizon.horizon_gridded(np.zeros(40000 * 3, np.float32), 40000, 1, a[3][:1, :1], a[4][:1, :1], 0, 0, 5.0)
    with pytest.raises((ValueError, TypeError)):  # float64 buffer rejected like the typed Cython signature
        hb.horizon.horizon_gridded(a[0].astype(np.float64), *a[1:])


def test_horizon_locations_validation_messages():
    """Same exceptions and wording as horizon.pyx:282-313."""
    c = syn.make_config("cfg1", n=48)
    co = np.zeros((3, 3), np.float32); v = np.zeros((3, 3), np.float32); v[:, 2] = 1
    with pytest.raises(ValueError, match="length\\(s\\) of 'coords' incorrect"):
        hb.horizon.horizon_locations(c["vert_grid"], 48, 48, co[:2], v, v, 5.0)
    with pytest.raises(ValueError, match="length of array 'ray_org_elev'"):
        hb.horizon.horizon_locations(c["vert_grid"], 48, 48, co, v, v, 5.0, ray_org_elev=np.ones(2, np.float32))
    with pytest.raises(TypeError, match="not implemented for horizon distance"):
        hb.horizon.horizon_locations(c["vert_grid"], 48, 48, co, v, v, 5.0, ray_algorithm="guess_constant", hori_dist_out=True)
    with pytest.raises(TypeError, match="minimal allowed value"):
        hb.horizon.horizon_locations(c["vert_grid"], 48, 48, co, v, v, 5.0, ray_org_elev=np.array([0.001], np.float32))


def test_terrain_validation_messages():
    """Same exceptions and wording as shadow.pyx:87-133, 165-168, 195-198."""
    n, rim = 40, 8
    x, y, z = syn.sinusoid_dem(n, n, 50.0, 100.0, 1000.0, 0, 2)
    vg = syn.rearrange_pad_buffer(x, y, z)
    tilt = syn.tilt_vectors(x, y, z, rim)
    ny = nx = n - 2 * rim
    norm, _ = syn.planar_frames(ny, nx)
    one = np.ones((ny, nx), np.float32); mask = np.ones((ny, nx), np.uint8)
    t = hb.shadow.Terrain()
    with pytest.raises(ValueError, match="'vert_grid', 'dem_dim_0' and 'dem_dim_1'"):
        t.initialise(vg[:10], n, n, rim, rim, tilt, norm, one, one, mask)
    with pytest.raises(ValueError, match="not normalised"):
        t.initialise(vg, n, n, rim, rim, tilt * 1.1, norm, one, one, mask)
    with pytest.raises(ValueError, match="not all input arrays are C-contiguous"):
        t.initialise(vg, n, n, rim, rim, tilt, norm, np.asfortranarray(one), one, mask)
    with pytest.raises(TypeError, match="'ang_max' must be in the range"):
        t.initialise(vg, n, n, rim, rim, tilt, norm, one, one, mask, ang_max=80.0)
    with pytest.raises(ValueError, match="'surf_enl_fac',  'elevation' and/or 'mask'"):
        t.initialise(vg, n, n, rim, rim, tilt, norm, one[:-1], one, mask)
    with pytest.raises(ValueError, match="'sun_position' has incorrect shape"):
        t.shadow(np.zeros(2, np.float32), np.zeros((ny, nx), np.uint8))
    with pytest.raises(RuntimeError, match="initialise"):
        t.shadow(np.zeros(3, np.float32), np.zeros((ny, nx), np.uint8))


def test_integral_validation_messages():
    """topo_param.pyx:400-405."""
    az = np.zeros(4, np.float32); h = np.zeros((2, 2, 4), np.float32); t = np.zeros((2, 2, 3), np.float32)
    with pytest.raises(ValueError, match="Inconsistent/incorrect shapes"):
        hb.topo_param.sky_view_factor(az[:3], h, t)
    with pytest.raises(ValueError, match="incorrect data type"):
        hb.topo_param.visible_sky_fraction(az, h.astype(np.float64), t)
    with pytest.raises(ValueError, match="Inconsistent/incorrect shapes"):
        hb.topo_param.topographic_openness(az[:2], h)


def test_compute_fails_loudly_without_gpu():
    """No CPU fallback: on a box without a CUDA device every compute entry point raises."""
    if resident.lib().hzb_device_count() > 0:
        pytest.skip("a GPU is present")
    c = syn.make_config("cfg1", n=48)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        hb.horizon.horizon_gridded(*_hg_args(c))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        hb.topo_param.sky_view_factor(np.zeros(4, np.float32), np.zeros((2, 2, 4), np.float32), np.ones((2, 2, 3), np.float32))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        resident.Scene(c["vert_grid"], 48, 48)
    # the additive entry points of round 2: fused SVF, multi-GPU, quantised output, replicated terrains
    tilt = np.zeros((c["ny"], c["nx"], 3), np.float32); tilt[..., 2] = 1.0
    with pytest.raises(RuntimeError, match="no CUDA device"):
        hb.horizon.horizon_gridded(*_hg_args(c), svf_vec_tilt=tilt)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        hb.horizon.horizon_gridded(*_hg_args(c), devices=0)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        hb.horizon.horizon_gridded_quantised(*_hg_args(c))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        hb.multi.MultiTerrain(devices=0)
    lon = np.array([[8.0, 8.1]]); lat = np.array([[46.0, 46.0]])
    with pytest.raises(RuntimeError, match="no CUDA device"):
        hb.transform.lonlat2ecef(lon, lat, np.zeros((1, 2), np.float32), ellps="WGS84")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        hb.direction.surf_norm(lon, lat)


def test_transform_direction_validation_messages():
    """Argument checks of transform.pyx:43-51, 129-139, 215-224 and direction.pyx:35-40, 104-115."""
    lon = np.array([[8.0, 8.1]]); lat = np.array([[46.0, 46.0]]); h = np.zeros((1, 2), np.float32)
    with pytest.raises(ValueError, match="Inconsistent shapes"):
        hb.transform.lonlat2ecef(lon, lat[:, :1], h, ellps="WGS84")
    with pytest.raises(ValueError, match="incorrect data type"):
        hb.transform.lonlat2ecef(lon, lat, h.astype(np.float64), ellps="WGS84")
    with pytest.raises(ValueError, match="Unknown value for 'ellps'"):
        hb.transform.lonlat2ecef(lon, lat, h, ellps="clarke")
    with pytest.raises(ValueError, match="must be instance of class 'TransformerEcef2enu'"):
        hb.transform.ecef2enu(lon, lat, lon, None)
    with pytest.raises(ValueError, match="Incorrect shape"):
        hb.transform.ecef2enu_vector(np.zeros(3, np.float32), hb.transform.TransformerEcef2enu(8.0, 46.0, "sphere"))
    with pytest.raises(ValueError, match="'lon_or' is outside of valid range"):
        hb.transform.TransformerEcef2enu(181.0, 46.0, "sphere")
    with pytest.raises(ValueError, match="Unknown value for 'ellps'"):
        hb.transform.TransformerEcef2enu(8.0, 46.0, "clarke")
    with pytest.raises(ValueError, match="incorrect data type"):
        hb.direction.surf_norm(lon.astype(np.float32), lat)
    with pytest.raises(ValueError, match="Inconsistent shapes"):
        hb.direction.north_dir(lon, lat, lon, np.zeros((1, 3, 3), np.float32), ellps="WGS84")
    t = hb.transform.TransformerEcef2enu(8.0, 46.0, "WGS84")   # attributes as in transform.pyx:455-483
    assert abs(np.sqrt(t.x_ecef_or ** 2 + t.y_ecef_or ** 2 + t.z_ecef_or ** 2) - 6.3670e6) < 2e4


def test_product_never_touches_the_oracle():
    """The product tree must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "horayzon_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".pyx", ".cu", ".cuh", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in text and "hzb_oracle" not in text and "orc_" not in text, f
    deps = os.popen("ldd %s" % os.path.join(pkg, "libhorayzon_b200.so")).read()
    assert "oracle" not in deps


def test_vertex_buffer_wire_format():
    """auxiliary.py:49-133: [y][x][xyz] float32 + >= 16 zeros, 16-byte multiple."""
    x = np.arange(6, dtype=np.float32).reshape(2, 3); y = x + 10; z = x + 20
    b = syn.rearrange_pad_buffer(x, y, z)
    assert b.dtype == np.float32 and b.nbytes % 16 == 0 and len(b) >= 18 + 16
    assert np.array_equal(b[:18].reshape(2, 3, 3)[..., 0], x) and np.array_equal(b[:18].reshape(2, 3, 3)[..., 2], z)
    assert np.all(b[18:] == 0)
    with pytest.raises(TypeError):
        syn.rearrange_pad_buffer(x.astype(np.float64), y, z)


def test_signature_parameter_order():
    # Cython functions expose no inspect.signature by default; the docstring's first line does (embedsignature off)
    # -> check through a positional call instead: all 8 leading positionals as in the reference
    c = syn.make_config("cfg1", n=48)
    with pytest.raises(ValueError, match="limit of hori_acc"):
        hb.horizon.horizon_gridded(c["vert_grid"], 48, 48, c["vec_norm"], c["vec_north"], 16, 16, 5.0, 8, 11.0)


def test_product_tables_equal_the_oracle_tables():
    """The elevation / azimuth tables are the alphabet of the horizon output (horizon_comp.cpp:711-731): the
    product's host-side builder (hzb_horizon_tables, no device needed) must give the oracle's bits, for the
    defaults and for odd parameters."""
    import ctypes
    import oracle
    L = resident.lib()
    f32p = ctypes.POINTER(ctypes.c_float)
    L.hzb_horizon_tables.argtypes = [ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                     f32p, f32p, f32p, f32p, f32p]
    for azim_num, dist, acc, low in ((360, 50.0, 0.25, -15.0), (37, 8.0, 0.1, -40.0), (1, 1.0, 4.0, -2.0), (720, 12.0, 0.15, -25.0)):
        want = oracle.tables(azim_num, dist, acc, low)
        n = L.hzb_horizon_tables(azim_num, dist, acc, low, 0, None, None, None, None, None)
        assert n == len(want["elev_ang"])
        bufs = [np.empty(n, np.float32) for _ in range(3)] + [np.empty(azim_num, np.float32) for _ in range(2)]
        assert L.hzb_horizon_tables(azim_num, dist, acc, low, n, *[b.ctypes.data_as(f32p) for b in bufs]) == n
        for got, key in zip(bufs, ("elev_ang", "elev_sin", "elev_cos", "azim_sin", "azim_cos")):
            assert np.array_equal(got, want[key]), (key, azim_num, acc)



def test_round2_additive_api_validation():
    """Argument checks of the additive keywords (no device needed: they fire before any native call)."""
    c = syn.make_config("cfg1", n=48)
    a = _hg_args(c)
    tilt = np.zeros((c["ny"], c["nx"], 3), np.float32); tilt[..., 2] = 1.0
    with pytest.raises(ValueError, match="incorrect shape"):
        hb.horizon.horizon_gridded(*a, svf_vec_tilt=tilt[1:])
    with pytest.raises(ValueError, match="azim_first"):
        hb.horizon.horizon_gridded(*a, svf_vec_tilt=tilt, azim_first=True)
    with pytest.raises(ValueError, match="at least two azimuth"):
        hb.horizon.horizon_gridded(*a, svf_vec_tilt=tilt, azim_num=1)
    with pytest.raises(ValueError, match="must not be negative"):
        hb.horizon.horizon_gridded(*a, vert_simp=np.zeros(9, np.float32), num_vert_simp=3,
                                   tri_ind_simp=np.array([0, 1, -2], np.int32), num_tri_simp=1)
    with pytest.raises(ValueError, match="16-bit output"):
        hb.horizon.horizon_gridded_quantised(*a, hori_acc=0.001)          # 525 000 table entries
    idx = np.array([[[0xFFFF, 3, 5], [0xFFFF, 0xFFFF, 0xFFFF]]], np.uint16)
    got = hb.horizon.dequantise(idx, np.array([[0.25, -1.0]], np.float32), np.arange(10, dtype=np.float32))
    assert got.dtype == np.float32 and np.array_equal(got, np.array([[[0.25, 3, 5], [-1, -1, -1]]], np.float32))
    # the debug switches are explicit calls, not environment variables; unknown names are refused
    assert resident.lib().hzb_debug_option(b"reset", 0) == 0
    assert resident.lib().hzb_debug_option(b"no_such_option", 1) != 0
    assert resident.lib().hzb_shard_rows(1199, 0, 8) == 152 and resident.lib().hzb_shard_rows(1199, 7, 8) == 148
    assert sum(resident.lib().hzb_shard_rows(1199, r, 8) for r in range(8)) == 1200


def _plan(H, W, lo, hi, off0, off1, ny, nx, r0, r1, rank, world, K, acc=0.25, low=-15.0, alg="guess_constant", ctas=148 * 6, dist=50.0):
    L = resident.lib()
    out = (ctypes.c_longlong * 7)()
    lo_a, hi_a = (ctypes.c_float * 3)(*lo), (ctypes.c_float * 3)(*hi)
    rc = L.hzb_plan_queue(H, W, lo_a, hi_a, off0, off1, ny, nx, r0, r1, rank, world, K, ctypes.c_float(dist), ctypes.c_float(acc),
                          ctypes.c_float(low), alg.encode(), ctas, out)
    assert rc == 0
    return dict(zip(("seg", "by0", "by1", "bx", "tail", "total", "tiles"), [int(v) for v in out]))


def test_queue_plan_of_the_bench_workloads(built):
    """The host-side queue layout (hzb_plan_queue: pure host code): which launches get a split tail, and where the band
    along the DEM's edge ends (cells closer to the edge than relief / tan(-low limit) may see below the elevation table
    and are never split)."""
    # cfg2: 1201 x 1201 at 90 m, relief ~3.6 km -> band ~150 cells; one GPU and one of eight interleaved shards
    lo, hi = (0.0, 0.0, -1800.0), (108000.0, 108000.0, 1800.0)
    p1 = _plan(1201, 1201, lo, hi, 1, 1, 1199, 1199, 0, 1199, 0, 1, 360)
    band = int(np.ceil(3600.0 / np.tan(np.deg2rad(15.0)) / 90.0))
    assert p1["seg"] == 4 and p1["tail"] == 2 * 148 * 6 * 4
    assert p1["by0"] == (band - 1 + 3) // 4 and p1["bx"] == (band - 1 + 7) // 8
    assert p1["tiles"] == 150 * 300 and p1["total"] == p1["tiles"] + 3 * p1["tail"]
    p8 = [_plan(1201, 1201, lo, hi, 1, 1, 1199, 1199, 0, 1199, r, 8, 360) for r in range(8)]
    assert all(p["seg"] == 4 for p in p8)
    assert sum(p["tiles"] for p in p8) == p1["tiles"]
    assert sum(p["by1"] - p["by0"] for p in p8) == p1["by1"] - p1["by0"]      # the shards' interiors partition the interior
    assert all(p["tail"] == (p["by1"] - p["by0"]) * (150 - 2 * p["bx"]) for p in p8)      # at N = 8 the whole interior is split
    # the north-star workload on one GPU: 40+ cells per lane, the tail is noise -> whole chains only
    lo4, hi4 = (0.0, 0.0, -790.0), (11998.0, 11998.0, 790.0)
    assert _plan(6000, 6000, lo4, hi4, 1, 1, 5998, 5998, 0, 5998, 0, 1, 360, dist=12.0)["seg"] == 1
    # inner domain far from the DEM's edge (the reference's usual set-up): no band at all
    far = _plan(1201, 1201, lo, hi, 300, 300, 601, 601, 0, 601, 0, 1, 360)
    assert far["seg"] == 4 and far["by0"] == 0 and far["bx"] == 0 and far["by1"] == (601 + 3) // 4
    # few azimuths, tiny elevation table, low limit above the horizontal (guess_constant): never split
    assert _plan(1201, 1201, lo, hi, 1, 1, 1199, 1199, 0, 1199, 0, 1, 36)["seg"] == 1
    assert _plan(1201, 1201, lo, hi, 1, 1, 1199, 1199, 0, 1199, 0, 1, 360, acc=30.0)["seg"] == 1
    assert _plan(1201, 1201, lo, hi, 1, 1, 1199, 1199, 0, 1199, 0, 1, 360, low=5.0)["seg"] == 1
    # search distance shorter than relief / tan(-low limit): a slope may fall below the table all the way -> never split
    assert _plan(1201, 1201, lo, hi, 1, 1, 1199, 1199, 0, 1199, 0, 1, 360, dist=10.0)["seg"] == 1
    assert _plan(1201, 1201, lo, hi, 1, 1, 1199, 1199, 0, 1199, 0, 1, 360, dist=14.0)["seg"] == 4
    # independent azimuths need no band
    ind = _plan(1201, 1201, lo, hi, 1, 1, 1199, 1199, 0, 1199, 0, 1, 360, alg="binary_search")
    assert ind["seg"] == 4 and ind["by0"] == 0 and ind["bx"] == 0


def test_ctypes_stats_struct_matches_the_header(built, tmp_path):
    """resident.Stats (ctypes) must have the size and field offsets of hzb_stats in include/horayzon_b200.h: the library
    writes the struct through the pointer the binding hands it."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("gcc unavailable")
    fields = [n for n, _ in resident.Stats._fields_]
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "horayzon_b200.h"\nint main(void){printf("%zu", sizeof(hzb_stats));'
                   + "".join('printf(" %%zu", offsetof(hzb_stats, %s));' % f for f in fields) + "return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call([gcc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    vals = [int(v) for v in subprocess.check_output([str(exe)], text=True).split()]
    assert vals[0] == ctypes.sizeof(resident.Stats)
    assert vals[1:] == [getattr(resident.Stats, f).offset for f in fields]
