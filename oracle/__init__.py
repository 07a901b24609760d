"""CPU oracle -- TEST INFRASTRUCTURE ONLY (see oracle/hzb_oracle.cpp header).

ctypes front-end for ``oracle/libhzb_oracle.so`` with the call shapes of the
reference's Python API (``horizon.pyx:29-197, 218-370``; ``shadow.pyx:17-200``;
``topo_param.pyx:377-603``).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs may import this package; the product
(``horayzon_b200``) never does.

Parity status: ray path "parity unpinned" (Embree absent, reference has no
golden vectors); SVF/VSF/openness pinned by ``tests/golden/`` vectors generated
from the reference's own compiled ``topo_param.pyx``.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhzb_oracle.so")

_f32p = ctypes.POINTER(ctypes.c_float)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_i32p = ctypes.POINTER(ctypes.c_int32)


def build(force=False):
    """Compile the oracle (gcc) if the shared object is missing or stale."""
    srcs = [os.path.join(_HERE, "hzb_oracle.cpp"),
            # the product's host/device sources (search state machine, triangle tests) are compiled in as units under test
            os.path.join(os.path.dirname(_HERE), "horayzon_b200", "csrc", "hzb_search.cuh"),
            os.path.join(os.path.dirname(_HERE), "horayzon_b200", "csrc", "hzb_tri.cuh"),
            os.path.join(os.path.dirname(_HERE), "horayzon_b200", "csrc", "hzb_box.cuh"),
            os.path.join(os.path.dirname(_HERE), "horayzon_b200", "csrc", "hzb_hd.cuh"),
            os.path.join(os.path.dirname(_HERE), "horayzon_b200", "csrc", "hzb_queue.cuh")]
    if (force or not os.path.exists(_SO)
            or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs if os.path.exists(f))):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libhzb_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_last_error.restype = ctypes.c_char_p
        _lib.orc_terrain_create.restype = ctypes.c_void_p
        _lib.orc_terrain_destroy.argtypes = [ctypes.c_void_p]
    return _lib


def _p(a, typ):
    return a.ctypes.data_as(typ)


def _check(rc):
    if rc != 0:
        raise RuntimeError(lib().orc_last_error().decode())


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    """OpenMP threads of the oracle (overrides OMP_NUM_THREADS, which torchrun exports as 1)."""
    lib().orc_set_num_threads(int(n))


def set_exact_predicate(on):
    """Evaluate every ray/triangle test in double on the exact float inputs (no epsilon) instead of the
    specified fp32 arithmetic: the reference point for how much the outputs depend on rounding."""
    lib().orc_set_exact_predicate(1 if on else 0)


def set_fast_traversal(on):
    """Scenes built while this is on also get an implicit 4-ary hierarchy over the grid quads (four boxes per SSE step) and
    their any-hit casts use it: same decisions (boxes only cull), several times faster.  For the CPU timing legs of
    bench.py; the parity checks keep the plain binary-BVH walker."""
    lib().orc_set_fast_traversal(1 if on else 0)


def last_timing():
    """(BVH build seconds, ray tracing seconds) of the last oracle call."""
    b, t = ctypes.c_double(), ctypes.c_double()
    lib().orc_last_timing(ctypes.byref(b), ctypes.byref(t))
    return b.value, t.value


def tables(azim_num, dist_search, hori_acc, elev_ang_low_lim):
    """Elevation / azimuth tables as horizon_comp.cpp:711-731 builds them."""
    L = lib()
    n = L.orc_tables(int(azim_num), ctypes.c_float(dist_search), ctypes.c_float(hori_acc),
                     ctypes.c_float(elev_ang_low_lim), 0, None, None, None, None, None)
    ea, es, ec = (np.empty(n, np.float32) for _ in range(3))
    asn, acs = (np.empty(azim_num, np.float32) for _ in range(2))
    L.orc_tables(int(azim_num), ctypes.c_float(dist_search), ctypes.c_float(hori_acc),
                 ctypes.c_float(elev_ang_low_lim), n, _p(ea, _f32p), _p(es, _f32p), _p(ec, _f32p),
                 _p(asn, _f32p), _p(acs, _f32p))
    return dict(elev_ang=ea, elev_sin=es, elev_cos=ec, azim_sin=asn, azim_cos=acs)


def _azim(azim_num):
    azim = np.empty(azim_num, dtype=np.float32)
    for i in range(azim_num):
        azim[i] = ((2 * np.pi) / azim_num * i)
    return azim


def horizon_gridded(vert_grid, dem_dim_0, dem_dim_1, vec_norm, vec_north, offset_0, offset_1,
                    dist_search, azim_num=360, hori_acc=0.25, ray_algorithm="guess_constant",
                    geom_type="grid", vert_simp=None, num_vert_simp=1, tri_ind_simp=None,
                    num_tri_simp=1, elev_ang_low_lim=-15.0, mask=None, hori_fill=0.0,
                    ray_org_elev=0.01, brute_force=False, return_rays=False):
    """Oracle twin of ``horizon.horizon_gridded`` (horizon.pyx:29-197)."""
    if vert_simp is None:
        vert_simp = np.zeros(4, np.float32)
    if tri_ind_simp is None:
        tri_ind_simp = np.zeros(4, np.int32)
    if mask is None:
        mask = np.ones(vec_norm.shape[:2], np.uint8)
    vert_grid = np.ascontiguousarray(vert_grid, np.float32)
    vec_norm = np.ascontiguousarray(vec_norm, np.float32)
    vec_north = np.ascontiguousarray(vec_north, np.float32)
    vert_simp = np.ascontiguousarray(vert_simp, np.float32)
    tri_ind_simp = np.ascontiguousarray(tri_ind_simp, np.int32)
    mask = np.ascontiguousarray(mask, np.uint8)
    ny, nx = vec_norm.shape[:2]
    hori = np.full((ny, nx, azim_num), np.nan, np.float32)
    rays = ctypes.c_ulonglong(0)
    _check(lib().orc_horizon_gridded(
        _p(vert_grid, _f32p), int(dem_dim_0), int(dem_dim_1), _p(vec_norm, _f32p),
        _p(vec_north, _f32p), int(offset_0), int(offset_1), _p(hori, _f32p), ny, nx,
        int(azim_num), ctypes.c_float(dist_search), ctypes.c_float(hori_acc),
        ray_algorithm.encode(), geom_type.encode(), _p(vert_simp, _f32p), int(num_vert_simp),
        _p(tri_ind_simp, _i32p), int(num_tri_simp), ctypes.c_float(elev_ang_low_lim),
        _p(mask, _u8p), ctypes.c_float(hori_fill), ctypes.c_float(ray_org_elev),
        int(bool(brute_force)), ctypes.byref(rays)))
    if return_rays:
        return hori, _azim(azim_num), rays.value
    return hori, _azim(azim_num)


class Scene:
    """DEM (+ optional TIN) with its BVH kept between calls: the oracle twin of the reference's scene
    (initializeScene, horizon_comp.cpp:101-231) for computing selected inner-domain ROWS without
    rebuilding the BVH per call (stratified CPU-baseline samples, sampled-row parity checks)."""

    def __init__(self, vert_grid, dem_dim_0, dem_dim_1, vert_simp=None, num_vert_simp=0, tri_ind_simp=None,
                 num_tri_simp=0):
        self.vert_grid = np.ascontiguousarray(vert_grid, np.float32)     # must outlive the native scene
        self.dims = (int(dem_dim_0), int(dem_dim_1))
        vs = ti = None
        if vert_simp is not None and num_vert_simp >= 3:
            vs = np.ascontiguousarray(vert_simp, np.float32); ti = np.ascontiguousarray(tri_ind_simp, np.int32)
        self._keep = (vs, ti)
        L = lib()
        L.orc_scene_create.restype = ctypes.c_void_p
        L.orc_scene_destroy.argtypes = [ctypes.c_void_p]
        self._h = ctypes.c_void_p(L.orc_scene_create(
            _p(self.vert_grid, _f32p), self.dims[0], self.dims[1], _p(vs, _f32p) if vs is not None else None,
            int(num_vert_simp) if vs is not None else 0, _p(ti, _i32p) if ti is not None else None,
            int(num_tri_simp) if vs is not None else 0))
        self.build_s = last_timing()[0]

    def close(self):
        if getattr(self, "_h", None):
            lib().orc_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def horizon_rows(self, rows, vec_norm, vec_north, offset_0, offset_1, dist_search, azim_num=360, hori_acc=0.25,
                     ray_algorithm="guess_constant", elev_ang_low_lim=-15.0, mask=None, hori_fill=0.0,
                     ray_org_elev=0.01, return_rays=False):
        """Horizon of the inner-domain rows ``rows`` (``vec_norm`` / ``vec_north`` / ``mask`` are the FULL
        inner-domain arrays).  Returns float32 (len(rows), nx, azim_num)."""
        rows = np.ascontiguousarray(rows, np.int32)
        vn = np.ascontiguousarray(vec_norm[rows], np.float32)
        vno = np.ascontiguousarray(vec_north[rows], np.float32)
        nx = vn.shape[1]
        mk = np.ones((len(rows), nx), np.uint8) if mask is None else np.ascontiguousarray(mask[rows], np.uint8)
        hori = np.full((len(rows), nx, azim_num), np.nan, np.float32)
        rays = ctypes.c_ulonglong(0)
        _check(lib().orc_scene_horizon_rows(
            self._h, _p(rows, _i32p), len(rows), nx, _p(vn, _f32p), _p(vno, _f32p), int(offset_0), int(offset_1),
            _p(hori, _f32p), int(azim_num), ctypes.c_float(dist_search), ctypes.c_float(hori_acc),
            ray_algorithm.encode(), ctypes.c_float(elev_ang_low_lim), _p(mk, _u8p), ctypes.c_float(hori_fill),
            ctypes.c_float(ray_org_elev), ctypes.byref(rays)))
        return (hori, rays.value) if return_rays else hori


def horizon_locations(vert_grid, dem_dim_0, dem_dim_1, coords, vec_norm, vec_north, dist_search,
                      azim_num=360, hori_acc=0.25, ray_algorithm="binary_search", geom_type="grid",
                      elev_ang_low_lim=-89.98, ray_org_elev=None, hori_dist_out=False,
                      brute_force=False):
    """Oracle twin of ``horizon.horizon_locations`` (horizon.pyx:218-370)."""
    if ray_org_elev is None:
        ray_org_elev = np.array([0.01], np.float32)
    n = coords.shape[0]
    if len(ray_org_elev) != n:
        ray_org_elev = np.repeat(ray_org_elev, n)
    vert_grid = np.ascontiguousarray(vert_grid, np.float32)
    coords = np.ascontiguousarray(coords, np.float32)
    vec_norm = np.ascontiguousarray(vec_norm, np.float32)
    vec_north = np.ascontiguousarray(vec_north, np.float32)
    ray_org_elev = np.ascontiguousarray(ray_org_elev, np.float32)
    hori = np.full((n, azim_num), np.nan, np.float32)
    dist = np.full((n if hori_dist_out else 1, azim_num), np.nan, np.float32)
    _check(lib().orc_horizon_locations(
        _p(vert_grid, _f32p), int(dem_dim_0), int(dem_dim_1), _p(coords, _f32p),
        _p(vec_norm, _f32p), _p(vec_north, _f32p), _p(hori, _f32p), _p(dist, _f32p), n,
        int(azim_num), ctypes.c_float(dist_search), ctypes.c_float(hori_acc),
        ray_algorithm.encode(), geom_type.encode(), ctypes.c_float(elev_ang_low_lim),
        _p(ray_org_elev, _f32p), int(bool(hori_dist_out)), int(bool(brute_force)), None))
    if hori_dist_out:
        return hori, dist, _azim(azim_num)
    return hori, _azim(azim_num)


class Terrain:
    """Oracle twin of ``shadow.Terrain`` (shadow.pyx:17-200)."""

    def __init__(self):
        self._h = ctypes.c_void_p(lib().orc_terrain_create())

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_terrain_destroy(self._h)
            self._h = None

    def initialise(self, vert_grid, dem_dim_0, dem_dim_1, offset_0, offset_1, vec_tilt, vec_norm,
                   surf_enl_fac, elevation, mask, geom_type="grid", sw_dir_cor_fill=np.nan,
                   ang_max=89.0, refrac_cor=False, brute_force=False):
        a = [np.ascontiguousarray(x, np.float32) for x in
             (vert_grid, vec_tilt, vec_norm, surf_enl_fac, elevation)]
        mask = np.ascontiguousarray(mask, np.uint8)
        _check(lib().orc_terrain_initialise(
            self._h, _p(a[0], _f32p), int(dem_dim_0), int(dem_dim_1), int(offset_0), int(offset_1),
            _p(a[1], _f32p), _p(a[2], _f32p), vec_tilt.shape[0], vec_tilt.shape[1],
            _p(a[3], _f32p), _p(a[4], _f32p), _p(mask, _u8p), geom_type.encode(),
            ctypes.c_float(sw_dir_cor_fill), ctypes.c_float(ang_max), int(bool(refrac_cor)),
            int(bool(brute_force))))

    def shadow(self, sun_position, shadow_buffer):
        sp = np.ascontiguousarray(sun_position, np.float32)
        _check(lib().orc_terrain_shadow(self._h, _p(sp, _f32p), _p(shadow_buffer, _u8p)))

    def sw_dir_cor(self, sun_position, sw_dir_cor_buffer):
        sp = np.ascontiguousarray(sun_position, np.float32)
        _check(lib().orc_terrain_sw_dir_cor(self._h, _p(sp, _f32p), _p(sw_dir_cor_buffer, _f32p)))


def _integral(fn, azim, hori, vec_tilt):
    azim = np.ascontiguousarray(azim, np.float32)
    hori = np.ascontiguousarray(hori, np.float32)
    ny, nx, K = hori.shape
    out = np.empty((ny, nx), np.float32)
    if vec_tilt is None:
        _check(fn(_p(azim, _f32p), _p(hori, _f32p), ny, nx, K, _p(out, _f32p)))
    else:
        vec_tilt = np.ascontiguousarray(vec_tilt, np.float32)
        _check(fn(_p(azim, _f32p), _p(hori, _f32p), _p(vec_tilt, _f32p), ny, nx, K, _p(out, _f32p)))
    return out


def sky_view_factor(azim, hori, vec_tilt):
    return _integral(lib().orc_sky_view_factor, azim, hori, vec_tilt)


def visible_sky_fraction(azim, hori, vec_tilt):
    return _integral(lib().orc_visible_sky_fraction, azim, hori, vec_tilt)


def topographic_openness(azim, hori):
    return _integral(lib().orc_topographic_openness, azim, hori, None)


def _slope(fn, x, y, z, rot_mat, output_rot):
    x, y, z = (np.ascontiguousarray(a, np.float32) for a in (x, y, z))
    ny, nx = x.shape
    out = np.empty((ny, nx, 3), np.float32)
    rm = None if rot_mat is None else np.ascontiguousarray(rot_mat, np.float32)
    _check(fn(_p(x, _f32p), _p(y, _f32p), _p(z, _f32p), None if rm is None else _p(rm, _f32p), ny, nx,
              int(bool(output_rot)), _p(out, _f32p)))
    return out


def slope_plane_meth(x, y, z, rot_mat=None, output_rot=False):
    """Oracle twin of ``topo_param.slope_plane_meth`` (topo_param.pyx:16-225)."""
    return _slope(lib().orc_slope_plane_meth, x, y, z, rot_mat, output_rot)


def slope_vector_meth(x, y, z, rot_mat=None, output_rot=False):
    """Oracle twin of ``topo_param.slope_vector_meth`` (topo_param.pyx:230-372)."""
    return _slope(lib().orc_slope_vector_meth, x, y, z, rot_mat, output_rot)


# ---------------------------------------------------------------------------
# coordinate preparation (scope row 8f-3): call shapes of horayzon.transform /
# horayzon.direction (transform.pyx:15-57, 108-149, 194-228, 266-303, 349-387,
# 490-530; direction.pyx:15-45, 75-122)
# ---------------------------------------------------------------------------
_f64p = ctypes.POINTER(ctypes.c_double)


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64).ravel()


def _c32(a):
    return np.ascontiguousarray(a, dtype=np.float32).ravel()


def lonlat2ecef(lon, lat, h, ellps):
    a, b, c = _c64(lon), _c64(lat), _c32(h)
    x, y, z = (np.empty(a.size, np.float64) for _ in range(3))
    _check(lib().orc_lonlat2ecef(_p(a, _f64p), _p(b, _f64p), _p(c, _f32p), ctypes.c_longlong(a.size), ellps.encode(),
                                 _p(x, _f64p), _p(y, _f64p), _p(z, _f64p)))
    return x.reshape(lon.shape), y.reshape(lon.shape), z.reshape(lon.shape)


def ecef2enu(x_ecef, y_ecef, z_ecef, trans):
    a, b, c = _c64(x_ecef), _c64(y_ecef), _c64(z_ecef)
    x, y, z = (np.empty(a.size, np.float32) for _ in range(3))
    _check(lib().orc_ecef2enu(_p(a, _f64p), _p(b, _f64p), _p(c, _f64p), ctypes.c_longlong(a.size),
                              ctypes.c_double(trans.x_ecef_or), ctypes.c_double(trans.y_ecef_or),
                              ctypes.c_double(trans.z_ecef_or), ctypes.c_double(trans.lon_or), ctypes.c_double(trans.lat_or),
                              _p(x, _f32p), _p(y, _f32p), _p(z, _f32p)))
    return x.reshape(x_ecef.shape), y.reshape(x_ecef.shape), z.reshape(x_ecef.shape)


def ecef2enu_vector(vec_ecef, trans):
    v = _c32(vec_ecef)
    o = np.empty(v.size, np.float32)
    _check(lib().orc_ecef2enu_vector(_p(v, _f32p), ctypes.c_longlong(v.size // 3), ctypes.c_double(trans.lon_or),
                                     ctypes.c_double(trans.lat_or), _p(o, _f32p)))
    return o.reshape(vec_ecef.shape)


def surf_norm(lon, lat):
    a, b = _c64(lon), _c64(lat)
    o = np.empty(a.size * 3, np.float32)
    _check(lib().orc_surf_norm(_p(a, _f64p), _p(b, _f64p), ctypes.c_longlong(a.size), _p(o, _f32p)))
    return o.reshape(lon.shape + (3,))


def north_dir(x_ecef, y_ecef, z_ecef, vec_norm_ecef, ellps):
    a, b, c, v = _c64(x_ecef), _c64(y_ecef), _c64(z_ecef), _c32(vec_norm_ecef)
    o = np.empty(a.size * 3, np.float32)
    _check(lib().orc_north_dir(_p(a, _f64p), _p(b, _f64p), _p(c, _f64p), _p(v, _f32p), ctypes.c_longlong(a.size),
                               ellps.encode(), _p(o, _f32p)))
    return o.reshape(x_ecef.shape + (3,))


def wgs2swiss(lon, lat, h_wgs):
    a, b, c = _c64(lon), _c64(lat), _c32(h_wgs)
    e, n = np.empty(a.size, np.float64), np.empty(a.size, np.float64)
    h = np.empty(a.size, np.float32)
    _check(lib().orc_wgs2swiss(_p(a, _f64p), _p(b, _f64p), _p(c, _f32p), ctypes.c_longlong(a.size), _p(e, _f64p),
                               _p(n, _f64p), _p(h, _f32p)))
    return e.reshape(lon.shape), n.reshape(lon.shape), h.reshape(lon.shape)


def swiss2wgs(e, n, h_ch):
    a, b, c = _c64(e), _c64(n), _c32(h_ch)
    lon, lat = np.empty(a.size, np.float64), np.empty(a.size, np.float64)
    h = np.empty(a.size, np.float32)
    _check(lib().orc_swiss2wgs(_p(a, _f64p), _p(b, _f64p), _p(c, _f32p), ctypes.c_longlong(a.size), _p(lon, _f64p),
                               _p(lat, _f64p), _p(h, _f32p)))
    return lon.reshape(e.shape), lat.reshape(e.shape), h.reshape(e.shape)


def rotation_matrix_glob2loc(vec_north_enu, vec_norm_enu):
    ny, nx = vec_north_enu.shape[:2]
    a, b = _c32(vec_north_enu), _c32(vec_norm_enu)
    o = np.empty((ny + 2, nx + 2, 3, 3), np.float32)
    _check(lib().orc_rotation_matrix_glob2loc(_p(a, _f32p), _p(b, _f32p), ny, nx, _p(o, _f32p)))
    return o


def selftest_state_machine(vert_grid, dem_dim_0, dem_dim_1, offset_0, offset_1, dim_in_0, dim_in_1, azim_num,
                           dist_search, hori_acc=0.25, elev_ang_low_lim=-15.0, ray_org_elev=0.01,
                           ray_algorithm="guess_constant", refuse_every=0):
    """Drive the PRODUCT's search state machine (csrc/hzb_search.cuh, host build) with the oracle's casts.
    Returns (cells that differ from the oracle's algorithm, reference casts, state-machine casts,
    companion results consumed)."""
    L = lib()
    L.orc_selftest_state_machine.restype = ctypes.c_longlong
    a, b, c = ctypes.c_longlong(0), ctypes.c_longlong(0), ctypes.c_longlong(0)
    vg = np.ascontiguousarray(vert_grid, dtype=np.float32)
    bad = L.orc_selftest_state_machine(_p(vg, _f32p), int(dem_dim_0), int(dem_dim_1), int(offset_0), int(offset_1),
                                       int(dim_in_0), int(dim_in_1), int(azim_num), ctypes.c_float(dist_search),
                                       ctypes.c_float(hori_acc), ctypes.c_float(elev_ang_low_lim),
                                       ctypes.c_float(ray_org_elev), ray_algorithm.encode(), int(refuse_every),
                                       ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    return int(bad), a.value, b.value, c.value



def selftest_segments(vert_grid, dem_dim_0, dem_dim_1, offset_0, offset_1, dim_in_0, dim_in_1, azim_num, dist_search,
                      hori_acc=0.25, elev_ang_low_lim=-15.0, ray_org_elev=0.01, ray_algorithm="guess_constant",
                      segments=4, tilted=True):
    """The PRODUCT's state machine in azimuth-segment mode (csrc/hzb_search.cuh) on the oracle's casts: every
    segment of every cell runs as a task of its own.  Returns (cells whose verified segments differ from the chain,
    reference casts, segment tasks, prelude guesses that missed the chain's index, prelude casts)."""
    L = lib()
    L.orc_selftest_segments.restype = ctypes.c_longlong
    a = ctypes.c_longlong(0)
    st = (ctypes.c_longlong * 3)()
    vg = np.ascontiguousarray(vert_grid, dtype=np.float32)
    bad = L.orc_selftest_segments(_p(vg, _f32p), int(dem_dim_0), int(dem_dim_1), int(offset_0), int(offset_1),
                                  int(dim_in_0), int(dim_in_1), int(azim_num), ctypes.c_float(dist_search),
                                  ctypes.c_float(hori_acc), ctypes.c_float(elev_ang_low_lim),
                                  ctypes.c_float(ray_org_elev), ray_algorithm.encode(), int(segments), int(bool(tilted)),
                                  ctypes.byref(a), st)
    return int(bad), a.value, int(st[0]), int(st[1]), int(st[2])


def selftest_queue(seed=1, iters=2000):
    """Enumerate the PRODUCT's work queue (csrc/hzb_queue.cuh, host build) for random launch geometries; returns the
    number of violations (tiles / segments not covered exactly once, predicates that disagree with the enumeration)."""
    L = lib()
    L.orc_selftest_queue.restype = ctypes.c_longlong
    return int(L.orc_selftest_queue(ctypes.c_ulonglong(seed), ctypes.c_longlong(iters)))
