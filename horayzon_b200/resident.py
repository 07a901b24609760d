"""Resident (device-pointer) tier of the C ABI, for benchmarking and sharding.

Thin ctypes binding of the ``hzb_scene_*`` / ``*_dev`` entry points of
``libhorayzon_b200.so`` (``include/horayzon_b200.h``).  Device memory and
streams come from PyTorch (plumbing only): tensors are passed as raw device
pointers, streams as ``cudaStream_t``.  These calls have no counterpart in the
reference; they keep DEM, BVH and outputs in HBM between calls and let several
processes shard the rows of the inner domain (the reference's own partitioning
axis, ``horizon_comp.cpp:739-744``) over the GPUs of one box.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("HZB_LIB", os.path.join(_HERE, "libhorayzon_b200.so"))  # HZB_LIB: A/B experiments only


class Stats(ctypes.Structure):
    _fields_ = [("rays", ctypes.c_ulonglong), ("node_visits", ctypes.c_ulonglong),
                ("prim_tests", ctypes.c_ulonglong), ("units", ctypes.c_ulonglong),
                ("warp_node_visits", ctypes.c_ulonglong),
                ("t_h2d", ctypes.c_double), ("t_build", ctypes.c_double), ("t_trace", ctypes.c_double),
                ("t_d2h", ctypes.c_double), ("t_total", ctypes.c_double),
                ("num_prims", ctypes.c_ulonglong), ("num_nodes", ctypes.c_ulonglong),
                ("bvh_bytes", ctypes.c_ulonglong), ("fallback_packets", ctypes.c_ulonglong),
                ("segment_tasks", ctypes.c_ulonglong), ("segment_redos", ctypes.c_ulonglong)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def lib():
    """The C-ABI library (loaded once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError("libhorayzon_b200.so is not built; run `python horayzon_b200/_build.py`")
        L = ctypes.CDLL(_LIB_PATH)
        L.hzb_last_error.restype = ctypes.c_char_p
        L.hzb_version.restype = ctypes.c_char_p
        L.hzb_scene_create.restype = ctypes.c_void_p
        L.hzb_scene_create.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                       ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.hzb_scene_destroy.argtypes = [ctypes.c_void_p]
        L.hzb_scene_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(Stats)]
        L.hzb_get_stats.argtypes = [ctypes.POINTER(Stats)]
        L.hzb_horizon_gridded_dev.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float,
            ctypes.c_char_p, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]
        L.hzb_horizon_gridded_dev_layout.argtypes = L.hzb_horizon_gridded_dev.argtypes[:-1] + [ctypes.c_int, ctypes.c_void_p]
        L.hzb_horizon_gridded_dev_sharded.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_char_p, ctypes.c_float,
            ctypes.c_float, ctypes.c_float, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.hzb_debug_option.argtypes = [ctypes.c_char_p, ctypes.c_int]
        L.hzb_set_device.argtypes = [ctypes.c_int]
        L.hzb_trim.restype = None
        for name in ("hzb_sky_view_factor_dev", "hzb_visible_sky_fraction_dev"):
            getattr(L, name).argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong,
                                         ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.hzb_topographic_openness_dev.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong,
                                                   ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _lib = L
    return _lib


def debug_option(name, value):
    """Test-only switch of the library (``hzb_debug_option``): second implementations, tuning knobs."""
    _check(lib().hzb_debug_option(name.encode(), int(value)))


def trim():
    """Release the idle pooled device / page-locked memory of this process (``hzb_trim``)."""
    lib().hzb_trim()


def last_error():
    return lib().hzb_last_error().decode("utf-8", "replace")


def _check(rc):
    if rc != 0:
        raise RuntimeError("horayzon_b200: " + last_error())


def last_stats():
    """Counters / timings of the most recent host-tier call on this thread."""
    st = Stats()
    _check(lib().hzb_get_stats(ctypes.byref(st)))
    return st.asdict()


def _np_ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def _stream_ptr(stream):
    if stream is None:
        import torch
        stream = torch.cuda.current_stream()
    return ctypes.c_void_p(int(stream.cuda_stream))


class Scene:
    """DEM (+ optional TIN) and its BVH resident on one GPU."""

    def __init__(self, vert_grid, dem_dim_0, dem_dim_1, vert_simp=None, num_vert_simp=0,
                 tri_ind_simp=None, num_tri_simp=0, device=0):
        vert_grid = np.ascontiguousarray(vert_grid, np.float32)
        if len(vert_grid) < dem_dim_0 * dem_dim_1 * 3:
            raise ValueError("inconsistency between input arguments vert_grid, dem_dim_0 and dem_dim_1")
        vs = ti = None
        if vert_simp is not None and num_vert_simp >= 3:
            vs = np.ascontiguousarray(vert_simp, np.float32)
            ti = np.ascontiguousarray(tri_ind_simp, np.int32)
        self.dem_dim_0, self.dem_dim_1, self.device = int(dem_dim_0), int(dem_dim_1), int(device)
        h = lib().hzb_scene_create(_np_ptr(vert_grid), int(dem_dim_0), int(dem_dim_1),
                                   _np_ptr(vs) if vs is not None else None, int(num_vert_simp) if vs is not None else 0,
                                   _np_ptr(ti) if ti is not None else None, int(num_tri_simp) if vs is not None else 0,
                                   int(device))
        if not h:
            raise RuntimeError("horayzon_b200: " + last_error())
        self._h = ctypes.c_void_p(h)

    def close(self):
        if getattr(self, "_h", None):
            lib().hzb_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown: module globals may be gone
            pass

    def stats(self):
        st = Stats()
        _check(lib().hzb_scene_stats(self._h, ctypes.byref(st)))
        return st.asdict()

    def horizon_gridded(self, vec_norm, vec_north, mask, offset_0, offset_1, hori_out, row_begin=0,
                        row_end=None, dist_search=50.0, hori_acc=0.25, ray_algorithm="guess_constant",
                        elev_ang_low_lim=-15.0, hori_fill=0.0, ray_org_elev=0.01, stream=None, azim_first=False):
        """Asynchronous horizon computation for inner-domain rows
        ``[row_begin, row_end)``.  ``vec_norm`` / ``vec_north`` (ny, nx, 3) float32,
        ``mask`` (ny, nx) uint8 and ``hori_out`` (ny, nx, K) float32 -- or (K, ny, nx)
        with ``azim_first=True`` -- are CUDA tensors of the FULL inner domain; only the
        selected rows are touched."""
        if azim_first:
            K, ny, nx = hori_out.shape
        else:
            ny, nx, K = hori_out.shape
        if row_end is None:
            row_end = ny
        assert vec_norm.is_cuda and vec_north.is_cuda and mask.is_cuda and hori_out.is_cuda
        assert vec_norm.is_contiguous() and vec_north.is_contiguous() and mask.is_contiguous() and hori_out.is_contiguous()
        assert tuple(vec_norm.shape) == (ny, nx, 3) and tuple(mask.shape) == (ny, nx)
        _check(lib().hzb_horizon_gridded_dev_layout(
            self._h, ctypes.c_void_p(vec_norm.data_ptr()), ctypes.c_void_p(vec_north.data_ptr()),
            ctypes.c_void_p(mask.data_ptr()), int(offset_0), int(offset_1), int(ny), int(nx), int(row_begin),
            int(row_end), int(K), float(dist_search), float(hori_acc), ray_algorithm.encode(),
            float(elev_ang_low_lim), float(hori_fill), float(ray_org_elev),
            ctypes.c_void_p(hori_out.data_ptr()), 1 if azim_first else 0, _stream_ptr(stream)))


    def horizon_gridded_sharded(self, vec_norm, vec_north, mask, offset_0, offset_1, hori_out, shard_rank, shard_count,
                                azim_num, packed=True, dist_search=50.0, hori_acc=0.25, ray_algorithm="guess_constant",
                                elev_ang_low_lim=-15.0, hori_fill=0.0, ray_org_elev=0.01, stream=None):
        """Asynchronous horizon computation for the 4-row blocks ``b`` of the inner domain with
        ``b % shard_count == shard_rank`` (``hzb_horizon_gridded_dev_sharded``).  ``packed=True``: ``hori_out`` is this
        shard's send buffer ``(shard_rows(ny, rank, count), nx, K)``; ``packed=False``: the full ``(ny, nx, K)`` array."""
        ny, nx = int(mask.shape[0]), int(mask.shape[1])
        assert vec_norm.is_cuda and vec_north.is_cuda and mask.is_cuda and hori_out.is_cuda
        assert vec_norm.is_contiguous() and vec_north.is_contiguous() and mask.is_contiguous() and hori_out.is_contiguous()
        need = shard_rows(ny, shard_rank, shard_count) if packed else ny
        assert hori_out.numel() >= need * nx * int(azim_num), "output buffer too small for this shard"
        _check(lib().hzb_horizon_gridded_dev_sharded(
            self._h, ctypes.c_void_p(vec_norm.data_ptr()), ctypes.c_void_p(vec_north.data_ptr()),
            ctypes.c_void_p(mask.data_ptr()), int(offset_0), int(offset_1), ny, nx, int(azim_num), float(dist_search),
            float(hori_acc), ray_algorithm.encode(), float(elev_ang_low_lim), float(hori_fill), float(ray_org_elev),
            ctypes.c_void_p(hori_out.data_ptr()), int(shard_rank), int(shard_count), 1 if packed else 0,
            _stream_ptr(stream)))


def shard_rows(num_rows, shard_rank, shard_count):
    """Rows (whole 4-row blocks) in the packed buffer of one shard (``hzb_shard_rows``)."""
    return int(lib().hzb_shard_rows(int(num_rows), int(shard_rank), int(shard_count)))


def sky_view_factor_dev(azim, hori, vec_tilt, out, stream=None):
    cells = hori.numel() // hori.shape[-1]
    _check(lib().hzb_sky_view_factor_dev(ctypes.c_void_p(azim.data_ptr()), ctypes.c_void_p(hori.data_ptr()),
                                         ctypes.c_void_p(vec_tilt.data_ptr()), cells, int(hori.shape[-1]),
                                         ctypes.c_void_p(out.data_ptr()), _stream_ptr(stream)))


def visible_sky_fraction_dev(azim, hori, vec_tilt, out, stream=None):
    cells = hori.numel() // hori.shape[-1]
    _check(lib().hzb_visible_sky_fraction_dev(ctypes.c_void_p(azim.data_ptr()), ctypes.c_void_p(hori.data_ptr()),
                                              ctypes.c_void_p(vec_tilt.data_ptr()), cells, int(hori.shape[-1]),
                                              ctypes.c_void_p(out.data_ptr()), _stream_ptr(stream)))


def topographic_openness_dev(azim, hori, out, stream=None):
    cells = hori.numel() // hori.shape[-1]
    _check(lib().hzb_topographic_openness_dev(ctypes.c_void_p(azim.data_ptr()), ctypes.c_void_p(hori.data_ptr()),
                                              cells, int(hori.shape[-1]), ctypes.c_void_p(out.data_ptr()),
                                              _stream_ptr(stream)))


def prep_enu_dev(lon, lat, elev, ellps, trans, offset_0, offset_1, dim_in_0, dim_in_1, vert_grid, vec_norm=None,
                 vec_north=None, stream=None):
    """Fused coordinate preparation in HBM (``hzb_prep_enu_dev``): ``lon`` [nx], ``lat`` [ny] (float64 tensors),
    ``elev`` [ny][nx] (float32) -> ``vert_grid`` [ny][nx][3] and the inner-domain ``vec_norm`` / ``vec_north``
    [dim_in_0][dim_in_1][3] in ENU coordinates.  ``trans`` carries the TransformerEcef2enu attributes."""
    L = lib()
    L.hzb_prep_enu_dev.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_char_p, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                   ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    ny, nx = int(elev.shape[0]), int(elev.shape[1])
    _check(L.hzb_prep_enu_dev(lon.data_ptr(), lat.data_ptr(), elev.data_ptr(), ny, nx, ellps.encode(),
                              float(trans.x_ecef_or), float(trans.y_ecef_or), float(trans.z_ecef_or),
                              float(trans.lon_or), float(trans.lat_or), int(offset_0), int(offset_1), int(dim_in_0),
                              int(dim_in_1), vert_grid.data_ptr(),
                              vec_norm.data_ptr() if vec_norm is not None else None,
                              vec_north.data_ptr() if vec_north is not None else None, _stream_ptr(stream)))

