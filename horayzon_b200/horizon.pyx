# cython: language_level=3
"""Horizon computation -- drop-in for ``horayzon.horizon``.

Same Python signatures, defaults, validation and return values as the
reference wrapper (``horayzon/horizon.pyx:29-197`` and ``:218-370``), but the
work is done by the C-ABI CUDA library ``libhorayzon_b200.so``
(``include/horayzon_b200.h``) instead of Embree/TBB
(``horayzon/horizon_comp.cpp``).  There is no CPU fallback: without a CUDA
device the calls raise ``RuntimeError``.
"""
cimport numpy as np
import numpy as np
from libc.stdint cimport int32_t, uint8_t

np.import_array()

from cpython.ref cimport Py_INCREF
import os

cdef extern from "horayzon_b200.h":
    const char* hzb_last_error()
    void* hzb_host_alloc(size_t nbytes) nogil
    void hzb_host_free(void* p) nogil
    int hzb_horizon_gridded(
        const float* vert_grid, int dem_dim_0, int dem_dim_1,
        const float* vec_norm, const float* vec_north,
        int offset_0, int offset_1, float* hori_buffer,
        int dim_in_0, int dim_in_1, int azim_num, float dist_search,
        float hori_acc, const char* ray_algorithm, const char* geom_type,
        const float* vert_simp, int num_vert_simp,
        const int32_t* tri_ind_simp, int num_tri_simp,
        float elev_ang_low_lim, const uint8_t* mask, float hori_fill,
        float ray_org_elev) nogil
    int hzb_horizon_gridded_layout(
        const float* vert_grid, int dem_dim_0, int dem_dim_1,
        const float* vec_norm, const float* vec_north,
        int offset_0, int offset_1, float* hori_buffer,
        int dim_in_0, int dim_in_1, int azim_num, float dist_search,
        float hori_acc, const char* ray_algorithm, const char* geom_type,
        const float* vert_simp, int num_vert_simp,
        const int32_t* tri_ind_simp, int num_tri_simp,
        float elev_ang_low_lim, const uint8_t* mask, float hori_fill,
        float ray_org_elev, int azim_first) nogil
    int hzb_horizon_gridded_svf(
        const float* vert_grid, int dem_dim_0, int dem_dim_1,
        const float* vec_norm, const float* vec_north,
        int offset_0, int offset_1, float* hori_buffer,
        int dim_in_0, int dim_in_1, int azim_num, float dist_search,
        float hori_acc, const char* ray_algorithm, const char* geom_type,
        const float* vert_simp, int num_vert_simp,
        const int32_t* tri_ind_simp, int num_tri_simp,
        float elev_ang_low_lim, const uint8_t* mask, float hori_fill,
        float ray_org_elev, const float* vec_tilt, float* svf_buffer) nogil
    int hzb_horizon_gridded_quantised(
        const float* vert_grid, int dem_dim_0, int dem_dim_1,
        const float* vec_norm, const float* vec_north,
        int offset_0, int offset_1, unsigned short* idx_buffer, float* first_buffer,
        int dim_in_0, int dim_in_1, int azim_num, float dist_search,
        float hori_acc, const char* geom_type,
        const float* vert_simp, int num_vert_simp,
        const int32_t* tri_ind_simp, int num_tri_simp,
        float elev_ang_low_lim, const uint8_t* mask, float hori_fill,
        float ray_org_elev) nogil
    int hzb_horizon_tables(int azim_num, float dist_search, float hori_acc, float elev_ang_low_lim, int cap,
                           float* elev_ang, float* elev_sin, float* elev_cos, float* azim_sin, float* azim_cos) nogil
    int hzb_horizon_gridded_multi(
        const float* vert_grid, int dem_dim_0, int dem_dim_1,
        const float* vec_norm, const float* vec_north,
        int offset_0, int offset_1, float* hori_buffer,
        int dim_in_0, int dim_in_1, int azim_num, float dist_search,
        float hori_acc, const char* ray_algorithm, const char* geom_type,
        const float* vert_simp, int num_vert_simp,
        const int32_t* tri_ind_simp, int num_tri_simp,
        float elev_ang_low_lim, const uint8_t* mask, float hori_fill,
        float ray_org_elev, const float* vec_tilt, float* svf_buffer,
        int n_devices, int n_shards, int device_gather) nogil
    int hzb_horizon_locations(
        const float* vert_grid, int dem_dim_0, int dem_dim_1,
        const float* coords, const float* vec_norm, const float* vec_north,
        float* hori_buffer, float* hori_dist_buffer, int num_loc,
        int azim_num, float dist_search, float hori_acc,
        const char* ray_algorithm, const char* geom_type,
        float elev_ang_low_lim, const float* ray_org_elev,
        int hori_dist_out) nogil

_ALGORITHMS = ("discrete_sampling", "binary_search", "guess_constant")
_GEOM_TYPES = ("triangle", "quad", "grid")


def _azimuth_axis(int azim_num):
    # float32 azimuth of every sector, as horizon.pyx:190-195 rebuilds it
    cdef np.ndarray[np.float32_t, ndim = 1, mode = "c"] azim = np.empty(azim_num, dtype=np.float32)
    cdef int i
    for i in range(azim_num):
        azim[i] = ((2 * np.pi) / azim_num * i)
    return azim


def _raise_native():
    raise RuntimeError("horayzon_b200: " + hzb_last_error().decode("utf-8", "replace"))


cdef class _PinnedBlock:
    """Owner of one pooled page-locked block (``hzb_host_alloc``); the ndarray built on it
    keeps it alive through ``base`` and the block returns to the pool with the array."""
    cdef void* p

    def __dealloc__(self):
        if self.p != NULL:
            hzb_host_free(self.p)
            self.p = NULL


_big_outputs = 0   # large outputs handed out by this process so far


cdef object _output_array(tuple shape):
    """float32 C-contiguous output array.  Large arrays (the 2 GB horizon of a 1201 x 1201 x 360
    run) can be backed by pooled page-locked memory so that the device writes / reads them by
    DMA (0.1 s per call for that size); page-locking a fresh block costs more than it saves
    (0.8 s for 2 GB), so by default it starts with the SECOND large output of a process -- a
    single-shot script never pays for it, a loop pays once.  ``HZB_PINNED_OUTPUT=1`` / ``0``
    force it on / off.  Anything else, or any failure, is plain ``np.empty`` as in the
    reference wrapper (horizon.pyx:170-173)."""
    global _big_outputs
    cdef size_t n = 4
    for d in shape:
        n *= <size_t> d
    cdef void* p = NULL
    cdef _PinnedBlock blk
    cdef np.npy_intp dims[3]
    cdef np.ndarray arr
    # between 64 MB and 8 GB: smaller arrays do not matter, larger ones (the 52 GB horizon of a 6000 x 6000 x 360 run)
    # would keep too much memory page-locked in the pool and take tens of seconds to lock
    cdef bint big = n >= (<size_t> 64 << 20) and n <= (<size_t> 8 << 30) and len(shape) == 3
    mode = os.environ.get("HZB_PINNED_OUTPUT", "auto")
    cdef bint use_pinned = big and (mode == "1" or (mode != "0" and _big_outputs >= 1))
    if big:
        _big_outputs += 1
    if use_pinned:
        with nogil:
            p = hzb_host_alloc(n)
        if p != NULL:
            blk = _PinnedBlock.__new__(_PinnedBlock)
            blk.p = p
            dims[0] = shape[0]; dims[1] = shape[1]; dims[2] = shape[2]
            arr = np.PyArray_SimpleNewFromData(3, dims, np.NPY_FLOAT32, p)
            Py_INCREF(blk)                       # PyArray_SetBaseObject steals a reference
            np.PyArray_SetBaseObject(arr, blk)
            return arr
    return np.empty(shape, dtype=np.float32)


def horizon_gridded(
        np.ndarray[np.float32_t, ndim = 1] vert_grid,
        int dem_dim_0, int dem_dim_1,
        np.ndarray[np.float32_t, ndim = 3] vec_norm,
        np.ndarray[np.float32_t, ndim = 3] vec_north,
        int offset_0, int offset_1,
        float dist_search,
        int azim_num=360,
        float hori_acc=0.25,
        str ray_algorithm="guess_constant",
        str geom_type="grid",
        np.ndarray[np.float32_t, ndim = 1]
        vert_simp=np.array([0.0, 0.0, 0.0, 0.0], dtype=np.float32),
        int num_vert_simp=1,
        np.ndarray[np.int32_t, ndim = 1]
        tri_ind_simp=np.array([0, 0, 0, 0], dtype=np.int32),
        int num_tri_simp=1,
        float elev_ang_low_lim = -15.0,
        np.ndarray[np.uint8_t, ndim = 2] mask=None,
        float hori_fill=0.0,
        float ray_org_elev=0.01,
        bint azim_first=False,
        np.ndarray[np.float32_t, ndim = 3] svf_vec_tilt=None,
        int devices=1,
        bint device_gather=False,
        int _shards=0):
    """Horizon of every unmasked cell of a gridded inner domain.

    Arguments, units and defaults are those of ``horayzon.horizon.horizon_gridded``
    (``horizon.pyx:29-106``): ``vert_grid`` flat float32 vertex buffer [m],
    ``vec_norm`` / ``vec_north`` float32 (y, x, 3), ``dist_search`` [km],
    ``hori_acc`` and ``elev_ang_low_lim`` [degree], ``hori_fill`` [radian],
    ``ray_org_elev`` [m].  ``geom_type`` is accepted for compatibility; the three
    Embree geometry types describe the same surface and share one GPU BVH.

    Returns ``(hori_buffer, azim)``: float32 (y, x, azim_num) horizon [radian]
    and float32 (azim_num,) azimuth [radian].

    Additive keyword (not in the reference): ``azim_first=True`` returns the horizon as
    (azim_num, y, x) -- what the reference's examples produce with
    ``np.moveaxis(hori, 2, 0)`` before writing NetCDF
    (``examples/horizon/gridded_curved_DEM.py:113-125``) -- written in that order by
    the kernel, with the same values.

    Additive keyword: ``svf_vec_tilt`` (float32 (y, x, 3), the ``vec_tilt`` argument of
    ``topo_param.sky_view_factor``) makes the call return ``(hori_buffer, azim, svf)``: the sky
    view factor is integrated on the device right behind the search, from the horizon array
    still resident in HBM -- the same values as calling ``topo_param.sky_view_factor(azim, hori,
    vec_tilt)`` afterwards (``examples/horizon/gridded_curved_DEM.py:104-144``), without uploading
    the horizon array again.

    Additive keyword: ``devices`` -- number of GPUs of this box to use (``0``: all visible; default ``1``).
    With more than one, the 4-row blocks of the inner domain are dealt out to the GPUs in turn (the
    reference's TBB row partition, ``horizon_comp.cpp:739-744``), one host thread per GPU, DEM and BVH
    replicated; the result is identical to the single-GPU call.  ``device_gather=True`` joins the shards by
    one NCCL all-gather over NVLink instead of per-GPU copies to the host array.
    """
    # argument checks, in the reference's order and wording (horizon.pyx:109-156)
    if len(vert_grid) < (dem_dim_0 * dem_dim_1 * 3):
        raise ValueError("inconsistency between input arguments vert_grid, "
                         "dem_dim_0 and dem_dim_1")
    if ((offset_0 + vec_norm.shape[0] > dem_dim_0)
            or (offset_1 + vec_norm.shape[1] > dem_dim_1)):
        raise ValueError("inconsistency between input arguments dem_dim_0, "
                         "dem_dim_1, offset_0, offset_1 and vec_norm")
    if ((vec_norm.ndim != 3) or (vec_north.ndim != 3)
            or (vec_norm.shape[0] != vec_north.shape[0])
            or (vec_norm.shape[1] != vec_north.shape[1])
            or (vec_norm.shape[2] != vec_north.shape[2])):
        raise ValueError("dimension (lengths) of vec_norm and/or vec_north "
                         "is/are erroneous")
    if ray_algorithm not in _ALGORITHMS:
        raise ValueError("invalid input argument for ray_algorithm")
    if geom_type not in _GEOM_TYPES:
        raise ValueError("invalid input argument for geom_type")
    if len(vert_simp) < (num_vert_simp * 3):
        raise ValueError("inconsistency between input arguments vert_simp "
                         "and num_vert_simp")
    if len(tri_ind_simp) < (num_tri_simp * 3):
        raise ValueError("inconsistency between input arguments tri_ind_simp "
                         "and num_tri_simp")
    if tri_ind_simp.max() > (num_vert_simp - 1):
        raise ValueError("triangle indices of simplified outer domain exceed "
                         "number of vertices")
    if hori_acc > 10.0:
        raise ValueError("limit of hori_acc (10 degree) is exceeded")
    if mask is None:
        mask = np.ones((vec_norm.shape[0], vec_norm.shape[1]), dtype=np.uint8)
    if (mask.shape[0] != vec_norm.shape[0]) \
            or (mask.shape[1] != vec_norm.shape[1]):
        raise ValueError("shape of mask is inconsistent with other input")
    if mask.dtype != "uint8":
        raise TypeError("data type of mask must be 'uint8'")
    if ray_org_elev < 0.005:
        raise TypeError("minimal allowed value for 'ray_org_elev' is 0.005 m")
    if (dem_dim_0 > 32767) or (dem_dim_1 > 32767):
        raise ValueError("maximal allowed input length for dem_dim_0 and "
                         "dem_dim_1 is 32'767")
    if vert_simp.nbytes > (16.0 * 10 ** 9):
        raise ValueError("vertex buffer vert_simp is larger than 16 GB")
    if tri_ind_simp.min() < 0:
        raise ValueError("triangle indices of simplified outer domain must not "
                         "be negative")
    if svf_vec_tilt is not None:
        if azim_first:
            raise ValueError("svf_vec_tilt needs the reference layout (azim_first=False)")
        if ((svf_vec_tilt.shape[0] != vec_norm.shape[0]) or (svf_vec_tilt.shape[1] != vec_norm.shape[1])
                or (svf_vec_tilt.shape[2] != 3)):   # topo_param.pyx:400-405
            raise ValueError("Input array(s) has/have incorrect shape(s)")
        if azim_num < 2:
            raise ValueError("the sky view factor needs at least two azimuth sectors")

    cdef np.ndarray[np.float32_t, ndim = 1, mode = "c"] vg = np.ascontiguousarray(vert_grid)
    cdef np.ndarray[np.float32_t, ndim = 3, mode = "c"] vn = np.ascontiguousarray(vec_norm)
    cdef np.ndarray[np.float32_t, ndim = 3, mode = "c"] vno = np.ascontiguousarray(vec_north)
    cdef np.ndarray[np.float32_t, ndim = 1, mode = "c"] vs = np.ascontiguousarray(vert_simp)
    cdef np.ndarray[np.int32_t, ndim = 1, mode = "c"] ti = np.ascontiguousarray(tri_ind_simp)
    cdef np.ndarray[np.uint8_t, ndim = 2, mode = "c"] mk = np.ascontiguousarray(mask)
    cdef bytes alg_b = ray_algorithm.encode("utf-8")
    cdef bytes geom_b = geom_type.encode("utf-8")
    cdef const char* alg_c = alg_b
    cdef const char* geom_c = geom_b
    cdef int ny = vn.shape[0], nx = vn.shape[1]

    cdef np.ndarray[np.float32_t, ndim = 3, mode = "c"] hori_buffer = \
        _output_array((azim_num, ny, nx) if azim_first else (ny, nx, azim_num))
    cdef int layout = 1 if azim_first else 0
    # The reference pre-fills NaN (horizon.pyx:170-173).  The native call writes every
    # element (masked cells get hori_fill), so the 2 GB-scale fill pass is skipped and
    # the pages are first touched by the overlapped device-to-host copy instead.
    cdef int rc = 0
    cdef np.ndarray[np.float32_t, ndim = 3, mode = "c"] tilt
    cdef np.ndarray[np.float32_t, ndim = 2, mode = "c"] svf
    cdef const float* tilt_p = NULL
    cdef float* svf_p = NULL
    cdef int gather = 1 if device_gather else 0
    if devices != 1 or _shards > 0:
        if azim_first:
            raise ValueError("devices != 1 needs the reference layout (azim_first=False)")
        if devices < 0:
            raise ValueError("devices must be >= 0")
        svf = None
        if svf_vec_tilt is not None:
            tilt = np.ascontiguousarray(svf_vec_tilt)
            svf = np.empty((ny, nx), dtype=np.float32)
            tilt_p = <const float*> tilt.data
            svf_p = <float*> svf.data
        if ny > 0 and nx > 0:
            with nogil:
                rc = hzb_horizon_gridded_multi(
                    <const float*> vg.data, dem_dim_0, dem_dim_1,
                    <const float*> vn.data, <const float*> vno.data,
                    offset_0, offset_1, <float*> hori_buffer.data, ny, nx,
                    azim_num, dist_search, hori_acc, alg_c, geom_c,
                    <const float*> vs.data, num_vert_simp,
                    <const int32_t*> ti.data, num_tri_simp,
                    elev_ang_low_lim, <const uint8_t*> mk.data, hori_fill,
                    ray_org_elev, tilt_p, svf_p, devices, _shards, gather)
        if rc != 0:
            _raise_native()
        if svf is not None:
            return hori_buffer, _azimuth_axis(azim_num), svf
        return hori_buffer, _azimuth_axis(azim_num)
    if svf_vec_tilt is not None:
        tilt = np.ascontiguousarray(svf_vec_tilt)
        svf = np.empty((ny, nx), dtype=np.float32)
        if ny > 0 and nx > 0:
            with nogil:
                rc = hzb_horizon_gridded_svf(
                    <const float*> vg.data, dem_dim_0, dem_dim_1,
                    <const float*> vn.data, <const float*> vno.data,
                    offset_0, offset_1, <float*> hori_buffer.data, ny, nx,
                    azim_num, dist_search, hori_acc, alg_c, geom_c,
                    <const float*> vs.data, num_vert_simp,
                    <const int32_t*> ti.data, num_tri_simp,
                    elev_ang_low_lim, <const uint8_t*> mk.data, hori_fill,
                    ray_org_elev, <const float*> tilt.data, <float*> svf.data)
        if rc != 0:
            _raise_native()
        return hori_buffer, _azimuth_axis(azim_num), svf
    if ny > 0 and nx > 0:
        with nogil:
            rc = hzb_horizon_gridded_layout(
                <const float*> vg.data, dem_dim_0, dem_dim_1,
                <const float*> vn.data, <const float*> vno.data,
                offset_0, offset_1, <float*> hori_buffer.data, ny, nx,
                azim_num, dist_search, hori_acc, alg_c, geom_c,
                <const float*> vs.data, num_vert_simp,
                <const int32_t*> ti.data, num_tri_simp,
                elev_ang_low_lim, <const uint8_t*> mk.data, hori_fill,
                ray_org_elev, layout)
    if rc != 0:
        _raise_native()
    return hori_buffer, _azimuth_axis(azim_num)


def horizon_gridded_quantised(
        np.ndarray[np.float32_t, ndim = 1] vert_grid,
        int dem_dim_0, int dem_dim_1,
        np.ndarray[np.float32_t, ndim = 3] vec_norm,
        np.ndarray[np.float32_t, ndim = 3] vec_north,
        int offset_0, int offset_1,
        float dist_search,
        int azim_num=360,
        float hori_acc=0.25,
        str geom_type="grid",
        np.ndarray[np.float32_t, ndim = 1]
        vert_simp=np.array([0.0, 0.0, 0.0, 0.0], dtype=np.float32),
        int num_vert_simp=1,
        np.ndarray[np.int32_t, ndim = 1]
        tri_ind_simp=np.array([0, 0, 0, 0], dtype=np.int32),
        int num_tri_simp=1,
        float elev_ang_low_lim = -15.0,
        np.ndarray[np.uint8_t, ndim = 2] mask=None,
        float hori_fill=0.0,
        float ray_org_elev=0.01):
    """Additive (not in the reference): ``horizon_gridded`` with ``ray_algorithm="guess_constant"`` and a
    lossless 16-bit output.  Every result but the first azimuth's is an entry of the elevation table
    (``horizon_comp.cpp:490-494``), so the call returns

    ``(idx, first, table, azim)``: uint16 (y, x, azim_num) table indices (``0xFFFF`` = "take ``first``":
    azimuth 0 and masked cells), float32 (y, x) first-azimuth horizon (or ``hori_fill``), float32 (elev_num,)
    elevation table, float32 (azim_num,) azimuths.  ``dequantise(idx, first, table)`` reproduces the array of
    ``horizon_gridded`` bit for bit; the 2 GB array of a 1201 x 1201 x 360 run shrinks to 1 GB on the bus,
    in memory and on disk (``examples/horizon/gridded_curved_DEM.py:113-125``)."""
    if len(vert_grid) < (dem_dim_0 * dem_dim_1 * 3):
        raise ValueError("inconsistency between input arguments vert_grid, "
                         "dem_dim_0 and dem_dim_1")
    if ((offset_0 + vec_norm.shape[0] > dem_dim_0)
            or (offset_1 + vec_norm.shape[1] > dem_dim_1)):
        raise ValueError("inconsistency between input arguments dem_dim_0, "
                         "dem_dim_1, offset_0, offset_1 and vec_norm")
    if ((vec_norm.shape[0] != vec_north.shape[0]) or (vec_norm.shape[1] != vec_north.shape[1])
            or (vec_norm.shape[2] != vec_north.shape[2])):
        raise ValueError("dimension (lengths) of vec_norm and/or vec_north "
                         "is/are erroneous")
    if geom_type not in _GEOM_TYPES:
        raise ValueError("invalid input argument for geom_type")
    if len(vert_simp) < (num_vert_simp * 3):
        raise ValueError("inconsistency between input arguments vert_simp "
                         "and num_vert_simp")
    if len(tri_ind_simp) < (num_tri_simp * 3):
        raise ValueError("inconsistency between input arguments tri_ind_simp "
                         "and num_tri_simp")
    if tri_ind_simp.max() > (num_vert_simp - 1) or tri_ind_simp.min() < 0:
        raise ValueError("triangle indices of simplified outer domain exceed "
                         "number of vertices")
    if hori_acc > 10.0:
        raise ValueError("limit of hori_acc (10 degree) is exceeded")
    if mask is None:
        mask = np.ones((vec_norm.shape[0], vec_norm.shape[1]), dtype=np.uint8)
    if (mask.shape[0] != vec_norm.shape[0]) or (mask.shape[1] != vec_norm.shape[1]):
        raise ValueError("shape of mask is inconsistent with other input")
    if ray_org_elev < 0.005:
        raise TypeError("minimal allowed value for 'ray_org_elev' is 0.005 m")
    if (dem_dim_0 > 32767) or (dem_dim_1 > 32767):
        raise ValueError("maximal allowed input length for dem_dim_0 and "
                         "dem_dim_1 is 32'767")
    cdef int elev_num
    with nogil:
        elev_num = hzb_horizon_tables(azim_num, dist_search, hori_acc, elev_ang_low_lim, 0, NULL, NULL, NULL, NULL, NULL)
    if elev_num < 2 or elev_num > 65534:
        raise ValueError("the elevation table of these parameters has %d entries; the 16-bit output holds 2 ... 65534" % elev_num)
    cdef np.ndarray[np.float32_t, ndim = 1, mode = "c"] table = np.empty(elev_num, dtype=np.float32)
    cdef np.ndarray[np.float32_t, ndim = 1, mode = "c"] tsin = np.empty(elev_num, dtype=np.float32)
    cdef np.ndarray[np.float32_t, ndim = 1, mode = "c"] tcos = np.empty(elev_num, dtype=np.float32)
    with nogil:
        hzb_horizon_tables(azim_num, dist_search, hori_acc, elev_ang_low_lim, elev_num, <float*> table.data,
                           <float*> tsin.data, <float*> tcos.data, NULL, NULL)
    cdef np.ndarray[np.float32_t, ndim = 1, mode = "c"] vg = np.ascontiguousarray(vert_grid)
    cdef np.ndarray[np.float32_t, ndim = 3, mode = "c"] vn = np.ascontiguousarray(vec_norm)
    cdef np.ndarray[np.float32_t, ndim = 3, mode = "c"] vno = np.ascontiguousarray(vec_north)
    cdef np.ndarray[np.float32_t, ndim = 1, mode = "c"] vs = np.ascontiguousarray(vert_simp)
    cdef np.ndarray[np.int32_t, ndim = 1, mode = "c"] ti = np.ascontiguousarray(tri_ind_simp)
    cdef np.ndarray[np.uint8_t, ndim = 2, mode = "c"] mk = np.ascontiguousarray(mask)
    cdef bytes geom_b = geom_type.encode("utf-8")
    cdef const char* geom_c = geom_b
    cdef int ny = vn.shape[0], nx = vn.shape[1]
    cdef np.ndarray[np.uint16_t, ndim = 3, mode = "c"] idx = np.empty((ny, nx, azim_num), dtype=np.uint16)
    cdef np.ndarray[np.float32_t, ndim = 2, mode = "c"] first = np.empty((ny, nx), dtype=np.float32)
    cdef int rc = 0
    if ny > 0 and nx > 0:
        with nogil:
            rc = hzb_horizon_gridded_quantised(
                <const float*> vg.data, dem_dim_0, dem_dim_1,
                <const float*> vn.data, <const float*> vno.data,
                offset_0, offset_1, <unsigned short*> idx.data, <float*> first.data, ny, nx,
                azim_num, dist_search, hori_acc, geom_c,
                <const float*> vs.data, num_vert_simp,
                <const int32_t*> ti.data, num_tri_simp,
                elev_ang_low_lim, <const uint8_t*> mk.data, hori_fill, ray_org_elev)
    if rc != 0:
        _raise_native()
    return idx, first, table, _azimuth_axis(azim_num)


def dequantise(idx, first, table):
    """float32 (y, x, azim_num) horizon from the output of ``horizon_gridded_quantised``: bit for bit the
    array ``horizon_gridded`` returns."""
    idx = np.asarray(idx)
    take_first = idx == 0xFFFF
    out = np.asarray(table, dtype=np.float32)[np.where(take_first, 0, idx)]
    return np.where(take_first, np.asarray(first, dtype=np.float32)[..., None], out).astype(np.float32, copy=False)


def horizon_locations(
        np.ndarray[np.float32_t, ndim = 1] vert_grid,
        int dem_dim_0, int dem_dim_1,
        np.ndarray[np.float32_t, ndim = 2] coords,
        np.ndarray[np.float32_t, ndim = 2] vec_norm,
        np.ndarray[np.float32_t, ndim = 2] vec_north,
        float dist_search,
        int azim_num=360,
        float hori_acc=0.25,
        str ray_algorithm="binary_search",
        str geom_type="grid",
        float elev_ang_low_lim = -89.98,
        np.ndarray[np.float32_t, ndim = 1] ray_org_elev \
        = np.array([0.01], dtype=np.float32),
        bint hori_dist_out=False):
    """Horizon (and optionally distance to the horizon) at arbitrary locations.

    Arguments, units and defaults are those of
    ``horayzon.horizon.horizon_locations`` (``horizon.pyx:218-279``).  Locations
    whose normal line does not meet the DEM surface within 100 km keep NaN.

    Returns ``(hori_buffer, azim)`` or, with ``hori_dist_out=True``,
    ``(hori_buffer, hori_dist_buffer, azim)``.
    """
    # argument checks, in the reference's order and wording (horizon.pyx:282-313)
    if len(vert_grid) < (dem_dim_0 * dem_dim_1 * 3):
        raise ValueError("inconsistency between input arguments vert_grid, "
                         "dem_dim_0 and dem_dim_1")
    if ((coords.ndim != 2) or (coords.shape[0] != vec_norm.shape[0])
            or (coords.shape[1] !=3)):
        raise ValueError("'number of dimensions and/or dimension "
                         + "length(s) of 'coords' incorrect")
    if ((vec_norm.ndim != 2) or (vec_north.ndim != 2)
            or (vec_norm.shape[0] != vec_north.shape[0])
            or (vec_norm.shape[1] != vec_north.shape[1])):
        raise ValueError("dimension (lengths) of vec_norm and/or vec_north "
                         "is/are erroneous")
    if ray_algorithm not in _ALGORITHMS:
        raise ValueError("invalid input argument for ray_algorithm")
    if geom_type not in _GEOM_TYPES:
        raise ValueError("invalid input argument for geom_type")
    if hori_acc > 10.0:
        raise ValueError("limit of hori_acc (10 degree) is exceeded")
    if (len(ray_org_elev) != 1) and (len(ray_org_elev) != coords.shape[0]):
        raise ValueError("length of array 'ray_org_elev' must be either "
                         + "one or correspond to the number of locations")
    if ray_org_elev.min() < 0.005:
        raise TypeError("minimal allowed value for 'ray_org_elev' is 0.005 m")
    if hori_dist_out and (ray_algorithm == "guess_constant"):
        raise TypeError("horizon detection algorithm 'guess_constant' not "
                        + "implemented for horizon distance computation")
    if (dem_dim_0 > 32767) or (dem_dim_1 > 32767):
        raise ValueError("maximal allowed input length for dem_dim_0 and "
                         "dem_dim_1 is 32'767")

    if len(ray_org_elev) != coords.shape[0]:
        ray_org_elev = np.repeat(ray_org_elev, coords.shape[0])

    cdef np.ndarray[np.float32_t, ndim = 1, mode = "c"] vg = np.ascontiguousarray(vert_grid)
    cdef np.ndarray[np.float32_t, ndim = 2, mode = "c"] co = np.ascontiguousarray(coords)
    cdef np.ndarray[np.float32_t, ndim = 2, mode = "c"] vn = np.ascontiguousarray(vec_norm)
    cdef np.ndarray[np.float32_t, ndim = 2, mode = "c"] vno = np.ascontiguousarray(vec_north)
    cdef np.ndarray[np.float32_t, ndim = 1, mode = "c"] roe = np.ascontiguousarray(ray_org_elev)
    cdef bytes alg_b = ray_algorithm.encode("utf-8")
    cdef bytes geom_b = geom_type.encode("utf-8")
    cdef const char* alg_c = alg_b
    cdef const char* geom_c = geom_b
    cdef int num_loc = vn.shape[0]
    cdef int dist_flag = 1 if hori_dist_out else 0

    cdef np.ndarray[np.float32_t, ndim = 2, mode = "c"] hori_buffer = \
        np.empty((num_loc, azim_num), dtype=np.float32)
    hori_buffer.fill(np.nan)
    cdef np.ndarray[np.float32_t, ndim = 2, mode = "c"] hori_dist_buffer = \
        np.empty((num_loc if hori_dist_out else 1, azim_num), dtype=np.float32)
    hori_dist_buffer.fill(np.nan)  # 'dummy length' 1 without distance output (horizon.pyx:336-344)

    cdef int rc = 0
    if num_loc > 0:
        with nogil:
            rc = hzb_horizon_locations(
                <const float*> vg.data, dem_dim_0, dem_dim_1,
                <const float*> co.data, <const float*> vn.data,
                <const float*> vno.data, <float*> hori_buffer.data,
                <float*> hori_dist_buffer.data, num_loc, azim_num,
                dist_search, hori_acc, alg_c, geom_c, elev_ang_low_lim,
                <const float*> roe.data, dist_flag)
    if rc != 0:
        _raise_native()
    if hori_dist_out:
        return hori_buffer, hori_dist_buffer, _azimuth_axis(azim_num)
    return hori_buffer, _azimuth_axis(azim_num)
