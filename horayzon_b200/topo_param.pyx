# cython: language_level=3
"""Azimuthal integrals over the horizon -- drop-in for the SVF part of
``horayzon.topo_param``.

``sky_view_factor``, ``visible_sky_fraction`` and ``topographic_openness`` keep
the reference's signatures and checks (``horayzon/topo_param.pyx:377-409,
465-496, 548-574``) and run on the GPU through ``libhorayzon_b200.so``
(``hzb_sky_view_factor`` etc., replacing the single-threaded loops at
``topo_param.pyx:412-460, 499-543, 577-603``).  ``slope_plane_meth`` and
``slope_vector_meth`` (input preparation for SVF / ``Terrain``; scope row "next 1",
``topo_param.pyx:16-372``) run on the GPU as well.  No CPU fallback.
"""
cimport numpy as np
import numpy as np

np.import_array()

cdef extern from "horayzon_b200.h":
    const char* hzb_last_error()
    int hzb_sky_view_factor(const float* azim, const float* hori, const float* vec_tilt,
                            int ny, int nx, int K, float* out) nogil
    int hzb_visible_sky_fraction(const float* azim, const float* hori, const float* vec_tilt,
                                 int ny, int nx, int K, float* out) nogil
    int hzb_topographic_openness(const float* azim, const float* hori,
                                 int ny, int nx, int K, float* out) nogil
    int hzb_slope_plane_meth(const float* x, const float* y, const float* z, const float* rot_mat,
                             int ny, int nx, int output_rot, float* out) nogil
    int hzb_slope_vector_meth(const float* x, const float* y, const float* z, const float* rot_mat,
                              int ny, int nx, int output_rot, float* out) nogil


def _raise_native():
    raise RuntimeError("horayzon_b200: " + hzb_last_error().decode("utf-8", "replace"))


def _check_inputs(azim, hori, vec_tilt):
    # topo_param.pyx:400-405 / :487-492 / :567-570
    if vec_tilt is None:
        if len(azim) != hori.shape[2]:
            raise ValueError("Inconsistent/incorrect shapes of input arrays")
        if (azim.dtype != "float32") or (hori.dtype != "float32"):
            raise ValueError("Input array(s) has/have incorrect data type(s)")
        return
    if (len(azim) != hori.shape[2]) or (hori.shape[:2] != vec_tilt.shape[:2])\
            or (vec_tilt.shape[2] != 3):
        raise ValueError("Inconsistent/incorrect shapes of input arrays")
    if ((azim.dtype != "float32") or (hori.dtype != "float32")
            or (vec_tilt.dtype != "float32")):
        raise ValueError("Input array(s) has/have incorrect data type(s)")


cdef _integral(int kind, azim, hori, vec_tilt):
    cdef np.ndarray[np.float32_t, ndim = 1, mode = "c"] a = np.ascontiguousarray(azim)
    cdef np.ndarray[np.float32_t, ndim = 3, mode = "c"] h = np.ascontiguousarray(hori)
    cdef np.ndarray[np.float32_t, ndim = 3, mode = "c"] t
    cdef int ny = h.shape[0], nx = h.shape[1], K = h.shape[2]
    cdef np.ndarray[np.float32_t, ndim = 2, mode = "c"] out = np.empty((ny, nx), dtype=np.float32)
    cdef int rc = 0
    cdef const float* tp = NULL
    if kind != 2:
        t = np.ascontiguousarray(vec_tilt)
        tp = <const float*> t.data
    if ny > 0 and nx > 0:
        with nogil:
            if kind == 0:
                rc = hzb_sky_view_factor(<const float*> a.data, <const float*> h.data, tp, ny, nx, K, <float*> out.data)
            elif kind == 1:
                rc = hzb_visible_sky_fraction(<const float*> a.data, <const float*> h.data, tp, ny, nx, K, <float*> out.data)
            else:
                rc = hzb_topographic_openness(<const float*> a.data, <const float*> h.data, ny, nx, K, <float*> out.data)
    if rc != 0:
        _raise_native()
    return out


def sky_view_factor(azim, hori, vec_tilt):
    """Sky view factor in the local horizontal frame: float32 (y, x).
    ``azim`` (K,) [rad], ``hori`` (y, x, K) [rad], ``vec_tilt`` (y, x, 3)."""
    _check_inputs(azim, hori, vec_tilt)
    return _integral(0, azim, hori, vec_tilt)


def visible_sky_fraction(azim, hori, vec_tilt):
    """Visible sky fraction (solid angle of visible sky): float32 (y, x)."""
    _check_inputs(azim, hori, vec_tilt)
    return _integral(1, azim, hori, vec_tilt)


def topographic_openness(azim, hori):
    """Positive topographic openness (Yokoyama et al. 2002) [rad]: float32 (y, x)."""
    _check_inputs(azim, hori, None)
    return _integral(2, azim, hori, None)


cdef _slope(int method, x, y, z, rot_mat, bint output_rot):
    cdef np.ndarray[np.float32_t, ndim = 2, mode = "c"] xa = np.ascontiguousarray(x)
    cdef np.ndarray[np.float32_t, ndim = 2, mode = "c"] ya = np.ascontiguousarray(y)
    cdef np.ndarray[np.float32_t, ndim = 2, mode = "c"] za = np.ascontiguousarray(z)
    cdef np.ndarray[np.float32_t, ndim = 4, mode = "c"] ra
    cdef int ny = xa.shape[0], nx = xa.shape[1]
    cdef np.ndarray[np.float32_t, ndim = 3, mode = "c"] out = np.empty((ny, nx, 3), dtype=np.float32)
    cdef const float* rp = NULL
    cdef int orot = 1 if output_rot else 0
    cdef int rc = 0
    if rot_mat is not None:
        ra = np.ascontiguousarray(rot_mat)
        rp = <const float*> ra.data
    if ny > 0 and nx > 0:
        with nogil:
            if method == 0:
                rc = hzb_slope_plane_meth(<const float*> xa.data, <const float*> ya.data, <const float*> za.data, rp, ny, nx, orot, <float*> out.data)
            else:
                rc = hzb_slope_vector_meth(<const float*> xa.data, <const float*> ya.data, <const float*> za.data, rp, ny, nx, orot, <float*> out.data)
    if rc != 0:
        _raise_native()
    return out


def _check_slope_inputs(x, y, z, rot_mat):
    # topo_param.pyx:60-72 / :261-275
    if (x.shape != y.shape) or (y.shape != z.shape):
        raise ValueError("Inconsistent shapes / number of dimensions of "
                         + "input arrays")
    if ((x.dtype != "float32") or (y.dtype != "float32")
            or (z.dtype != "float32")):
        raise ValueError("Input array(s) has/have incorrect data type(s)")
    if rot_mat is not None:
        if ((x.shape[0] != rot_mat.shape[0])
                or (x.shape[1] != rot_mat.shape[1])):
            raise ValueError("Inconsistent shapes / number of dimensions of "
                             + "input arrays")
        if rot_mat.dtype != "float32":
            raise ValueError("'rot mat' has incorrect data type")


def slope_plane_meth(x, y, z, rot_mat=None, output_rot=False):
    """Tilted-surface normals from a least-squares plane through the 3 x 3
    neighbourhood (``topo_param.pyx:16-225``).  ``x, y, z`` float32 (y, x);
    ``rot_mat`` optional float32 (y, x, 3, 3) rotation to a frame whose z-axis is
    local up (identity if omitted); ``output_rot`` keeps the result in that frame.
    Returns float32 (y, x, 3); border cells are NaN."""
    _check_slope_inputs(x, y, z, rot_mat)
    return _slope(0, x, y, z, rot_mat, output_rot)


def slope_vector_meth(x, y, z, rot_mat=None, output_rot=False):
    """Tilted-surface normals as the average of the four adjacent triangle normals
    (``topo_param.pyx:230-372``; Corripio 2003).  Arguments as ``slope_plane_meth``;
    ``rot_mat`` is only applied to the output when ``output_rot`` is true."""
    _check_slope_inputs(x, y, z, rot_mat)
    if output_rot and (rot_mat is None):
        raise ValueError("'rot_mat' must be provided for 'output_rot = True'")
    return _slope(1, x, y, z, rot_mat, output_rot)
