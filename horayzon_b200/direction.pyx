# cython: language_level=3
"""Unit vectors of the local frame -- drop-in for ``horayzon.direction`` (scope row "next 3").

``surf_norm`` and ``north_dir`` keep the reference's signatures and checks
(``horayzon/direction.pyx:15-45, 75-122``); the element loops (``_surf_norm_1d``
``:48-70``, ``_north_dir_1d`` ``:125-178``) run on the GPU through
``libhorayzon_b200.so``.  No CPU fallback.
"""
cimport numpy as np
import numpy as np
from math import prod

np.import_array()

cdef extern from "horayzon_b200.h":
    const char* hzb_last_error()
    int hzb_surf_norm(const double* lon, const double* lat, long long n, float* out) nogil
    int hzb_north_dir(const double* x, const double* y, const double* z, const float* vec_norm, long long n,
                      const char* ellps, float* out) nogil


def _raise_native():
    raise RuntimeError("horayzon_b200: " + hzb_last_error().decode("utf-8", "replace"))


def surf_norm(lon, lat):
    """Surface normal unit vectors of the ellipsoid in ECEF coordinates
    (direction.pyx:15-45)."""
    if lon.shape != lat.shape:
        raise ValueError("Inconsistent shapes / number of dimensions of "
                         + "input arrays")
    if ((lon.dtype != "float64") or (lat.dtype != "float64")):
        raise ValueError("Input array(s) has/have incorrect data type(s)")
    shp = lon.shape
    cdef np.ndarray[np.float64_t, ndim = 1, mode = "c"] a = np.ascontiguousarray(lon.ravel())
    cdef np.ndarray[np.float64_t, ndim = 1, mode = "c"] b = np.ascontiguousarray(lat.ravel())
    cdef long long n = a.shape[0]
    cdef np.ndarray[np.float32_t, ndim = 2, mode = "c"] o = np.empty((n, 3), dtype=np.float32)
    cdef int rc
    with nogil:
        rc = hzb_surf_norm(<const double*> a.data, <const double*> b.data, n, <float*> o.data)
    if rc != 0:
        _raise_native()
    return o.reshape(shp + (3,))


def north_dir(x_ecef, y_ecef, z_ecef, vec_norm_ecef, ellps):
    """Unit vectors pointing towards North, perpendicular to the surface normals,
    in ECEF coordinates (direction.pyx:75-122)."""
    if (x_ecef.shape != y_ecef.shape) or (y_ecef.shape != z_ecef.shape) \
            or (z_ecef.shape != vec_norm_ecef
                                .shape[:(vec_norm_ecef.ndim - 1)]):
        raise ValueError("Inconsistent shapes / number of dimensions of "
                         + "input arrays")
    if ((x_ecef.dtype != "float64") or (y_ecef.dtype != "float64")
            or (z_ecef.dtype != "float64")
            or (vec_norm_ecef.dtype != "float32")):
        raise ValueError("Input array(s) has/have incorrect data type(s)")
    if ellps not in ("sphere", "GRS80", "WGS84"):
        raise ValueError("Unknown value for 'ellps'")
    shp = x_ecef.shape
    cdef np.ndarray[np.float64_t, ndim = 1, mode = "c"] a = np.ascontiguousarray(x_ecef.ravel())
    cdef np.ndarray[np.float64_t, ndim = 1, mode = "c"] b = np.ascontiguousarray(y_ecef.ravel())
    cdef np.ndarray[np.float64_t, ndim = 1, mode = "c"] c = np.ascontiguousarray(z_ecef.ravel())
    cdef np.ndarray[np.float32_t, ndim = 2, mode = "c"] v = np.ascontiguousarray(vec_norm_ecef.reshape(prod(shp), 3))
    cdef long long n = a.shape[0]
    cdef np.ndarray[np.float32_t, ndim = 2, mode = "c"] o = np.empty((n, 3), dtype=np.float32)
    cdef bytes el = ellps.encode("ascii")
    cdef const char* elp = el
    cdef int rc
    with nogil:
        rc = hzb_north_dir(<const double*> a.data, <const double*> b.data, <const double*> c.data,
                           <const float*> v.data, n, elp, <float*> o.data)
    if rc != 0:
        _raise_native()
    return o.reshape(shp + (3,))
