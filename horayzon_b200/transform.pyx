# cython: language_level=3
"""Coordinate transformations -- drop-in for ``horayzon.transform`` (scope row "next 3").

Same functions, argument checks and messages as the reference
(``horayzon/transform.pyx:15-57, 108-149, 194-228, 266-303, 349-387, 437-530``);
the element loops (``_lonlat2ecef_1d`` ``:60-103``, ``_ecef2enu_1d`` ``:152-189``,
``_ecef2enu_vector_1d`` ``:231-261``, ``_wgs2swiss_1d`` ``:306-344``, ``_swiss2wgs_1d``
``:390-432``) and ``rotation_matrix_glob2loc`` (``:490-530``) run on the GPU through
``libhorayzon_b200.so``.  No CPU fallback.  ``TransformerEcef2enu`` only stores five
scalars (``:437-485``) and stays on the host, like in the reference.
"""
cimport numpy as np
import numpy as np

np.import_array()

cdef extern from "horayzon_b200.h":
    const char* hzb_last_error()
    int hzb_lonlat2ecef(const double* lon, const double* lat, const float* h, long long n, const char* ellps,
                        double* x, double* y, double* z) nogil
    int hzb_ecef2enu(const double* x, const double* y, const double* z, long long n, double x_or, double y_or,
                     double z_or, double lon_or, double lat_or, float* xe, float* ye, float* ze) nogil
    int hzb_ecef2enu_vector(const float* v, long long n, double lon_or, double lat_or, float* o) nogil
    int hzb_wgs2swiss(const double* lon, const double* lat, const float* h, long long n, double* e, double* nn,
                      float* hc) nogil
    int hzb_swiss2wgs(const double* e, const double* nn, const float* hc, long long n, double* lon, double* lat,
                      float* h) nogil
    int hzb_rotation_matrix_glob2loc(const float* north, const float* norm, int ny, int nx, float* out) nogil


def _raise_native():
    raise RuntimeError("horayzon_b200: " + hzb_last_error().decode("utf-8", "replace"))


def lonlat2ecef(lon, lat, h, ellps):
    """Geodetic longitude/latitude [degree, float64] and elevation [m, float32] to
    ECEF coordinates [m, float64] (transform.pyx:15-57)."""
    if (lon.shape != lat.shape) or (lat.shape != h.shape):
        raise ValueError("Inconsistent shapes / number of dimensions of "
                         + "input arrays")
    if ((lon.dtype != "float64") or (lat.dtype != "float64")
            or (h.dtype != "float32")):
        raise ValueError("Input array(s) has/have incorrect data type(s)")
    if ellps not in ("sphere", "GRS80", "WGS84"):
        raise ValueError("Unknown value for 'ellps'")
    shp = lon.shape
    cdef np.ndarray[np.float64_t, ndim = 1, mode = "c"] a = np.ascontiguousarray(lon.ravel())
    cdef np.ndarray[np.float64_t, ndim = 1, mode = "c"] b = np.ascontiguousarray(lat.ravel())
    cdef np.ndarray[np.float32_t, ndim = 1, mode = "c"] c = np.ascontiguousarray(h.ravel())
    cdef long long n = a.shape[0]
    cdef np.ndarray[np.float64_t, ndim = 1, mode = "c"] x = np.empty(n, dtype=np.float64)
    cdef np.ndarray[np.float64_t, ndim = 1, mode = "c"] y = np.empty(n, dtype=np.float64)
    cdef np.ndarray[np.float64_t, ndim = 1, mode = "c"] z = np.empty(n, dtype=np.float64)
    cdef bytes el = ellps.encode("ascii")
    cdef const char* elp = el
    cdef int rc
    with nogil:
        rc = hzb_lonlat2ecef(<const double*> a.data, <const double*> b.data, <const float*> c.data, n, elp,
                             <double*> x.data, <double*> y.data, <double*> z.data)
    if rc != 0:
        _raise_native()
    return x.reshape(shp), y.reshape(shp), z.reshape(shp)


def ecef2enu(x_ecef, y_ecef, z_ecef, trans_ecef2enu):
    """ECEF [m, float64] to local tangent plane (ENU) coordinates [m, float32]
    (transform.pyx:108-149)."""
    if (x_ecef.shape != y_ecef.shape) or (y_ecef.shape != z_ecef.shape):
        raise ValueError("Inconsistent shapes / number of dimensions of "
                         + "input arrays")
    if ((x_ecef.dtype != "float64") or (y_ecef.dtype != "float64")
            or (z_ecef.dtype != "float64")):
        raise ValueError("Input array(s) has/have incorrect data type(s)")
    if not isinstance(trans_ecef2enu, TransformerEcef2enu):
        raise ValueError("Last input argument must be instance of class "
                         + "'TransformerEcef2enu'")
    shp = x_ecef.shape
    cdef np.ndarray[np.float64_t, ndim = 1, mode = "c"] a = np.ascontiguousarray(x_ecef.ravel())
    cdef np.ndarray[np.float64_t, ndim = 1, mode = "c"] b = np.ascontiguousarray(y_ecef.ravel())
    cdef np.ndarray[np.float64_t, ndim = 1, mode = "c"] c = np.ascontiguousarray(z_ecef.ravel())
    cdef long long n = a.shape[0]
    cdef np.ndarray[np.float32_t, ndim = 1, mode = "c"] x = np.empty(n, dtype=np.float32)
    cdef np.ndarray[np.float32_t, ndim = 1, mode = "c"] y = np.empty(n, dtype=np.float32)
    cdef np.ndarray[np.float32_t, ndim = 1, mode = "c"] z = np.empty(n, dtype=np.float32)
    cdef double x0 = trans_ecef2enu.x_ecef_or, y0 = trans_ecef2enu.y_ecef_or, z0 = trans_ecef2enu.z_ecef_or
    cdef double lo = trans_ecef2enu.lon_or, la = trans_ecef2enu.lat_or
    cdef int rc
    with nogil:
        rc = hzb_ecef2enu(<const double*> a.data, <const double*> b.data, <const double*> c.data, n, x0, y0, z0,
                          lo, la, <float*> x.data, <float*> y.data, <float*> z.data)
    if rc != 0:
        _raise_native()
    return x.reshape(shp), y.reshape(shp), z.reshape(shp)


def ecef2enu_vector(vec_ecef, trans_ecef2enu):
    """Vectors (components in the last dimension, float32) from ECEF to ENU
    (transform.pyx:194-228)."""
    if (vec_ecef.ndim < 2) or (vec_ecef.shape[vec_ecef.ndim - 1] != 3):
        raise ValueError("Incorrect shape / number of dimensions of input "
                         + "array")
    if vec_ecef.dtype != "float32":
        raise ValueError("Input array has incorrect data type")
    if not isinstance(trans_ecef2enu, TransformerEcef2enu):
        raise ValueError("Last input argument must be instance of class "
                         + "'TransformerEcef2enu'")
    shp = vec_ecef.shape[:(vec_ecef.ndim - 1)]
    cdef np.ndarray[np.float32_t, ndim = 2, mode = "c"] v = np.ascontiguousarray(vec_ecef.reshape(-1, 3))
    cdef long long n = v.shape[0]
    cdef np.ndarray[np.float32_t, ndim = 2, mode = "c"] o = np.empty((n, 3), dtype=np.float32)
    cdef double lo = trans_ecef2enu.lon_or, la = trans_ecef2enu.lat_or
    cdef int rc
    with nogil:
        rc = hzb_ecef2enu_vector(<const float*> v.data, n, lo, la, <float*> o.data)
    if rc != 0:
        _raise_native()
    return o.reshape(shp + (3,))


def wgs2swiss(lon, lat, h_wgs):
    """WGS84 lon/lat [degree] and ellipsoidal height to LV95 (transform.pyx:266-303)."""
    if (lon.shape != lat.shape) or (lat.shape != h_wgs.shape):
        raise ValueError("Inconsistent shapes / number of dimensions of "
                         + "input arrays")
    if ((lon.dtype != "float64") or (lat.dtype != "float64")
            or (h_wgs.dtype != "float32")):
        raise ValueError("Input array(s) has/have incorrect data type(s)")
    shp = lon.shape
    cdef np.ndarray[np.float64_t, ndim = 1, mode = "c"] a = np.ascontiguousarray(lon.ravel())
    cdef np.ndarray[np.float64_t, ndim = 1, mode = "c"] b = np.ascontiguousarray(lat.ravel())
    cdef np.ndarray[np.float32_t, ndim = 1, mode = "c"] c = np.ascontiguousarray(h_wgs.ravel())
    cdef long long n = a.shape[0]
    cdef np.ndarray[np.float64_t, ndim = 1, mode = "c"] e = np.empty(n, dtype=np.float64)
    cdef np.ndarray[np.float64_t, ndim = 1, mode = "c"] nn = np.empty(n, dtype=np.float64)
    cdef np.ndarray[np.float32_t, ndim = 1, mode = "c"] hc = np.empty(n, dtype=np.float32)
    cdef int rc
    with nogil:
        rc = hzb_wgs2swiss(<const double*> a.data, <const double*> b.data, <const float*> c.data, n,
                           <double*> e.data, <double*> nn.data, <float*> hc.data)
    if rc != 0:
        _raise_native()
    return e.reshape(shp), nn.reshape(shp), hc.reshape(shp)


def swiss2wgs(e, n, h_ch):
    """LV95 to WGS84 lon/lat [degree] and ellipsoidal height (transform.pyx:349-387)."""
    if (e.shape != n.shape) or (n.shape != h_ch.shape):
        raise ValueError("Inconsistent shapes / number of dimensions of "
                         + "input arrays")
    if ((e.dtype != "float64") or (n.dtype != "float64")
            or (h_ch.dtype != "float32")):
        raise ValueError("Input array(s) has/have incorrect data type(s)")
    shp = e.shape
    cdef np.ndarray[np.float64_t, ndim = 1, mode = "c"] a = np.ascontiguousarray(e.ravel())
    cdef np.ndarray[np.float64_t, ndim = 1, mode = "c"] b = np.ascontiguousarray(n.ravel())
    cdef np.ndarray[np.float32_t, ndim = 1, mode = "c"] c = np.ascontiguousarray(h_ch.ravel())
    cdef long long cnt = a.shape[0]
    cdef np.ndarray[np.float64_t, ndim = 1, mode = "c"] lon = np.empty(cnt, dtype=np.float64)
    cdef np.ndarray[np.float64_t, ndim = 1, mode = "c"] lat = np.empty(cnt, dtype=np.float64)
    cdef np.ndarray[np.float32_t, ndim = 1, mode = "c"] hw = np.empty(cnt, dtype=np.float32)
    cdef int rc
    with nogil:
        rc = hzb_swiss2wgs(<const double*> a.data, <const double*> b.data, <const float*> c.data, cnt,
                           <double*> lon.data, <double*> lat.data, <float*> hw.data)
    if rc != 0:
        _raise_native()
    return lon.reshape(shp), lat.reshape(shp), hw.reshape(shp)


class TransformerEcef2enu:
    """Attributes of the ECEF -> ENU transformation; the ENU origin lies on the
    surface of the sphere / ellipsoid (transform.pyx:437-485)."""

    def __init__(self, lon_or, lat_or, ellps):
        if (lon_or < -180.0) or (lon_or > 180.0):
            raise ValueError("Value for 'lon_or' is outside of valid range")
        if (lat_or < -90.0) or (lat_or > 90.0):
            raise ValueError("Value for 'lat_or' is outside of valid range")
        self.lon_or = lon_or
        self.lat_or = lat_or
        if ellps == "sphere":
            r = 6370997.0  # earth radius [m]
            self.x_ecef_or = r * np.cos(np.deg2rad(self.lat_or)) \
                * np.cos(np.deg2rad(self.lon_or))
            self.y_ecef_or = r * np.cos(np.deg2rad(self.lat_or)) \
                * np.sin(np.deg2rad(self.lon_or))
            self.z_ecef_or = r * np.sin(np.deg2rad(self.lat_or))
        elif ellps in ("GRS80", "WGS84"):
            a = 6378137.0  # equatorial radius (semi-major axis) [m]
            if ellps == "GRS80":
                f = (1.0 / 298.257222101)  # flattening [-]
            else:  # WGS84
                f = (1.0 / 298.257223563)  # flattening [-]
            b = a * (1.0 - f)  # polar radius (semi-minor axis) [m]
            e_2 = 1.0 - (b ** 2 / a ** 2)  # squared num. eccentricity [-]
            n = a / np.sqrt(1.0 - e_2 * np.sin(np.deg2rad(self.lat_or)) ** 2)
            self.x_ecef_or = n * np.cos(np.deg2rad(self.lat_or)) \
                * np.cos(np.deg2rad(self.lon_or))
            self.y_ecef_or = n * np.cos(np.deg2rad(self.lat_or)) \
                * np.sin(np.deg2rad(self.lon_or))
            self.z_ecef_or = (b ** 2 / a ** 2 * n) \
                * np.sin(np.deg2rad(self.lat_or))
        else:
            raise ValueError("Unknown value for 'ellps'")


def rotation_matrix_glob2loc(vec_north_enu, vec_norm_enu):
    """Matrices that rotate vectors from global to local ENU coordinates: rows
    east = north x norm, north, norm; one NaN cell on each side so that the
    shape matches the DEM domain used for slope (transform.pyx:490-530)."""
    if vec_north_enu.shape != vec_norm_enu.shape:
        raise ValueError("Inconsistent shapes / number of dimensions of "
                         + "input arrays")
    cdef np.ndarray[np.float32_t, ndim = 3, mode = "c"] a = np.ascontiguousarray(vec_north_enu, dtype=np.float32)
    cdef np.ndarray[np.float32_t, ndim = 3, mode = "c"] b = np.ascontiguousarray(vec_norm_enu, dtype=np.float32)
    if a.shape[2] != 3:
        raise ValueError("Inconsistent shapes / number of dimensions of "
                         + "input arrays")
    cdef int ny = a.shape[0], nx = a.shape[1]
    cdef np.ndarray[np.float32_t, ndim = 4, mode = "c"] out = np.empty((ny + 2, nx + 2, 3, 3), dtype=np.float32)
    cdef int rc
    with nogil:
        rc = hzb_rotation_matrix_glob2loc(<const float*> a.data, <const float*> b.data, ny, nx, <float*> out.data)
    if rc != 0:
        _raise_native()
    return out
