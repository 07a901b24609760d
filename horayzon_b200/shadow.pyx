# cython: language_level=3
"""Shadow mask and shortwave correction factor -- drop-in for ``horayzon.shadow``.

``Terrain`` keeps the Python interface of the reference class
(``horayzon/shadow.pyx:17-200``) and forwards to the opaque-handle family
``hzb_terrain_*`` of ``libhorayzon_b200.so`` (``include/horayzon_b200.h``), which
replaces ``shapes::CppTerrain`` (``horayzon/shadow_comp.cpp:304-605``).  Unlike
the reference, ``initialise`` copies its inputs to the GPU, so the arrays may
be released afterwards.  No CPU fallback.
"""
cimport numpy as np
import numpy as np
from libc.stdint cimport uint8_t

np.import_array()

cdef extern from "horayzon_b200.h":
    ctypedef struct hzb_terrain:
        pass
    const char* hzb_last_error()
    hzb_terrain* hzb_terrain_create()
    void hzb_terrain_destroy(hzb_terrain* t)
    int hzb_terrain_initialise(
        hzb_terrain* t, const float* vert_grid, int dem_dim_0, int dem_dim_1,
        int offset_0, int offset_1, const float* vec_tilt,
        const float* vec_norm, int dim_in_0, int dim_in_1,
        const float* surf_enl_fac, const float* elevation,
        const uint8_t* mask, const char* geom_type, float sw_dir_cor_fill,
        float ang_max, int refrac_cor) nogil
    int hzb_terrain_shadow(hzb_terrain* t, const float* sun_position,
                           uint8_t* shadow_buffer) nogil
    int hzb_terrain_sw_dir_cor(hzb_terrain* t, const float* sun_position,
                               float* sw_dir_cor_buffer) nogil
    int hzb_terrain_shadow_batch(hzb_terrain* t, const float* sun_positions,
                                 int n_sun, uint8_t* shadow_buffer) nogil
    int hzb_terrain_sw_dir_cor_batch(hzb_terrain* t, const float* sun_positions,
                                     int n_sun, float* sw_dir_cor_buffer) nogil


def _raise_native():
    raise RuntimeError("horayzon_b200: " + hzb_last_error().decode("utf-8", "replace"))


cdef class Terrain:
    """Terrain resident on the GPU; one occlusion ray per cell and sun position."""

    cdef hzb_terrain* handle
    cdef int dim_in_0, dim_in_1
    cdef bint initialised

    def __cinit__(self):
        self.handle = hzb_terrain_create()
        self.initialised = False
        self.dim_in_0 = 0
        self.dim_in_1 = 0

    def __dealloc__(self):
        if self.handle != NULL:
            hzb_terrain_destroy(self.handle)
            self.handle = NULL

    def initialise(self, np.ndarray[np.float32_t, ndim = 1] vert_grid,
                   int dem_dim_0, int dem_dim_1,
                   int offset_0, int offset_1,
                   np.ndarray[np.float32_t, ndim = 3] vec_tilt,
                   np.ndarray[np.float32_t, ndim = 3] vec_norm,
                   np.ndarray[np.float32_t, ndim = 2] surf_enl_fac,
                   np.ndarray[np.float32_t, ndim = 2] elevation,
                   np.ndarray[np.uint8_t, ndim = 2] mask,
                   str geom_type="grid",
                   float sw_dir_cor_fill=np.nan,
                   float ang_max=89.0,
                   bint refrac_cor=False):
        """Upload DEM and per-cell auxiliaries and build the BVH.

        Arguments as ``horayzon.shadow.Terrain.initialise`` (``shadow.pyx:27-84``):
        ``vec_tilt`` / ``vec_norm`` unit vectors in global ENU (y, x, 3),
        ``surf_enl_fac`` and ``elevation`` (y, x), ``mask`` uint8 (0 ignored,
        1 considered), ``ang_max`` [degree], ``refrac_cor`` atmospheric
        refraction on/off.
        """
        # argument checks, in the reference's order and wording (shadow.pyx:87-133)
        if len(vert_grid) < (dem_dim_0 * dem_dim_1 * 3):
            raise ValueError("inconsistency between input arguments "
                             + "'vert_grid', 'dem_dim_0' and 'dem_dim_1'")
        if ((offset_0 + vec_tilt.shape[0] > dem_dim_0)
                or (offset_1 + vec_tilt.shape[1] > dem_dim_1)):
            raise ValueError("inconsistency between input arguments "
                             + "'dem_dim_0', 'dem_dim_1', 'offset_0', "
                             + "'offset_1' and 'vec_norm'")
        if ((vec_tilt.ndim != 3) or (vec_norm.ndim != 3)
                or (vec_tilt.shape[2] != 3)
                or (vec_tilt.shape[0] != vec_norm.shape[0])
                or (vec_tilt.shape[1] != vec_norm.shape[1])
                or (vec_tilt.shape[2] != vec_norm.shape[2])):
            raise ValueError("Inconsistent/incorrect shape of 'vec_tilt' "
                             + "and/or 'vec_norm'")
        if ((surf_enl_fac.ndim != 2) or (elevation.ndim != 2)
                or (mask.ndim != 2)
                or (surf_enl_fac.shape[0] != vec_tilt.shape[0])
                or (surf_enl_fac.shape[1] != vec_tilt.shape[1])
                or (elevation.shape[0] != vec_tilt.shape[0])
                or (elevation.shape[1] != vec_tilt.shape[1])
                or (mask.shape[0] != vec_tilt.shape[0])
                or (mask.shape[1] != vec_tilt.shape[1])):
            raise ValueError("Inconsistent/incorrect shape of 'surf_enl_fac', "
                             + " 'elevation' and/or 'mask'")
        if ((not vert_grid.flags["C_CONTIGUOUS"])
                or (not vec_tilt.flags["C_CONTIGUOUS"])
                or (not vec_norm.flags["C_CONTIGUOUS"])
                or (not surf_enl_fac.flags["C_CONTIGUOUS"])
                or (not elevation.flags["C_CONTIGUOUS"])
                or (not mask.flags["C_CONTIGUOUS"])):
            raise ValueError("not all input arrays are C-contiguous")
        if ((np.abs((vec_tilt ** 2).sum(axis=2) - 1.0).max() > 1.0e-5)
                or (np.abs((vec_norm ** 2).sum(axis=2) - 1.0).max() > 1.0e-5)):
            raise ValueError("Vectors in 'vec_tilt' and/or 'vec_norm' are "
                             + "not normalised")
        if geom_type not in ("triangle", "quad", "grid"):
            raise ValueError("invalid input argument for geom_type")
        if mask.dtype != "uint8":
            raise TypeError("data type of mask must be 'uint8'")
        if (ang_max < 85.0) or (ang_max > 89.99):
            raise TypeError("'ang_max' must be in the range [85.0, 89.99]")
        if (dem_dim_0 > 32767) or (dem_dim_1 > 32767):
            raise ValueError("maximal allowed input length for dem_dim_0 and "
                             "dem_dim_1 is 32'767")

        cdef bytes geom_b = geom_type.encode("utf-8")
        cdef const char* geom_c = geom_b
        cdef int ny = vec_tilt.shape[0], nx = vec_tilt.shape[1]
        cdef int refrac = 1 if refrac_cor else 0
        cdef int rc
        with nogil:
            rc = hzb_terrain_initialise(
                self.handle, <const float*> vert_grid.data, dem_dim_0, dem_dim_1,
                offset_0, offset_1, <const float*> vec_tilt.data,
                <const float*> vec_norm.data, ny, nx,
                <const float*> surf_enl_fac.data, <const float*> elevation.data,
                <const uint8_t*> mask.data, geom_c, sw_dir_cor_fill, ang_max, refrac)
        if rc != 0:
            _raise_native()
        self.dim_in_0 = ny
        self.dim_in_1 = nx
        self.initialised = True

    cdef _check_ready(self, shape):
        if not self.initialised:
            raise RuntimeError("horayzon_b200: Terrain.initialise() has not been called")
        if shape[0] != self.dim_in_0 or shape[1] != self.dim_in_1:
            raise ValueError("output buffer shape does not match the inner domain")

    def shadow(self, np.ndarray[np.float32_t, ndim = 1] sun_position,
               np.ndarray[np.uint8_t, ndim = 2] shadow_buffer):
        """Shadow mask for one sun position (global ENU [m]), written in place:
        0 illuminated, 1 self-shaded, 2 terrain-shaded, 3 masked
        (``shadow.pyx:149-170``, ``shadow_comp.cpp:386-491``)."""
        if (sun_position.ndim != 1) or (sun_position.size != 3):
            raise ValueError("array 'sun_position' has incorrect shape")
        if not shadow_buffer.flags["C_CONTIGUOUS"]:
            raise ValueError("array 'shadow_buffer' is not C-contiguous")
        self._check_ready((shadow_buffer.shape[0], shadow_buffer.shape[1]))
        cdef np.ndarray[np.float32_t, ndim = 1, mode = "c"] sp = np.ascontiguousarray(sun_position)
        cdef int rc
        with nogil:
            rc = hzb_terrain_shadow(self.handle, <const float*> sp.data,
                                    <uint8_t*> shadow_buffer.data)
        if rc != 0:
            _raise_native()

    def sw_dir_cor(self, np.ndarray[np.float32_t, ndim = 1] sun_position,
                   np.ndarray[np.float32_t, ndim = 2] sw_dir_cor_buffer):
        """Correction factor for direct downward shortwave radiation for one sun
        position, written in place (``shadow.pyx:172-200``,
        ``shadow_comp.cpp:495-605``; Mueller & Scherer 2005)."""
        if (sun_position.ndim != 1) or (sun_position.size != 3):
            raise ValueError("array 'sun_position' has incorrect shape")
        if not sw_dir_cor_buffer.flags["C_CONTIGUOUS"]:
            raise ValueError("array 'sw_dir_cor_buffer' is not C-contiguous")
        self._check_ready((sw_dir_cor_buffer.shape[0], sw_dir_cor_buffer.shape[1]))
        cdef np.ndarray[np.float32_t, ndim = 1, mode = "c"] sp = np.ascontiguousarray(sun_position)
        cdef int rc
        with nogil:
            rc = hzb_terrain_sw_dir_cor(self.handle, <const float*> sp.data,
                                        <float*> sw_dir_cor_buffer.data)
        if rc != 0:
            _raise_native()

    # ---- additive: many sun positions per call (one D2H copy for all) ----
    def shadow_batch(self, np.ndarray[np.float32_t, ndim = 2] sun_positions):
        """Shadow masks for ``sun_positions`` (n, 3) -> uint8 (n, y, x)."""
        if sun_positions.shape[1] != 3:
            raise ValueError("array 'sun_positions' has incorrect shape")
        if not self.initialised:
            raise RuntimeError("horayzon_b200: Terrain.initialise() has not been called")
        cdef np.ndarray[np.float32_t, ndim = 2, mode = "c"] sp = np.ascontiguousarray(sun_positions)
        cdef int n = sp.shape[0]
        cdef np.ndarray[np.uint8_t, ndim = 3, mode = "c"] out = \
            np.empty((n, self.dim_in_0, self.dim_in_1), dtype=np.uint8)
        cdef int rc
        with nogil:
            rc = hzb_terrain_shadow_batch(self.handle, <const float*> sp.data, n, <uint8_t*> out.data)
        if rc != 0:
            _raise_native()
        return out

    def sw_dir_cor_batch(self, np.ndarray[np.float32_t, ndim = 2] sun_positions):
        """Correction factors for ``sun_positions`` (n, 3) -> float32 (n, y, x)."""
        if sun_positions.shape[1] != 3:
            raise ValueError("array 'sun_positions' has incorrect shape")
        if not self.initialised:
            raise RuntimeError("horayzon_b200: Terrain.initialise() has not been called")
        cdef np.ndarray[np.float32_t, ndim = 2, mode = "c"] sp = np.ascontiguousarray(sun_positions)
        cdef int n = sp.shape[0]
        cdef np.ndarray[np.float32_t, ndim = 3, mode = "c"] out = \
            np.empty((n, self.dim_in_0, self.dim_in_1), dtype=np.float32)
        cdef int rc
        with nogil:
            rc = hzb_terrain_sw_dir_cor_batch(self.handle, <const float*> sp.data, n, <float*> out.data)
        if rc != 0:
            _raise_native()
        return out
