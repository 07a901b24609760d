/* horayzon_b200.h -- C ABI of libhorayzon_b200.so
 *
 * The drop-in boundary for HORAYZON's horizon / shadow / sky-view hot path.
 * Every entry point replaces one native interface of the reference (cited as
 * file:line relative to the reference's horayzon/ directory) and keeps its
 * argument order, units and meaning.  Plain pointers and sizes only.
 *
 * Conventions
 *   - All functions returning int return 0 on success; on failure the message
 *     is available from hzb_last_error() (thread-local).  The reference's
 *     native functions return void and print Embree errors with printf
 *     (horizon_comp.cpp:74-76); status codes replace that.
 *   - "host tier" functions take HOST pointers, exactly like the reference, and
 *     perform H2D copy, on-device BVH build, kernels and D2H copy inside the
 *     call.  Buffers are borrowed for the duration of the call only.
 *   - "resident tier" functions (suffix _dev, and hzb_scene_*) are additive:
 *     they take DEVICE pointers and a cudaStream_t (passed as void*) so that a
 *     caller can keep DEM, BVH and outputs in HBM, shard rows over GPUs and
 *     overlap work.  They have no counterpart in the reference.
 *   - Units at the boundary are the reference's (horizon_comp.cpp:667-670):
 *     dist_search [km], hori_acc / elev_ang_low_lim [degree], hori_fill
 *     [radian], ray_org_elev [m]; outputs in radian.
 *   - There is no CPU fallback: without a CUDA device every compute entry
 *     point fails with a non-zero status.
 */
#ifndef HORAYZON_B200_H
#define HORAYZON_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ status */
const char* hzb_last_error(void);
int hzb_device_count(void);       /* number of visible CUDA devices (0 if none) */
const char* hzb_version(void);

/* Work counters and timings of the most recent host-tier call on this thread.
 * rays = casts performed, the reference's own work counter "Number of rays
 * shot" (horizon_comp.cpp:739,799-810).  Seconds are wall-clock around
 * device-synchronised phases, mirroring the reference's prints "BVH build
 * time" (:225-227), "Ray tracing time" (:805), "Total run time" (:818). */
typedef struct hzb_stats {
    unsigned long long rays;        /* ray casts                               */
    unsigned long long node_visits; /* BVH nodes fetched (per ray; per packet for warp_node_visits) */
    unsigned long long prim_tests;  /* leaf primitives (quads / TIN triangles) tested */
    unsigned long long units;       /* unmasked cells x azimuths (or cells for shadow) */
    unsigned long long warp_node_visits; /* packet traversal: nodes fetched per warp */
    double t_h2d, t_build, t_trace, t_d2h, t_total;
    unsigned long long num_prims;   /* BVH primitives (grid quads + TIN triangles) */
    unsigned long long num_nodes;   /* wide-BVH nodes */
    unsigned long long bvh_bytes;   /* bytes of the traversal structure in HBM */
    unsigned long long fallback_packets; /* packets re-decided by the binary-BVH walker after a full traversal stack */
    unsigned long long segment_tasks;    /* azimuth segments >= 1 of split cells the horizon kernel ran (tail of a launch) */
    unsigned long long segment_redos;    /* ... of which the fix-up pass recomputed (start index not the chain's) */
} hzb_stats;
int hzb_get_stats(hzb_stats* out);

/* Additive, pure host code (works without a device): the trig tables of one call, built exactly like
 * horizon_comp.cpp:711-731 (azimuth sin/cos; elevation angle/sin/cos anchored at 89.98 degree, step
 * hori_acc/5).  Returns elev_num (or -1); the elevation arrays are filled when cap >= elev_num, the
 * azimuth arrays when non-NULL.  The table entries are the alphabet of the horizon output. */
int hzb_horizon_tables(int azim_num, float dist_search, float hori_acc, float elev_ang_low_lim, int cap,
                       float* elev_ang, float* elev_sin, float* elev_cos, float* azim_sin, float* azim_cos);

/* Additive: pooled page-locked host blocks for large arrays the WRAPPER allocates and returns
 * (the reference's wrapper allocates its outputs with np.empty, horizon.pyx:170-173).  Host-tier
 * calls recognise page-locked buffers and move them by plain DMA instead of staging.  NULL when
 * no device / no memory: the caller falls back to ordinary memory. */
void* hzb_host_alloc(size_t bytes);
void hzb_host_free(void* p);
/* Additive: release the idle pooled memory of this process (device blocks the host tier keeps
 * between calls, at most 6 GB; the one idle page-locked output block, at most 4 GB). */
void hzb_trim(void);
/* Additive, pure host code (no device needed; for test suites): the work-queue layout the horizon kernel would use for a
 * launch (band along the DEM's edge first, interior in row order, the last tiles as azimuth segments; DESIGN.md
 * section 5).  lo / hi: bounds of the DEM vertices (3 floats each); resident_ctas: CTAs of the persistent grid.
 * out[0..6] = segments per split cell (1: no split), first / end local block row and tile-column margin of the interior,
 * split tiles, queue entries, tiles of the launch. */
int hzb_plan_queue(int dem_dim_0, int dem_dim_1, const float* lo, const float* hi, int offset_0, int offset_1, int dim_in_0,
                   int dim_in_1, int row_begin, int row_end, int shard_rank, int shard_count, int azim_num, float dist_search,
                   float hori_acc, float elev_ang_low_lim, const char* ray_algorithm, int resident_ctas, long long* out);
/* Test-only switches: second implementations ("horizon_kernel" 1, "shadow_kernel" 1/2: reference-shaped
 * per-lane kernels on the binary BVH / nearest-first order), tuning knobs ("wrefill", "wwait"),
 * "no_overlap", "stack_limit" (forces the full-stack fallback), "tail_segments" / "tail_tiles" /
 * "tail_band" (azimuth segments in the tail of a horizon launch: off, forced, everywhere),
 * "ctas_per_sm", "reset".  Production code never calls it; no environment variable selects a kernel. */
int hzb_debug_option(const char* name, int value);

/* --------------------------------------------------------------- host tier */

/* Replaces horizon_gridded_comp (horizon_comp.h:8-20, horizon_comp.cpp:629-822).
 * hori_buffer: float32 [dim_in_0][dim_in_1][azim_num], radians. */
int hzb_horizon_gridded(const float* vert_grid, int dem_dim_0, int dem_dim_1,
                        const float* vec_norm, const float* vec_north,
                        int offset_0, int offset_1,
                        float* hori_buffer,
                        int dim_in_0, int dim_in_1,
                        int azim_num, float dist_search,
                        float hori_acc, const char* ray_algorithm, const char* geom_type,
                        const float* vert_simp, int num_vert_simp,
                        const int32_t* tri_ind_simp, int num_tri_simp,
                        float elev_ang_low_lim,
                        const uint8_t* mask, float hori_fill,
                        float ray_org_elev);

/* Additive (scope row 8f-4): the same call with the output in azimuth-first order
 * [azim_num][dim_in_0][dim_in_1] when azim_first != 0 -- the layout the reference's examples
 * transpose to before writing NetCDF (examples/horizon/gridded_curved_DEM.py:113-125), so a
 * 2 GB np.moveaxis + copy on the host disappears.  azim_first == 0 is hzb_horizon_gridded. */
int hzb_horizon_gridded_layout(const float* vert_grid, int dem_dim_0, int dem_dim_1,
                               const float* vec_norm, const float* vec_north,
                               int offset_0, int offset_1, float* hori_buffer,
                               int dim_in_0, int dim_in_1, int azim_num,
                               float dist_search, float hori_acc, const char* ray_algorithm,
                               const char* geom_type, const float* vert_simp, int num_vert_simp,
                               const int32_t* tri_ind_simp, int num_tri_simp,
                               float elev_ang_low_lim, const uint8_t* mask, float hori_fill,
                               float ray_org_elev, int azim_first);

/* Additive: horizon + sky view factor in one call -- horizon_gridded_comp followed by
 * _sky_view_factor_cy (topo_param.pyx:412-460) on the device-resident horizon, as
 * examples/horizon/gridded_curved_DEM.py:104-144 calls them back to back.  vec_tilt:
 * [dim_in_0][dim_in_1][3] (local frame); svf_buffer: [dim_in_0][dim_in_1].  Same values as
 * hzb_horizon_gridded + hzb_sky_view_factor; the horizon array is not uploaded a second time. */
int hzb_horizon_gridded_svf(const float* vert_grid, int dem_dim_0, int dem_dim_1,
                            const float* vec_norm, const float* vec_north,
                            int offset_0, int offset_1, float* hori_buffer,
                            int dim_in_0, int dim_in_1, int azim_num,
                            float dist_search, float hori_acc, const char* ray_algorithm,
                            const char* geom_type, const float* vert_simp, int num_vert_simp,
                            const int32_t* tri_ind_simp, int num_tri_simp,
                            float elev_ang_low_lim, const uint8_t* mask, float hori_fill,
                            float ray_org_elev, const float* vec_tilt, float* svf_buffer);

/* Additive (scope row 8f-4): quantised horizon output for ray_algorithm "guess_constant" (the default).  Every
 * result but the first azimuth's is an entry of the elevation table (horizon_comp.cpp:490-494), so the call
 * returns 16-bit table indices idx_buffer [dim_in_0][dim_in_1][azim_num] (0xFFFF: "take first_buffer": azimuth 0
 * and masked cells) and first_buffer [dim_in_0][dim_in_1] (the first azimuth's un-quantised midpoint, :428, or
 * hori_fill).  Lossless: with elev_ang from hzb_horizon_tables, elev_ang[idx] is bit for bit what
 * hzb_horizon_gridded stores.  Half the bytes to copy, keep and write to disk
 * (examples/horizon/gridded_curved_DEM.py:113-125 is where the 2-400 GB float array goes to NetCDF). */
int hzb_horizon_gridded_quantised(const float* vert_grid, int dem_dim_0, int dem_dim_1,
                                  const float* vec_norm, const float* vec_north,
                                  int offset_0, int offset_1, uint16_t* idx_buffer, float* first_buffer,
                                  int dim_in_0, int dim_in_1, int azim_num,
                                  float dist_search, float hori_acc, const char* geom_type,
                                  const float* vert_simp, int num_vert_simp,
                                  const int32_t* tri_ind_simp, int num_tri_simp,
                                  float elev_ang_low_lim, const uint8_t* mask, float hori_fill,
                                  float ray_org_elev);

/* Additive: multi-GPU twin of hzb_horizon_gridded (same leading arguments, horizon_comp.h:8-20) -- one
 * process, one host thread per GPU.  The reference parallelises over rows of the inner domain with TBB
 * (horizon_comp.cpp:739-744); here the 4-row blocks are dealt out to the GPUs in turn, every GPU builds the
 * BVH of the replicated DEM and computes its blocks.  n_devices <= 0: all visible devices; n_shards <= 0: one
 * shard per device (more shards than devices are dealt round-robin).  device_gather == 0: each GPU copies its
 * blocks to their places in the host array over its own PCIe link; != 0: the packed shards are joined by ONE
 * in-place ncclAllGather over NVLink (single process, ncclCommInitAll; NCCL is dlopen'ed), put in domain
 * order on GPU 0 and returned from there.  vec_tilt / svf_buffer: optional fused sky view factor
 * ([dim_in_0][dim_in_1][3] / [dim_in_0][dim_in_1]; both NULL: none).  Same values as the single-GPU call. */
int hzb_horizon_gridded_multi(const float* vert_grid, int dem_dim_0, int dem_dim_1,
                              const float* vec_norm, const float* vec_north,
                              int offset_0, int offset_1, float* hori_buffer,
                              int dim_in_0, int dim_in_1, int azim_num,
                              float dist_search, float hori_acc, const char* ray_algorithm,
                              const char* geom_type, const float* vert_simp, int num_vert_simp,
                              const int32_t* tri_ind_simp, int num_tri_simp,
                              float elev_ang_low_lim, const uint8_t* mask, float hori_fill,
                              float ray_org_elev, const float* vec_tilt, float* svf_buffer,
                              int n_devices, int n_shards, int device_gather);
/* Additive: select the CUDA device of the calling host thread for the host-tier calls that follow
 * (they run on the current device), e.g. one hzb_terrain per GPU driven from one thread each. */
int hzb_set_device(int device);

/* Replaces horizon_locations_comp (horizon_comp.h:23-34, horizon_comp.cpp:828-1094).
 * hori_buffer / hori_dist_buffer: float32 [num_loc][azim_num]; locations whose
 * normal line misses the surface are left untouched (the wrapper pre-fills NaN). */
int hzb_horizon_locations(const float* vert_grid, int dem_dim_0, int dem_dim_1,
                          const float* coords,
                          const float* vec_norm, const float* vec_north,
                          float* hori_buffer,
                          float* hori_dist_buffer,
                          int num_loc,
                          int azim_num, float dist_search,
                          float hori_acc, const char* ray_algorithm, const char* geom_type,
                          float elev_ang_low_lim,
                          const float* ray_org_elev,
                          int hori_dist_out);

/* Replaces class shapes::CppTerrain (shadow_comp.h:3-40, shadow_comp.cpp:304-605).
 * Unlike the reference, initialise COPIES every input to the device, so the
 * caller may free its arrays afterwards (shadow_comp.cpp:332-346 keeps raw
 * pointers). */
typedef struct hzb_terrain hzb_terrain;
hzb_terrain* hzb_terrain_create(void);                       /* CppTerrain()  :304-308 */
void hzb_terrain_destroy(hzb_terrain* t);                    /* ~CppTerrain() :310-316 */
int hzb_terrain_initialise(hzb_terrain* t,                   /* ::initialise  :318-380 */
                           const float* vert_grid,
                           int dem_dim_0, int dem_dim_1,
                           int offset_0, int offset_1,
                           const float* vec_tilt,
                           const float* vec_norm,
                           int dim_in_0, int dim_in_1,
                           const float* surf_enl_fac,
                           const float* elevation,
                           const uint8_t* mask,
                           const char* geom_type,
                           float sw_dir_cor_fill,
                           float ang_max,
                           int refrac_cor);
/* ::shadow :386-491 -- codes 0 lit, 1 self-shaded, 2 terrain-shaded, 3 masked */
int hzb_terrain_shadow(hzb_terrain* t, const float* sun_position, uint8_t* shadow_buffer);
/* ::sw_dir_cor :495-605 */
int hzb_terrain_sw_dir_cor(hzb_terrain* t, const float* sun_position, float* sw_dir_cor_buffer);
/* additive: n_sun positions in one call; outputs [n_sun][dim_in_0][dim_in_1] */
int hzb_terrain_shadow_batch(hzb_terrain* t, const float* sun_positions, int n_sun,
                             uint8_t* shadow_buffer);
int hzb_terrain_sw_dir_cor_batch(hzb_terrain* t, const float* sun_positions, int n_sun,
                                 float* sw_dir_cor_buffer);

/* Replace _sky_view_factor_cy / _visible_sky_fraction_cy / _topographic_openness_cy
 * (topo_param.pyx:412-460, 499-543, 577-603).  hori: [ny][nx][K], vec_tilt:
 * [ny][nx][3] (local frame), out: [ny][nx]. */
int hzb_sky_view_factor(const float* azim, const float* hori, const float* vec_tilt,
                        int ny, int nx, int K, float* out);
int hzb_visible_sky_fraction(const float* azim, const float* hori, const float* vec_tilt,
                             int ny, int nx, int K, float* out);
int hzb_topographic_openness(const float* azim, const float* hori,
                             int ny, int nx, int K, float* out);

/* "Next" row 8f-1: tilted-surface normals on the device.  Replace _slope_plane_meth_cy /
 * _slope_vector_meth_cy (topo_param.pyx:84-225, 284-372).  x, y, z: [ny][nx]; rot_mat:
 * [ny][nx][3][3] or NULL (identity / none); out: [ny][nx][3], border cells NaN. */
int hzb_slope_plane_meth(const float* x, const float* y, const float* z, const float* rot_mat,
                         int ny, int nx, int output_rot, float* out);
int hzb_slope_vector_meth(const float* x, const float* y, const float* z, const float* rot_mat,
                          int ny, int nx, int output_rot, float* out);

/* "Next" row 8f-3: coordinate preparation on the device.  Replace the loops of
 * _lonlat2ecef_1d (transform.pyx:60-103), _ecef2enu_1d (:152-189), _ecef2enu_vector_1d
 * (:231-261), _wgs2swiss_1d (:306-344), _swiss2wgs_1d (:390-432),
 * rotation_matrix_glob2loc (:490-530), _surf_norm_1d (direction.pyx:48-70) and
 * _north_dir_1d (direction.pyx:125-178).  n = number of points; vectors are [n][3];
 * ellps is "sphere", "GRS80" or "WGS84"; *_or are the TransformerEcef2enu attributes
 * (transform.pyx:437-485).  rot_mat: [(ny+2)][(nx+2)][3][3] with a NaN rim. */
int hzb_lonlat2ecef(const double* lon, const double* lat, const float* h, long long n, const char* ellps,
                    double* x_ecef, double* y_ecef, double* z_ecef);
int hzb_ecef2enu(const double* x_ecef, const double* y_ecef, const double* z_ecef, long long n,
                 double x_ecef_or, double y_ecef_or, double z_ecef_or, double lon_or, double lat_or,
                 float* x_enu, float* y_enu, float* z_enu);
int hzb_ecef2enu_vector(const float* vec_ecef, long long n, double lon_or, double lat_or, float* vec_enu);
int hzb_surf_norm(const double* lon, const double* lat, long long n, float* vec_norm_ecef);
int hzb_north_dir(const double* x_ecef, const double* y_ecef, const double* z_ecef, const float* vec_norm_ecef,
                  long long n, const char* ellps, float* vec_north_ecef);
int hzb_wgs2swiss(const double* lon, const double* lat, const float* h_wgs, long long n,
                  double* e, double* nn, float* h_ch);
int hzb_swiss2wgs(const double* e, const double* nn, const float* h_ch, long long n,
                  double* lon, double* lat, float* h_wgs);
int hzb_rotation_matrix_glob2loc(const float* vec_north_enu, const float* vec_norm_enu, int ny, int nx,
                                 float* rot_mat);

/* ----------------------------------------------------------- resident tier */

/* A scene = DEM vertices (+ optional TIN) and their BVH, resident on `device`.
 * Replaces initializeDevice + initializeScene (horizon_comp.cpp:79-86, 101-231):
 * Morton codes + radix sort + LBVH + wide-BVH collapse, all on the device.
 * vert_grid etc. are HOST pointers (copied). */
typedef struct hzb_scene hzb_scene;
hzb_scene* hzb_scene_create(const float* vert_grid, int dem_dim_0, int dem_dim_1,
                            const float* vert_simp, int num_vert_simp,
                            const int32_t* tri_ind_simp, int num_tri_simp,
                            int device);
void hzb_scene_destroy(hzb_scene* s);
int hzb_scene_stats(const hzb_scene* s, hzb_stats* out);

/* Horizon for rows [row_begin, row_end) of the inner domain.  d_* are DEVICE
 * pointers to the FULL inner-domain arrays ([dim_in_0][dim_in_1][...]); only
 * the selected rows are read / written.  stream is a cudaStream_t (NULL = the
 * legacy default stream).  Asynchronous: nothing on the launch path synchronises
 * once the tables of a parameter set (azim_num, dist_search, hori_acc,
 * elev_ang_low_lim) have been uploaded by their first use.  Launches on DIFFERENT
 * streams against one scene may overlap (every launch has its own work-queue
 * counter, the cached tables are never overwritten); calls on one scene must come
 * from one host thread at a time.  Counters accumulate into the scene and are read
 * (after synchronising) with hzb_scene_stats. */
int hzb_horizon_gridded_dev(hzb_scene* s,
                            const float* d_vec_norm, const float* d_vec_north,
                            const uint8_t* d_mask,
                            int offset_0, int offset_1,
                            int dim_in_0, int dim_in_1,
                            int row_begin, int row_end,
                            int azim_num, float dist_search, float hori_acc,
                            const char* ray_algorithm, float elev_ang_low_lim,
                            float hori_fill, float ray_org_elev,
                            float* d_hori_buffer, void* stream);

/* As hzb_horizon_gridded_dev, with d_hori_buffer in azimuth-first order
 * [azim_num][dim_in_0][dim_in_1] when azim_first != 0 (row sharding works unchanged). */
int hzb_horizon_gridded_dev_layout(hzb_scene* s,
                                   const float* d_vec_norm, const float* d_vec_north, const uint8_t* d_mask,
                                   int offset_0, int offset_1, int dim_in_0, int dim_in_1,
                                   int row_begin, int row_end, int azim_num,
                                   float dist_search, float hori_acc, const char* ray_algorithm,
                                   float elev_ang_low_lim, float hori_fill, float ray_org_elev,
                                   float* d_hori_buffer, int azim_first, void* stream);

/* Resident-tier form of hzb_horizon_gridded_quantised (rows [row_begin, row_end), device buffers). */
int hzb_horizon_gridded_dev_quantised(hzb_scene* s,
                                      const float* d_vec_norm, const float* d_vec_north, const uint8_t* d_mask,
                                      int offset_0, int offset_1, int dim_in_0, int dim_in_1,
                                      int row_begin, int row_end, int azim_num,
                                      float dist_search, float hori_acc, float elev_ang_low_lim,
                                      float hori_fill, float ray_org_elev,
                                      uint16_t* d_idx_buffer, float* d_first_buffer, void* stream);

/* Additive (multi-GPU): block-interleaved sharding.  Shard `shard_rank` of `shard_count` computes the 4-row
 * blocks b of the inner domain with b % shard_count == shard_rank -- the reference's row partition
 * (horizon_comp.cpp:739-744) dealt out in 4-row blocks, so that every GPU gets the same mix of cheap rim
 * rows and expensive centre rows.  packed == 0: results land in their places of the full
 * [dim_in_0][dim_in_1][azim_num] array; packed != 0: the shard's blocks are stored back to back from
 * d_hori_buffer on (hzb_shard_rows rows): the contiguous send buffer of one all-gather, after which
 * block j of shard r is block j * shard_count + r of the result. */
int hzb_horizon_gridded_dev_sharded(hzb_scene* s,
                                    const float* d_vec_norm, const float* d_vec_north, const uint8_t* d_mask,
                                    int offset_0, int offset_1, int dim_in_0, int dim_in_1, int azim_num,
                                    float dist_search, float hori_acc, const char* ray_algorithm,
                                    float elev_ang_low_lim, float hori_fill, float ray_org_elev,
                                    float* d_hori_buffer, int shard_rank, int shard_count, int packed,
                                    void* stream);
int hzb_shard_rows(int dim_in_0, int shard_rank, int shard_count);

/* Device-pointer variants of the azimuthal integrals (same layouts). */
int hzb_sky_view_factor_dev(const float* d_azim, const float* d_hori, const float* d_vec_tilt,
                            long long num_cells, int K, float* d_out, void* stream);
int hzb_visible_sky_fraction_dev(const float* d_azim, const float* d_hori,
                                 const float* d_vec_tilt, long long num_cells, int K,
                                 float* d_out, void* stream);
int hzb_topographic_openness_dev(const float* d_azim, const float* d_hori,
                                 long long num_cells, int K, float* d_out, void* stream);

/* Additive (no counterpart in the reference): the whole preparation chain of
 * examples/horizon/gridded_curved_DEM.py:60-90 fused into one kernel with inputs and
 * outputs in HBM -- lon [nx], lat [ny] (degree, double), elevation [ny][nx] (float) ->
 * vert_grid [ny][nx][3] (ENU, the wire format of hzb_scene_create / hzb_horizon_gridded
 * without the padding) and, for the inner domain, vec_norm / vec_north [dim_in_0][dim_in_1][3]
 * in ENU (both may be NULL).  Same device functions as the step-by-step entry points. */
int hzb_prep_enu_dev(const double* d_lon, const double* d_lat, const float* d_elev, int ny, int nx,
                     const char* ellps, double x_ecef_or, double y_ecef_or, double z_ecef_or,
                     double lon_or, double lat_or, int offset_0, int offset_1, int dim_in_0, int dim_in_1,
                     float* d_vert_grid, float* d_vec_norm, float* d_vec_north, void* stream);

/* Shadow / sw_dir_cor with a DEVICE output buffer (terrain already resident). */
int hzb_terrain_shadow_dev(hzb_terrain* t, const float* sun_position_host,
                           uint8_t* d_shadow_buffer, void* stream);
int hzb_terrain_sw_dir_cor_dev(hzb_terrain* t, const float* sun_position_host,
                               float* d_sw_dir_cor_buffer, void* stream);
int hzb_terrain_stats(const hzb_terrain* t, hzb_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* HORAYZON_B200_H */
