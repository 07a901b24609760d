// hzb_common.cuh -- shared declarations of libhorayzon_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/horayzon_b200.h"
#include "hzb_hd.cuh"
#include "hzb_queue.cuh"

namespace hzb {

// ------------------------------------------------------------------ errors
void set_error(const std::string& msg);
void set_last_stats(const hzb_stats& st);   // what hzb_get_stats returns on this thread
#define HZB_CUDA(call)                                                                     \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            ::hzb::set_error(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" +   \
                             __FILE__ + ":" + std::to_string(__LINE__) + ")");             \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)
#define HZB_TRY(expr)              \
    do {                           \
        if ((expr) != 0) return 1; \
    } while (0)

double now_s();

// --------------------------------------------------------------- BVH nodes
// Binary LBVH node (build-time structure; also the v1 traversal structure):
// both children's boxes live in the parent.  64 B = four 16-byte loads.
// child code: >= 0 internal node index; < 0 leaf holding primitive ~code.
struct __align__(16) Bvh2Node {
    float lo0[3], hi0[3], lo1[3], hi1[3];
    int c0, c1;
    int pad0, pad1;
};
static_assert(sizeof(Bvh2Node) == 64, "Bvh2Node must be 64 bytes");

// 4-wide BVH node with child boxes quantised to 16 bits on a scene-global grid
// (plane = qorg[a] + q * qstep[a]): no per-node header, so the lane that owns
// child k needs exactly one 16-byte load per node.  64 B = two 32-byte sectors.
//   qx/qy/qz : lo | hi << 16 (conservative: lo rounded down, hi rounded up)
//   ref      : 0xFFFFFFFF empty (box inverted); bit 31 set -> leaf holding
//              primitive (ref & 0x7FFFFFFF); else index of the child node.
struct __align__(16) WideChild { uint32_t qx, qy, qz, ref; };
struct __align__(64) Bvh4Node { WideChild c[4]; };
static_assert(sizeof(Bvh4Node) == 64, "Bvh4Node must be 64 bytes");
// WIDE_EMPTY / WIDE_LEAF: hzb_hd.cuh

// Device view of a scene: DEM vertices, optional TIN, BVH.
// Primitive p < num_quads is grid quad (i = p / (W-1), j = p % (W-1)), split
// along the diagonal (i,j+1)-(i+1,j) into triangles ((i,j),(i,j+1),(i+1,j))
// and ((i+1,j+1),(i+1,j),(i,j+1)) (reference horizon_comp.cpp:140-151,
// 163-171, 178-183).  Primitive p >= num_quads is TIN triangle p - num_quads.
struct SceneView {
    const float4* vert4;   // [H*W] (x, y, z, 0)
    const float4* tin4;    // [3*num_tin] (x, y, z, 0)
    const Bvh2Node* nodes2;
    const Bvh4Node* nodes4;
    float qorg[3], qstep[3];  // quantisation grid of nodes4
    const uint32_t* prim_ids;  // sorted (Morton) order -> primitive id
    int H, W;
    uint32_t num_quads, num_tin, num_prims;
    uint32_t num_nodes4;
};

struct Counters {  // device-side accumulators (one struct per scene / terrain)
    unsigned long long rays, node_visits, prim_tests, units, warp_node_visits;
    unsigned long long stack_overflow;     // binary-BVH walker ran out of its 96-entry stack (results invalid: reported as an error)
    unsigned long long fallback_packets;   // packets re-decided by the binary-BVH walker after a full shared-memory stack (informational)
    unsigned long long segment_tasks, segment_redos;   // azimuth segments >= 1 run as tasks of their own / recomputed by the fix-up pass
};

// Test-only switches (hzb_debug_option): second implementations and tuning knobs the parity tests
// and A/B runs select explicitly.  Production code never changes them; no environment variable does.
struct DebugOptions {
    int horizon_kernel = 0;   // 0 production packet kernel, 1 reference-shaped per-lane kernel on the binary BVH
    int shadow_kernel = 0;    // 0 production, 1 reference-shaped per-lane kernel, 2 production step with nearest-first order
    int wrefill = 24, wwait = 2;
    int no_overlap = 0;       // host tier: copy the horizon array after the kernel instead of while it runs
    int stack_limit = 1 << 20; // clamped to WQ_STACK_N; lowered by the tests to force the full-stack fallback
    int ctas_per_sm = 0;      // horizon kernel: resident CTAs per SM of the persistent grid (0: the compiled maximum)
    int tail_segments = -1;   // azimuth segments per cell in the tail of a launch: -1 automatic (4 where it pays), 1 off, 4 forced
    int tail_tiles = -1;      // tiles whose cells are split (-1: two per resident warp)
    int tail_band = -1;       // cells next to the DEM's edge that are never split (-1: from the scene's relief and the table's low limit)
};
DebugOptions& debug_options();

// ------------------------------------------------------------------- scene
struct Scene {
    int device = 0;
    int H = 0, W = 0;
    uint32_t num_quads = 0, num_tin = 0, num_prims = 0, num_nodes4 = 0;
    float4* d_vert4 = nullptr;
    float4* d_tin4 = nullptr;
    Bvh2Node* d_nodes2 = nullptr;
    Bvh4Node* d_nodes4 = nullptr;
    float qorg[3] = {0, 0, 0}, qstep[3] = {1, 1, 1};
    uint32_t* d_prim_ids = nullptr;
    Counters* d_counters = nullptr;
    // Work-queue counters: a ring of HZB_TILE_SLOTS words, one per launch in flight, so that launches on
    // different streams against the same scene never share (or reset) each other's queue.
    unsigned int* d_tile_counter = nullptr;
    unsigned int tile_slot = 0;
    float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0}, pad = 0.f;
    double t_h2d = 0, t_build = 0;
    size_t bvh_bytes = 0;
    // Device trig / elevation tables, cached per parameter set (azim_num, hori_acc, low limit, dist): an entry is
    // written once and never overwritten, so kernels in flight on other streams keep valid tables and a repeated
    // call uploads nothing (no stream synchronisation on the launch path).
    struct TableEntry { int azim_num; float acc_deg, low_deg, dist_km; float* d; int elev_num; };
    std::vector<TableEntry> tables;
    SceneView view() const {
        SceneView v;
        v.vert4 = d_vert4; v.tin4 = d_tin4; v.nodes2 = d_nodes2; v.nodes4 = d_nodes4;
        for (int a = 0; a < 3; ++a) { v.qorg[a] = qorg[a]; v.qstep[a] = qstep[a]; }
        v.prim_ids = d_prim_ids; v.H = H; v.W = W; v.num_quads = num_quads; v.num_tin = num_tin;
        v.num_prims = num_prims; v.num_nodes4 = num_nodes4;
        return v;
    }
};

// bvh_build.cu / bvh_wide.cu
int build_wide_bvh(Scene& s, cudaStream_t st);
int scene_upload_and_build(Scene& s, const float* vert_grid, int H, int W, const float* vert_simp,
                           int num_vert_simp, const int32_t* tri_ind_simp, int num_tri_simp);
void scene_free(Scene& s);

// ------------------------------------------------------ horizon parameters
struct HorizonTables {  // host copies; built exactly like horizon_comp.cpp:711-731
    int azim_num = 0, elev_num = 0;
    float acc = 0, low = 0, up = 0, dist = 0;
    double step = 0;  // (double)acc / 5.0
    std::vector<float> azim_sin, azim_cos, elev_ang, elev_sin, elev_cos;
    void make(int azim_num, float dist_km, float acc_deg, float low_deg, bool fill = true);
};

struct HorizonParams {
    // tables (device)
    const float* azim_sin; const float* azim_cos;
    const float* elev_ang; const float* elev_sin; const float* elev_cos;
    int azim_num, elev_num;
    float acc, low, up, dist; double step;
    int algorithm;  // 0 discrete_sampling, 1 binary_search, 2 guess_constant
    // inner domain
    const float* vec_norm; const float* vec_north; const uint8_t* mask;
    int offset_0, offset_1, dim_in_0, dim_in_1, row_begin, row_end;
    float hori_fill, ray_org_elev;
    // Block sharding (multi-GPU, cost-balanced): of the 4-row blocks of [row_begin, row_end) this launch computes
    // those with (block % blk_stride) == blk_offset; (1, 0) = all of them.  packed != 0: the launch's blocks are
    // stored back to back ([local block][4][dim_in_1][azim]) from `hori` on -- a contiguous all-gather send buffer.
    int blk_stride, blk_offset, packed;
    float* hori;
    unsigned short* hori_q; float* hori_first;   // quantised output instead of `hori` (both or none): 16-bit table indices
                                                 // [cell][azimuth] + the first azimuth's float per cell (scope row 8f-4)
    long long stride_c, stride_k;   // element (cell c, azimuth k) lives at hori[c * stride_c + k * stride_k]: (K, 1) = the reference's
                                    // [y][x][azim] layout, (1, cells) = azimuth-first [azim][y][x] (scope row "next 4")
    unsigned int* row_done;  // optional [ceil(rows/4)]: +1 per finished cell slot of that row block, 32 per 8x4 tile (host overlaps D2H)
    volatile unsigned int* row_flags;  // optional, MAPPED HOST memory [ceil(rows/4)]: set to 1 by the lane that completes a row block
    unsigned int row_full;   // cell slots per row block (32 per tile)
    // Work queue (horizon.cu, "Queue order and azimuth segments"): tiles of the band -- block rows outside [q_by0, q_by1),
    // tile columns outside [q_bx, tiles_x - q_bx) -- come first, then the interior in row order; the cells of the last
    // q_tail interior tiles are split into seg_count azimuth segments (one queue entry per tile and segment).
    int seg_count;           // 1: no segments
    int q_by0, q_by1, q_bx;
    unsigned int q_tail;
    // derived on the host (plan_queue): tile grid of the launch, interior width, section sizes of the queue
    int q_tiles_x, q_tiles_y, q_wi;
    unsigned int q_nA1, q_nA2, q_nA3, q_nI, q_total;
    int q_gb_end, q_gb_tail, q_tx_tail;
    SegRecord* seg;          // [q_tail * 32][SEG_COUNT]: per split cell, what the fix-up pass needs to know about segments 1.. (records
                             // 0 .. SEG_COUNT-2) and the index azimuth 0's bisection ended with (guess of the last record)
};

int parse_algorithm(const char* s);  // -1 if unknown
int parse_geom_type(const char* s);  // -1 if unknown

// horizon.cu
int launch_horizon_gridded(Scene& s, const HorizonParams& p, cudaStream_t st);
struct LocationParams {
    const float* coords; const float* vec_norm; const float* vec_north; const float* ray_org_elev;
    float* hori; float* hori_dist; int num_loc; int hori_dist_out;
};
int launch_horizon_locations(Scene& s, const HorizonParams& p, const LocationParams& lp, cudaStream_t st);
int scene_tables(Scene& s, int azim_num, float dist_km, float acc_deg, float low_deg, HorizonParams& p, cudaStream_t st);
void plan_queue_host(const Scene& s, HorizonParams& p, int grid_ctas);   // queue layout of a launch (no device needed)
void seg_pool_trim();      // releases the idle memory of the segment-record pools (hzb_trim)
unsigned int* scene_tile_counter(Scene& s, cudaStream_t st);   // fresh (zeroed on `st`) work-queue counter for one launch
constexpr unsigned int HZB_TILE_SLOTS = 64;

// shadow.cu
struct TerrainParams {
    const float* vec_tilt; const float* vec_norm; const float* surf_enl_fac; const float* elevation;
    const uint8_t* mask;
    int offset_0, offset_1, dim_in_0, dim_in_1;
    float sw_dir_cor_fill, ang_max; int refrac_cor;
    float dot_prod_min;  // cosf(deg2rad(ang_max)) computed on the host
    float t_ref, p_ref, lapse, expo;
};
int launch_shadow(Scene& s, const TerrainParams& tp, const float* sun_xyz /*host*/, uint8_t* d_out, cudaStream_t st);
int launch_sw_dir_cor(Scene& s, const TerrainParams& tp, const float* sun_xyz /*host*/, float* d_out, cudaStream_t st);

// topo.cu
int launch_svf(int kind, const float* d_azim, const float* d_hori, const float* d_tilt, long long cells,
               int K, float* d_out, cudaStream_t st);

int launch_slope(int method, const float* d_x, const float* d_y, const float* d_z, const float* d_rot, int ny, int nx,
                 int output_rot, float* d_out, cudaStream_t st);

// transform.cu (scope row "next 3"): coordinate preparation; all pointers are DEVICE pointers
int launch_lonlat2ecef(int ellps, const double* lon, const double* lat, const float* h, long long n, double* x, double* y, double* z, cudaStream_t st);
int launch_ecef2enu(const double* x, const double* y, const double* z, long long n, double x0, double y0, double z0, double lon_or,
                    double lat_or, float* xe, float* ye, float* ze, cudaStream_t st);
int launch_ecef2enu_vector(const float* v, long long n, double lon_or, double lat_or, float* o, cudaStream_t st);
int launch_surf_norm(const double* lon, const double* lat, long long n, float* o, cudaStream_t st);
int launch_north_dir(int ellps, const double* x, const double* y, const double* z, const float* nrm, long long n, float* o, cudaStream_t st);
int launch_wgs2swiss(const double* lon, const double* lat, const float* h, long long n, double* e, double* nn, float* hc, cudaStream_t st);
int launch_swiss2wgs(const double* e, const double* nn, const float* hc, long long n, double* lon, double* lat, float* h, cudaStream_t st);
int launch_rotmat(const float* north, const float* norm, int ny, int nx, float* out, cudaStream_t st);
int launch_prep_enu(int ellps, const double* lon, const double* lat, const float* elev, int ny, int nx, double x0, double y0, double z0,
                    double lon_or, double lat_or, int off0, int off1, int in0, int in1, float* vert_grid, float* vec_norm,
                    float* vec_north, cudaStream_t st);
int parse_ellps(const char* s);      // sphere 0, GRS80 1, WGS84 2, unknown -1

// hostcopy.cu: staging engine of the host tier (pinned ring + worker threads)
void host_prefault(void* p, size_t bytes);                                       // populate fresh pages (content untouched)
int staged_d2h(void* dst_host, const void* src_dev, size_t bytes, cudaStream_t st);   // returns when dst_host is complete
int staged_h2d(void* dst_dev, const void* src_host, size_t bytes, cudaStream_t st);   // returns when src_host has been read

void* host_block_alloc(size_t bytes);      // pooled page-locked block for a large output array (nullptr: not available)
void host_block_free(void* p);
void host_block_trim();                    // release the idle page-locked blocks
bool host_is_pinned(const void* p);       // page-locked (ours or cudaHostRegister'ed by the caller)

// api.cu: pooled device buffers of the host tier
void* pool_alloc(size_t bytes);            // nullptr + error set on failure
void pool_free(void* p);
void pool_trim();

// number of SMs of the current device (cached)
int sm_count();

}  // namespace hzb
