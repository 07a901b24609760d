// api.cu -- the C ABI of libhorayzon_b200.so (see include/horayzon_b200.h).
//
// Host-side drivers standing where horizon_gridded_comp / horizon_locations_comp
// (horizon_comp.cpp:629-822, 828-1094) and shapes::CppTerrain
// (shadow_comp.cpp:304-605) stand in the reference.  No CPU fallback: every
// compute entry point needs a CUDA device and fails with a status otherwise.
#include "hzb_common.cuh"
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <chrono>
#include <limits>
#include <algorithm>
#include <time.h>
#include <stdlib.h>
#include <mutex>
#include <type_traits>
#include <thread>

namespace hzb {

static thread_local std::string g_error;
static thread_local hzb_stats g_stats;

void set_error(const std::string& msg) { g_error = msg; }
void set_last_stats(const hzb_stats& st) { g_stats = st; }
double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 148;
        cached = prop.multiProcessorCount; cached_dev = dev;
    }
    return cached;
}
int parse_algorithm(const char* s) {
    if (!s) return -1;
    if (!strcmp(s, "discrete_sampling")) return 0;
    if (!strcmp(s, "binary_search")) return 1;
    if (!strcmp(s, "guess_constant")) return 2;
    return -1;
}
int parse_geom_type(const char* s) {
    if (!s) return -1;
    if (!strcmp(s, "triangle")) return 0;
    if (!strcmp(s, "quad")) return 1;
    if (!strcmp(s, "grid")) return 2;
    return -1;
}

int parse_ellps(const char* s) {
    if (!s) return -1;
    if (!strcmp(s, "sphere")) return 0;
    if (!strcmp(s, "GRS80")) return 1;
    if (!strcmp(s, "WGS84")) return 2;
    return -1;
}

DebugOptions& debug_options() { static DebugOptions o; return o; }

// ---- device buffer pool: EVERY device allocation of the library (scene, BVH, build temporaries, per-call
// buffers of the host tier such as the 2 GB horizon array of a 1201 x 1201 x 360 call) comes from here and goes
// back here, so that a repeated host-tier call makes no cudaMalloc / cudaFree at all -- on some hosts those cost
// tens of milliseconds EACH (a BVH build with its 30 allocations was measured at 6 ms on one box and 1.1 s on
// another).  Idle memory is bounded (HZB_POOL_IDLE_MAX, 96 blocks) and released by hzb_trim().
namespace {
struct DevPool {
    std::mutex mu;
    struct Blk { void* p; size_t cap; int dev; };
    std::vector<Blk> idle, live;
    size_t idle_bytes = 0;
};
DevPool& dev_pool() { static DevPool* p = new DevPool(); return *p; }
constexpr size_t HZB_POOL_IDLE_MAX = (size_t)6 << 30;
}  // namespace

void* pool_alloc(size_t bytes) {
    if (bytes == 0) bytes = 1;
    int dev = 0; cudaGetDevice(&dev);
    DevPool& P = dev_pool();
    std::lock_guard<std::mutex> l(P.mu);
    int best = -1;
    for (int i = 0; i < (int)P.idle.size(); ++i)
        if (P.idle[i].dev == dev && P.idle[i].cap >= bytes && P.idle[i].cap <= bytes + bytes / 4 + 65536 &&
            (best < 0 || P.idle[i].cap < P.idle[best].cap)) best = i;
    DevPool::Blk b{nullptr, 0, dev};
    if (best >= 0) { b = P.idle[best]; P.idle.erase(P.idle.begin() + best); P.idle_bytes -= b.cap; }
    else {
        if (cudaMalloc(&b.p, bytes) != cudaSuccess) {
            cudaGetLastError();
            for (DevPool::Blk& x : P.idle) if (x.dev == dev) cudaFree(x.p);      // make room and retry once
            P.idle.erase(std::remove_if(P.idle.begin(), P.idle.end(), [&](const DevPool::Blk& x) { return x.dev == dev; }), P.idle.end());
            P.idle_bytes = 0; for (DevPool::Blk& x : P.idle) P.idle_bytes += x.cap;
            if (cudaMalloc(&b.p, bytes) != cudaSuccess) { set_error(std::string("cudaMalloc of ") + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(cudaGetLastError())); return nullptr; }
        }
        b.cap = bytes;
    }
    P.live.push_back(b);
    return b.p;
}
void pool_free(void* p) {
    if (!p) return;
    DevPool& P = dev_pool();
    std::lock_guard<std::mutex> l(P.mu);
    for (size_t i = 0; i < P.live.size(); ++i)
        if (P.live[i].p == p) {
            DevPool::Blk b = P.live[i];
            P.live.erase(P.live.begin() + i);
            if (b.cap > HZB_POOL_IDLE_MAX) { cudaFree(b.p); return; }
            P.idle.push_back(b); P.idle_bytes += b.cap;
            while (P.idle_bytes > HZB_POOL_IDLE_MAX || P.idle.size() > 96) {       // oldest first
                cudaFree(P.idle[0].p); P.idle_bytes -= P.idle[0].cap; P.idle.erase(P.idle.begin());
            }
            return;
        }
    cudaFree(p);
}
void pool_trim() {
    DevPool& P = dev_pool();
    std::lock_guard<std::mutex> l(P.mu);
    int cur = 0; cudaGetDevice(&cur);
    for (DevPool::Blk& b : P.idle) { cudaSetDevice(b.dev); cudaFree(b.p); }
    cudaSetDevice(cur);
    P.idle.clear(); P.idle_bytes = 0;
}

// ---- per-thread, per-device host-tier context: two streams (compute / copy) created once
struct HostCtx {
    int dev = -1; cudaStream_t comp = nullptr, copy = nullptr;
    cudaEvent_t done = nullptr;                                   // "compute stream finished" of the call in flight
    unsigned int* h_flags = nullptr; unsigned int* d_flags = nullptr; size_t flags_cap = 0;   // mapped host memory: row-block progress flags
    int flags(size_t n) {
        if (n > flags_cap) {
            if (h_flags) cudaFreeHost(h_flags);
            h_flags = nullptr; flags_cap = 0;
            const size_t cap = std::max<size_t>(n, 8192);
            HZB_CUDA(cudaHostAlloc((void**)&h_flags, cap * sizeof(unsigned int), cudaHostAllocMapped));
            HZB_CUDA(cudaHostGetDevicePointer((void**)&d_flags, h_flags, 0));
            flags_cap = cap;
        }
        memset(h_flags, 0, n * sizeof(unsigned int));
        return 0;
    }
};
static int host_ctx(HostCtx** out) {
    static thread_local std::vector<HostCtx> ctxs;
    int dev = 0;
    HZB_CUDA(cudaGetDevice(&dev));
    for (HostCtx& c : ctxs) if (c.dev == dev) { *out = &c; return 0; }
    HostCtx c; c.dev = dev;
    HZB_CUDA(cudaStreamCreateWithFlags(&c.comp, cudaStreamNonBlocking));
    HZB_CUDA(cudaStreamCreateWithFlags(&c.copy, cudaStreamNonBlocking));
    HZB_CUDA(cudaEventCreateWithFlags(&c.done, cudaEventDisableTiming));
    ctxs.push_back(c);
    *out = &ctxs.back();
    return 0;
}

static int require_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        set_error(std::string("no CUDA device available (libhorayzon_b200 has no CPU fallback): ") +
                  (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
        cudaGetLastError();
        return 1;
    }
    return 0;
}

static int read_counters(const Scene& s, hzb_stats& st) {
    Counters c;
    HZB_CUDA(cudaMemcpy(&c, s.d_counters, sizeof(c), cudaMemcpyDeviceToHost));
    st.rays = c.rays; st.node_visits = c.node_visits; st.prim_tests = c.prim_tests; st.units = c.units;
    st.warp_node_visits = c.warp_node_visits; st.fallback_packets = c.fallback_packets;
    st.segment_tasks = c.segment_tasks; st.segment_redos = c.segment_redos;
    st.num_prims = s.num_prims; st.num_nodes = s.num_nodes4; st.bvh_bytes = s.bvh_bytes;
    st.t_h2d = s.t_h2d; st.t_build = s.t_build;
    if (c.stack_overflow) { set_error("binary-BVH walker stack overflow (results invalid)"); return 1; }
    return 0;
}

template <typename T>
struct DevBuf {  // RAII device buffer (pooled: large blocks survive the call, see pool_alloc)
    T* p = nullptr;
    ~DevBuf() { if (p) pool_free(p); }
    int alloc(size_t n) { p = (T*)pool_alloc((n ? n : 1) * sizeof(T)); return p ? 0 : 1; }
    int upload(const T* h, size_t n) {
        HZB_TRY(alloc(n));
        HZB_CUDA(cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice));
        return 0;
    }
};

// ------------------------------------------------- coordinate preparation (row "next 3")
// Host-tier helpers: upload / download through the staging engine when the array is large.
template <typename T>
static int up_big(DevBuf<T>& d, const T* h, size_t n) {
    HZB_TRY(d.alloc(n));
    if (n * sizeof(T) >= ((size_t)8 << 20)) { HZB_TRY(staged_h2d(d.p, h, n * sizeof(T), nullptr)); }
    else HZB_CUDA(cudaMemcpy(d.p, h, n * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}
template <typename T>
static int down_big(T* h, const DevBuf<T>& d, size_t n) {
    HZB_CUDA(cudaStreamSynchronize(nullptr));
    if (n * sizeof(T) >= ((size_t)8 << 20)) { HZB_TRY(staged_d2h(h, d.p, n * sizeof(T), nullptr)); }
    else HZB_CUDA(cudaMemcpy(h, d.p, n * sizeof(T), cudaMemcpyDeviceToHost));
    return 0;
}

}  // namespace hzb

using namespace hzb;

struct hzb_scene { Scene s; };
struct hzb_terrain {
    Scene s; bool ready = false; TerrainParams tp{};
    float *d_tilt = nullptr, *d_norm = nullptr, *d_enl = nullptr, *d_elev = nullptr; uint8_t* d_mask = nullptr;
    void* d_out = nullptr; size_t out_cap = 0;
    // created once per terrain (not per sun position): compute / copy streams and the double-buffer events
    cudaStream_t s_comp = nullptr, s_copy = nullptr;
    cudaEvent_t done[2] = {nullptr, nullptr}, freed[2] = {nullptr, nullptr};
    int streams() {
        if (s_comp) return 0;
        HZB_CUDA(cudaStreamCreateWithFlags(&s_comp, cudaStreamNonBlocking));
        HZB_CUDA(cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            HZB_CUDA(cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming));
            HZB_CUDA(cudaEventCreateWithFlags(&freed[b], cudaEventDisableTiming));
        }
        return 0;
    }
    void release() {
        pool_free(d_tilt); pool_free(d_norm); pool_free(d_enl); pool_free(d_elev); pool_free(d_mask); pool_free(d_out);
        d_tilt = d_norm = d_enl = d_elev = nullptr; d_mask = nullptr; d_out = nullptr; out_cap = 0;
        if (s_comp) cudaStreamDestroy(s_comp);
        if (s_copy) cudaStreamDestroy(s_copy);
        for (int b = 0; b < 2; ++b) { if (done[b]) cudaEventDestroy(done[b]); if (freed[b]) cudaEventDestroy(freed[b]); done[b] = freed[b] = nullptr; }
        s_comp = s_copy = nullptr;
        scene_free(s); ready = false;
    }
};

extern "C" {

const char* hzb_last_error(void) { return g_error.c_str(); }
const char* hzb_version(void) { return "horayzon_b200 0.1 (sm_100a)"; }
int hzb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int hzb_get_stats(hzb_stats* out) { if (!out) return 1; *out = g_stats; return 0; }
// Pure host code (no device needed): the tables the kernels use, for inspection and for the CPU tests.
int hzb_horizon_tables(int azim_num, float dist_search, float hori_acc, float elev_ang_low_lim, int cap, float* elev_ang,
                       float* elev_sin, float* elev_cos, float* azim_sin, float* azim_cos) {
    if (azim_num < 1 || !(hori_acc > 0.f)) { set_error("invalid table parameters"); return -1; }
    HorizonTables T; T.make(azim_num, dist_search, hori_acc, elev_ang_low_lim);
    if (elev_ang && elev_sin && elev_cos && cap >= T.elev_num) {
        memcpy(elev_ang, T.elev_ang.data(), sizeof(float) * (size_t)T.elev_num);
        memcpy(elev_sin, T.elev_sin.data(), sizeof(float) * (size_t)T.elev_num);
        memcpy(elev_cos, T.elev_cos.data(), sizeof(float) * (size_t)T.elev_num);
    }
    if (azim_sin && azim_cos) {
        memcpy(azim_sin, T.azim_sin.data(), sizeof(float) * (size_t)azim_num);
        memcpy(azim_cos, T.azim_cos.data(), sizeof(float) * (size_t)azim_num);
    }
    return T.elev_num;
}
// Test-only switches (see DebugOptions).  Returns 0, or 1 for an unknown name.
int hzb_debug_option(const char* name, int value) {
    DebugOptions& o = debug_options();
    if (!name) { set_error("null option name"); return 1; }
    if (!strcmp(name, "reset")) { o = DebugOptions(); return 0; }
    if (!strcmp(name, "horizon_kernel")) { o.horizon_kernel = value; return 0; }
    if (!strcmp(name, "shadow_kernel")) { o.shadow_kernel = value; return 0; }
    if (!strcmp(name, "wrefill")) { o.wrefill = value; return 0; }
    if (!strcmp(name, "wwait")) { o.wwait = value; return 0; }
    if (!strcmp(name, "no_overlap")) { o.no_overlap = value; return 0; }
    if (!strcmp(name, "stack_limit")) { o.stack_limit = value; return 0; }
    if (!strcmp(name, "ctas_per_sm")) { o.ctas_per_sm = value; return 0; }
    if (!strcmp(name, "tail_segments")) { o.tail_segments = value; return 0; }
    if (!strcmp(name, "tail_tiles")) { o.tail_tiles = value; return 0; }
    if (!strcmp(name, "tail_band")) { o.tail_band = value; return 0; }
    set_error(std::string("unknown debug option ") + name);
    return 1;
}
// Additive, pure host code (works without a device; for the CPU test suite): the queue layout the horizon kernel would use
// for a launch -- see horizon.cu, "Queue order and azimuth segments".  out[0..6] = segments per split cell (1: none), first /
// end local block row and tile-column margin of the interior, split tiles, queue entries, tiles of the launch.
int hzb_plan_queue(int dem_dim_0, int dem_dim_1, const float* lo, const float* hi, int offset_0, int offset_1, int dim_in_0,
                   int dim_in_1, int row_begin, int row_end, int shard_rank, int shard_count, int azim_num, float dist_search,
                   float hori_acc, float elev_ang_low_lim, const char* ray_algorithm, int resident_ctas, long long* out) {
    if (!lo || !hi || !out || !ray_algorithm) { set_error("null argument"); return 1; }
    const int alg = parse_algorithm(ray_algorithm);
    if (alg < 0 || shard_count < 1 || shard_rank < 0 || shard_rank >= shard_count) { set_error("invalid argument"); return 1; }
    Scene s; s.H = dem_dim_0; s.W = dem_dim_1;
    for (int a = 0; a < 3; ++a) { s.lo[a] = lo[a]; s.hi[a] = hi[a]; }
    HorizonTables T; T.make(azim_num, dist_search, hori_acc, elev_ang_low_lim, false);
    HorizonParams p{};
    p.algorithm = alg; p.azim_num = azim_num; p.elev_num = T.elev_num; p.low = T.low; p.dist = T.dist;
    p.offset_0 = offset_0; p.offset_1 = offset_1; p.dim_in_0 = dim_in_0; p.dim_in_1 = dim_in_1; p.row_begin = row_begin; p.row_end = row_end;
    p.blk_stride = shard_count; p.blk_offset = shard_rank; p.row_full = (unsigned int)((dim_in_1 + 7) / 8) * 32u;
    plan_queue_host(s, p, resident_ctas);
    out[0] = p.seg_count; out[1] = p.q_by0; out[2] = p.q_by1; out[3] = p.q_bx; out[4] = p.q_tail; out[5] = p.q_total;
    out[6] = (long long)p.q_tiles_x * p.q_tiles_y;
    return 0;
}
// Release the idle pooled memory of this process (device blocks of the host tier, page-locked output blocks).
void hzb_trim(void) { pool_trim(); host_block_trim(); seg_pool_trim(); }
void* hzb_host_alloc(size_t bytes) { if (require_device()) return nullptr; return host_block_alloc(bytes); }
void hzb_host_free(void* p) { host_block_free(p); }

// ------------------------------------------------------------------ scene
hzb_scene* hzb_scene_create(const float* vert_grid, int dem_dim_0, int dem_dim_1, const float* vert_simp,
                            int num_vert_simp, const int32_t* tri_ind_simp, int num_tri_simp, int device) {
    if (require_device()) return nullptr;
    if (!vert_grid) { set_error("null pointer argument"); return nullptr; }
    // the kernels pack a cell as (row << 16) | column and the reference's wrapper has the same limit (horizon.pyx:149-151)
    if (dem_dim_0 > 32767 || dem_dim_1 > 32767) { set_error("maximal allowed input length for dem_dim_0 and dem_dim_1 is 32'767"); return nullptr; }
    if (num_vert_simp >= 3 && num_tri_simp > 0) {
        if (!vert_simp || !tri_ind_simp) { set_error("null pointer argument"); return nullptr; }
        const size_t m = (size_t)num_tri_simp * 3;
        int32_t lo = 0, hi = 0;
        for (size_t i = 0; i < m; ++i) { lo = std::min(lo, tri_ind_simp[i]); hi = std::max(hi, tri_ind_simp[i]); }
        if (lo < 0 || hi >= num_vert_simp) { set_error("triangle indices of simplified outer domain exceed number of vertices"); return nullptr; }
    }
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice failed"); cudaGetLastError(); return nullptr; }
    hzb_scene* h = new hzb_scene();
    h->s.device = device;
    if (scene_upload_and_build(h->s, vert_grid, dem_dim_0, dem_dim_1, vert_simp, num_vert_simp, tri_ind_simp, num_tri_simp)) {
        scene_free(h->s); delete h; return nullptr;
    }
    return h;
}
void hzb_scene_destroy(hzb_scene* h) { if (h) { cudaSetDevice(h->s.device); scene_free(h->s); delete h; } }
int hzb_scene_stats(const hzb_scene* h, hzb_stats* out) {
    if (!h || !out) { set_error("null argument"); return 1; }
    memset(out, 0, sizeof(*out));
    return read_counters(h->s, *out);
}

static int horizon_gridded_launch(Scene& sc, const float* d_vec_norm, const float* d_vec_north, const uint8_t* d_mask,
                                  int offset_0, int offset_1, int dim_in_0, int dim_in_1, int row_begin, int row_end,
                                  int azim_num, float dist_search, float hori_acc, const char* ray_algorithm,
                                  float elev_ang_low_lim, float hori_fill, float ray_org_elev, float* d_hori_buffer,
                                  unsigned int* d_row_done, volatile unsigned int* row_flags, cudaStream_t st, int azim_first = 0,
                                  int shard_rank = 0, int shard_count = 1, int packed = 0,
                                  uint16_t* d_hori_q = nullptr, float* d_hori_first = nullptr) {
    const int alg = parse_algorithm(ray_algorithm);
    if (alg < 0) { set_error("invalid input argument for ray_algorithm"); return 1; }
    if (azim_num < 1 || dim_in_0 < 0 || dim_in_1 < 0 || row_begin < 0 || row_end > dim_in_0) { set_error("invalid dimensions"); return 1; }
    if (offset_0 < 0 || offset_1 < 0 || offset_0 + dim_in_0 > sc.H || offset_1 + dim_in_1 > sc.W) {
        set_error("inner domain exceeds DEM dimensions"); return 1;
    }
    if (!(hori_acc > 0.f)) { set_error("hori_acc must be positive"); return 1; }
    if (shard_count < 1 || shard_rank < 0 || shard_rank >= shard_count) { set_error("invalid shard"); return 1; }
    if ((shard_count > 1 || packed) && (d_row_done || azim_first)) { set_error("block sharding needs the reference layout and no progress counters"); return 1; }
    HZB_CUDA(cudaSetDevice(sc.device));
    HorizonParams p{};
    HZB_TRY(scene_tables(sc, azim_num, dist_search, hori_acc, elev_ang_low_lim, p, st));
    p.algorithm = alg; p.vec_norm = d_vec_norm; p.vec_north = d_vec_north; p.mask = d_mask;
    p.offset_0 = offset_0; p.offset_1 = offset_1; p.dim_in_0 = dim_in_0; p.dim_in_1 = dim_in_1;
    p.row_begin = row_begin; p.row_end = row_end; p.hori_fill = hori_fill; p.ray_org_elev = ray_org_elev;
    p.hori = d_hori_buffer; p.row_done = d_row_done; p.row_flags = row_flags;
    p.hori_q = d_hori_q; p.hori_first = d_hori_first;
    if ((d_hori_q == nullptr) != (d_hori_first == nullptr)) { set_error("quantised output needs both buffers"); return 1; }
    if (d_hori_q && (azim_first || packed)) { set_error("quantised output needs the reference layout"); return 1; }
    p.row_full = (unsigned int)((dim_in_1 + 7) / 8) * 32u;
    p.blk_stride = shard_count; p.blk_offset = shard_rank; p.packed = packed;
    p.stride_c = azim_first ? 1 : azim_num;
    p.stride_k = azim_first ? (long long)dim_in_0 * dim_in_1 : 1;
    return launch_horizon_gridded(sc, p, st);
}

int hzb_horizon_gridded_dev(hzb_scene* h, const float* d_vec_norm, const float* d_vec_north, const uint8_t* d_mask,
                            int offset_0, int offset_1, int dim_in_0, int dim_in_1, int row_begin, int row_end,
                            int azim_num, float dist_search, float hori_acc, const char* ray_algorithm,
                            float elev_ang_low_lim, float hori_fill, float ray_org_elev, float* d_hori_buffer,
                            void* stream) {
    if (!h) { set_error("null scene"); return 1; }
    return horizon_gridded_launch(h->s, d_vec_norm, d_vec_north, d_mask, offset_0, offset_1, dim_in_0, dim_in_1, row_begin,
                                  row_end, azim_num, dist_search, hori_acc, ray_algorithm, elev_ang_low_lim, hori_fill,
                                  ray_org_elev, d_hori_buffer, nullptr, nullptr, (cudaStream_t)stream);
}

// additive (scope row "next 4"): azimuth-first output [azim_num][dim_in_0][dim_in_1], the layout the
// reference's examples transpose to before writing NetCDF (examples/horizon/gridded_curved_DEM.py:113-125)
int hzb_horizon_gridded_dev_layout(hzb_scene* h, const float* d_vec_norm, const float* d_vec_north, const uint8_t* d_mask,
                                   int offset_0, int offset_1, int dim_in_0, int dim_in_1, int row_begin, int row_end,
                                   int azim_num, float dist_search, float hori_acc, const char* ray_algorithm,
                                   float elev_ang_low_lim, float hori_fill, float ray_org_elev, float* d_hori_buffer,
                                   int azim_first, void* stream) {
    if (!h) { set_error("null scene"); return 1; }
    return horizon_gridded_launch(h->s, d_vec_norm, d_vec_north, d_mask, offset_0, offset_1, dim_in_0, dim_in_1, row_begin,
                                  row_end, azim_num, dist_search, hori_acc, ray_algorithm, elev_ang_low_lim, hori_fill,
                                  ray_org_elev, d_hori_buffer, nullptr, nullptr, (cudaStream_t)stream, azim_first);
}

// additive (scope row 8f-4): quantised output.  Every guess_constant result but the first azimuth's is a table
// entry (horizon_comp.cpp:490-494), so d_idx_buffer [dim_in_0][dim_in_1][azim_num] receives 16-bit table indices
// (0xFFFF: "take the cell's float": azimuth 0 and masked cells) and d_first_buffer [dim_in_0][dim_in_1] the first
// azimuth's un-quantised midpoint (:428) or hori_fill.  Lossless: hzb_horizon_tables gives elev_ang, and
// elev_ang[index] is bit for bit the float the other entry points store.  Half the bytes to move and keep.
int hzb_horizon_gridded_dev_quantised(hzb_scene* h, const float* d_vec_norm, const float* d_vec_north, const uint8_t* d_mask,
                                      int offset_0, int offset_1, int dim_in_0, int dim_in_1, int row_begin, int row_end,
                                      int azim_num, float dist_search, float hori_acc, float elev_ang_low_lim, float hori_fill,
                                      float ray_org_elev, uint16_t* d_idx_buffer, float* d_first_buffer, void* stream) {
    if (!h) { set_error("null scene"); return 1; }
    if (!d_idx_buffer || !d_first_buffer) { set_error("null pointer argument"); return 1; }
    return horizon_gridded_launch(h->s, d_vec_norm, d_vec_north, d_mask, offset_0, offset_1, dim_in_0, dim_in_1, row_begin, row_end,
                                  azim_num, dist_search, hori_acc, "guess_constant", elev_ang_low_lim, hori_fill, ray_org_elev,
                                  nullptr, nullptr, nullptr, (cudaStream_t)stream, 0, 0, 1, 0, d_idx_buffer, d_first_buffer);
}

// additive (multi-GPU): shard `shard_rank` of `shard_count` computes the 4-row blocks b of the inner domain with
// b % shard_count == shard_rank (interleaved, so every shard sees the same mix of cheap rim rows and expensive
// centre rows).  packed == 0: results go to their places in the full [dim_in_0][dim_in_1][azim_num] array;
// packed != 0: the shard's blocks are stored back to back from d_hori_buffer on -- a contiguous all-gather send
// buffer of hzb_shard_rows(dim_in_0, rank, count) rows.
int hzb_horizon_gridded_dev_sharded(hzb_scene* h, const float* d_vec_norm, const float* d_vec_north, const uint8_t* d_mask,
                                    int offset_0, int offset_1, int dim_in_0, int dim_in_1, int azim_num, float dist_search,
                                    float hori_acc, const char* ray_algorithm, float elev_ang_low_lim, float hori_fill,
                                    float ray_org_elev, float* d_hori_buffer, int shard_rank, int shard_count, int packed,
                                    void* stream) {
    if (!h) { set_error("null scene"); return 1; }
    return horizon_gridded_launch(h->s, d_vec_norm, d_vec_north, d_mask, offset_0, offset_1, dim_in_0, dim_in_1, 0, dim_in_0,
                                  azim_num, dist_search, hori_acc, ray_algorithm, elev_ang_low_lim, hori_fill, ray_org_elev,
                                  d_hori_buffer, nullptr, nullptr, (cudaStream_t)stream, 0, shard_rank, shard_count, packed);
}
// rows (padded to whole 4-row blocks) a shard's packed buffer holds
int hzb_shard_rows(int dim_in_0, int shard_rank, int shard_count) {
    if (shard_count < 1 || shard_rank < 0 || shard_rank >= shard_count || dim_in_0 < 0) return -1;
    const int all = (dim_in_0 + 3) / 4;
    return all > shard_rank ? 4 * ((all - shard_rank + shard_count - 1) / shard_count) : 0;
}

// -------------------------------------------------------------- host tier
// One call = H2D of the inputs, on-device BVH build, horizon kernel, (optionally) the SVF integral on
// the device-resident horizon, D2H.  The kernel runs on one stream; finished 4-row blocks -- the last
// cell of a block raises a flag in mapped host memory, so the host polls plain memory, no CUDA call --
// leave for the caller's array on a second stream while the kernel is still running.
static int horizon_gridded_host(const float* vert_grid, int dem_dim_0, int dem_dim_1, const float* vec_norm,
                        const float* vec_north, int offset_0, int offset_1, float* hori_buffer, int dim_in_0,
                        int dim_in_1, int azim_num, float dist_search, float hori_acc, const char* ray_algorithm,
                        const char* geom_type, const float* vert_simp, int num_vert_simp,
                        const int32_t* tri_ind_simp, int num_tri_simp, float elev_ang_low_lim, const uint8_t* mask,
                        float hori_fill, float ray_org_elev, int azim_first,
                        const float* svf_vec_tilt = nullptr, float* svf_out = nullptr) {
    memset(&g_stats, 0, sizeof(g_stats));
    const double t_start = now_s();
    const bool timing = getenv("HZB_TIMING") != nullptr;   // stderr breakdown of this call
    if (require_device()) return 1;
    if (parse_geom_type(geom_type) < 0) { set_error("invalid input argument for geom_type"); return 1; }
    if (!vert_grid || !vec_norm || !vec_north || !hori_buffer || !mask) { set_error("null pointer argument"); return 1; }
    if ((svf_vec_tilt == nullptr) != (svf_out == nullptr)) { set_error("svf_vec_tilt and svf_out go together"); return 1; }
    if (svf_out && azim_num < 2) { set_error("the sky view factor needs at least two azimuth sectors"); return 1; }
    int dev = 0; cudaGetDevice(&dev);
    const double t_scene0 = now_s();
    hzb_scene* h = hzb_scene_create(vert_grid, dem_dim_0, dem_dim_1, vert_simp, num_vert_simp, tri_ind_simp, num_tri_simp, dev);
    if (!h) return 1;
    struct Guard { hzb_scene* h; ~Guard() { hzb_scene_destroy(h); } } guard{h};
    const size_t nc = (size_t)dim_in_0 * dim_in_1;
    const double t_scene = now_s() - t_scene0;
    double t0 = now_s();
    HostCtx* ctx = nullptr;
    HZB_TRY(host_ctx(&ctx));
    cudaStream_t s_comp = ctx->comp, s_copy = ctx->copy;
    DevBuf<float> d_norm, d_north, d_hori, d_tilt, d_azim, d_svf; DevBuf<uint8_t> d_mask;
    HZB_TRY(d_norm.upload(vec_norm, nc * 3)); HZB_TRY(d_north.upload(vec_north, nc * 3)); HZB_TRY(d_mask.upload(mask, nc));
    HZB_TRY(d_hori.alloc(nc * (size_t)azim_num));
    if (svf_out) {
        std::vector<float> az((size_t)azim_num);
        for (int i = 0; i < azim_num; ++i) az[i] = (float)((2 * M_PI) / azim_num * i);    // horizon.pyx:190-195
        HZB_TRY(d_tilt.upload(svf_vec_tilt, nc * 3)); HZB_TRY(d_azim.upload(az.data(), (size_t)azim_num));
        HZB_TRY(d_svf.alloc(nc));
    }
    const double t_h2d_extra = now_s() - t0;
    t0 = now_s();
    const int tiles_y = (dim_in_0 + 3) / 4;
    // (rows are contiguous byte ranges only in the reference layout: azimuth-first output is copied after the kernel)
    const DebugOptions& dbg = debug_options();
    const bool overlap = !azim_first && !dbg.no_overlap && dbg.horizon_kernel == 0;
    DevBuf<unsigned int> d_done;
    unsigned int* h_flags = nullptr; unsigned int* d_flags = nullptr;
    if (overlap) {
        HZB_TRY(d_done.alloc((size_t)tiles_y));
        HZB_CUDA(cudaMemsetAsync(d_done.p, 0, (size_t)tiles_y * sizeof(unsigned int), s_comp));
        HZB_TRY(ctx->flags((size_t)tiles_y));          // persistent mapped buffer of this thread / device, cleared
        h_flags = ctx->h_flags; d_flags = ctx->d_flags;
    }
    HZB_TRY(horizon_gridded_launch(h->s, d_norm.p, d_north.p, d_mask.p, offset_0, offset_1, dim_in_0, dim_in_1, 0, dim_in_0,
                                   azim_num, dist_search, hori_acc, ray_algorithm, elev_ang_low_lim, hori_fill,
                                   ray_org_elev, d_hori.p, overlap ? d_done.p : nullptr, overlap ? d_flags : nullptr, s_comp, azim_first));
    if (svf_out && nc > 0) {   // fused epilogue: the integral reads the horizon where the kernel left it (HBM), azimuth-first aware
        if (azim_first) { set_error("the fused sky view factor needs the reference layout (azim_first = 0)"); cudaStreamSynchronize(s_comp); return 1; }
        HZB_TRY(launch_svf(0, d_azim.p, d_hori.p, d_tilt.p, (long long)nc, azim_num, d_svf.p, s_comp));
    }
    cudaEvent_t ev_done = ctx->done;
    HZB_CUDA(cudaEventRecord(ev_done, s_comp));
    double t_d2h = 0.0, t_trace = 0.0;
    const size_t row_elems = (size_t)dim_in_1 * (size_t)azim_num;
    const double t_pf0 = now_s();
    const bool out_pinned = host_is_pinned(hori_buffer);   // page-locked output (hzb_host_alloc): plain DMA, no staging
    if (!out_pinned) host_prefault(hori_buffer, nc * (size_t)azim_num * sizeof(float));   // the kernel is running meanwhile
    const double t_prefault = now_s() - t_pf0;
    if (overlap) {
        const volatile unsigned int* flags = h_flags;
        const int min_blocks = std::max(1, (int)((64u << 20) / (row_elems * 4 * sizeof(float)) ));  // >= 64 MB per copy
        int copied_blocks = 0;
        while (copied_blocks < tiles_y) {
            const bool finished = cudaEventQuery(ev_done) == cudaSuccess;
            if (finished && t_trace == 0.0) t_trace = now_s() - t0;
            int ready = copied_blocks;
            while (ready < tiles_y && flags[ready] != 0u) ++ready;
            if (finished && ready < tiles_y) {
                cudaError_t e = cudaGetLastError();
                set_error(std::string("horizon kernel ended with unfinished rows") + (e != cudaSuccess ? std::string(": ") + cudaGetErrorString(e) : ""));
                return 1;
            }
            if (ready - copied_blocks >= min_blocks || (ready == tiles_y && ready > copied_blocks)) {
                const size_t r0 = (size_t)copied_blocks * 4, r1 = std::min<size_t>((size_t)ready * 4, (size_t)dim_in_0);
                const double tc = now_s();
                if (out_pinned) {
                    HZB_CUDA(cudaMemcpyAsync(hori_buffer + r0 * row_elems, d_hori.p + r0 * row_elems, (r1 - r0) * row_elems * sizeof(float),
                                             cudaMemcpyDeviceToHost, s_copy));
                    if (ready == tiles_y) HZB_CUDA(cudaStreamSynchronize(s_copy));   // earlier blocks keep flowing while we poll
                } else {
                    HZB_TRY(staged_d2h(hori_buffer + r0 * row_elems, d_hori.p + r0 * row_elems, (r1 - r0) * row_elems * sizeof(float), s_copy));
                }
                t_d2h += now_s() - tc;
                copied_blocks = ready;
            } else {
                std::this_thread::sleep_for(std::chrono::microseconds(100));
            }
        }
        HZB_CUDA(cudaStreamSynchronize(s_comp));
        HZB_CUDA(cudaStreamSynchronize(s_copy));
        if (t_trace == 0.0) t_trace = now_s() - t0;
        // Row blocks the fix-up kernel rewrote (cells whose traversal stack was full, azimuth segments that started
        // from a wrong index) AFTER they had been copied carry flag 2: they are copied again (rare; the tests force it).
        for (int b = 0; b < tiles_y; ++b) {
            if (flags[b] != 2u) continue;
            int e = b + 1;
            while (e < tiles_y && flags[e] == 2u) ++e;
            const size_t r0 = (size_t)b * 4, r1 = std::min<size_t>((size_t)e * 4, (size_t)dim_in_0);
            if (out_pinned) HZB_CUDA(cudaMemcpy(hori_buffer + r0 * row_elems, d_hori.p + r0 * row_elems, (r1 - r0) * row_elems * sizeof(float), cudaMemcpyDeviceToHost));
            else HZB_TRY(staged_d2h(hori_buffer + r0 * row_elems, d_hori.p + r0 * row_elems, (r1 - r0) * row_elems * sizeof(float), nullptr));
            b = e - 1;
        }
    } else {
        HZB_CUDA(cudaStreamSynchronize(s_comp));
        t_trace = now_s() - t0;
        const double tc = now_s();
        if (out_pinned) HZB_CUDA(cudaMemcpy(hori_buffer, d_hori.p, nc * (size_t)azim_num * sizeof(float), cudaMemcpyDeviceToHost));
        else HZB_TRY(staged_d2h(hori_buffer, d_hori.p, nc * (size_t)azim_num * sizeof(float), nullptr));
        t_d2h = now_s() - tc;
    }
    if (svf_out && nc > 0) HZB_CUDA(cudaMemcpy(svf_out, d_svf.p, nc * sizeof(float), cudaMemcpyDeviceToHost));
    HZB_CUDA(cudaGetLastError());
    HZB_TRY(read_counters(h->s, g_stats));
    g_stats.t_h2d += t_h2d_extra; g_stats.t_trace = t_trace; g_stats.t_d2h = t_d2h; g_stats.t_total = now_s() - t_start;
    if (timing)
        fprintf(stderr, "[hzb] horizon_gridded: scene %.3f (h2d %.3f build %.3f) inputs+alloc %.3f launch..end %.3f (prefault %.3f, kernel seen done %.3f, d2h calls %.3f) total %.3f s\n",
                t_scene, h->s.t_h2d, h->s.t_build, t_h2d_extra, now_s() - t0, t_prefault, t_trace, t_d2h, g_stats.t_total);
    return 0;
}

int hzb_horizon_gridded(const float* vert_grid, int dem_dim_0, int dem_dim_1, const float* vec_norm,
                        const float* vec_north, int offset_0, int offset_1, float* hori_buffer, int dim_in_0,
                        int dim_in_1, int azim_num, float dist_search, float hori_acc, const char* ray_algorithm,
                        const char* geom_type, const float* vert_simp, int num_vert_simp,
                        const int32_t* tri_ind_simp, int num_tri_simp, float elev_ang_low_lim, const uint8_t* mask,
                        float hori_fill, float ray_org_elev) {
    return horizon_gridded_host(vert_grid, dem_dim_0, dem_dim_1, vec_norm, vec_north, offset_0, offset_1, hori_buffer, dim_in_0,
                                dim_in_1, azim_num, dist_search, hori_acc, ray_algorithm, geom_type, vert_simp, num_vert_simp,
                                tri_ind_simp, num_tri_simp, elev_ang_low_lim, mask, hori_fill, ray_org_elev, 0);
}
// additive: same call, output [azim_num][dim_in_0][dim_in_1] when azim_first != 0
int hzb_horizon_gridded_layout(const float* vert_grid, int dem_dim_0, int dem_dim_1, const float* vec_norm,
                               const float* vec_north, int offset_0, int offset_1, float* hori_buffer, int dim_in_0,
                               int dim_in_1, int azim_num, float dist_search, float hori_acc, const char* ray_algorithm,
                               const char* geom_type, const float* vert_simp, int num_vert_simp,
                               const int32_t* tri_ind_simp, int num_tri_simp, float elev_ang_low_lim, const uint8_t* mask,
                               float hori_fill, float ray_org_elev, int azim_first) {
    return horizon_gridded_host(vert_grid, dem_dim_0, dem_dim_1, vec_norm, vec_north, offset_0, offset_1, hori_buffer, dim_in_0,
                                dim_in_1, azim_num, dist_search, hori_acc, ray_algorithm, geom_type, vert_simp, num_vert_simp,
                                tri_ind_simp, num_tri_simp, elev_ang_low_lim, mask, hori_fill, ray_org_elev, azim_first);
}

// host tier of the quantised output (guess_constant only; see hzb_horizon_gridded_dev_quantised)
int hzb_horizon_gridded_quantised(const float* vert_grid, int dem_dim_0, int dem_dim_1, const float* vec_norm,
                                  const float* vec_north, int offset_0, int offset_1, uint16_t* idx_buffer, float* first_buffer,
                                  int dim_in_0, int dim_in_1, int azim_num, float dist_search, float hori_acc,
                                  const char* geom_type, const float* vert_simp, int num_vert_simp,
                                  const int32_t* tri_ind_simp, int num_tri_simp, float elev_ang_low_lim, const uint8_t* mask,
                                  float hori_fill, float ray_org_elev) {
    memset(&g_stats, 0, sizeof(g_stats));
    const double t_start = now_s();
    if (require_device()) return 1;
    if (parse_geom_type(geom_type) < 0) { set_error("invalid input argument for geom_type"); return 1; }
    if (!vert_grid || !vec_norm || !vec_north || !idx_buffer || !first_buffer || !mask) { set_error("null pointer argument"); return 1; }
    int dev = 0; cudaGetDevice(&dev);
    hzb_scene* h = hzb_scene_create(vert_grid, dem_dim_0, dem_dim_1, vert_simp, num_vert_simp, tri_ind_simp, num_tri_simp, dev);
    if (!h) return 1;
    struct Guard { hzb_scene* h; ~Guard() { hzb_scene_destroy(h); } } guard{h};
    const size_t nc = (size_t)dim_in_0 * dim_in_1;
    HostCtx* ctx = nullptr;
    HZB_TRY(host_ctx(&ctx));
    double t0 = now_s();
    DevBuf<float> d_norm, d_north, d_first; DevBuf<uint8_t> d_mask; DevBuf<uint16_t> d_idx;
    HZB_TRY(d_norm.upload(vec_norm, nc * 3)); HZB_TRY(d_north.upload(vec_north, nc * 3)); HZB_TRY(d_mask.upload(mask, nc));
    HZB_TRY(d_idx.alloc(nc * (size_t)azim_num)); HZB_TRY(d_first.alloc(nc));
    const double t_h2d_extra = now_s() - t0;
    t0 = now_s();
    HZB_TRY(hzb_horizon_gridded_dev_quantised(h, d_norm.p, d_north.p, d_mask.p, offset_0, offset_1, dim_in_0, dim_in_1, 0, dim_in_0, azim_num,
                                              dist_search, hori_acc, elev_ang_low_lim, hori_fill, ray_org_elev, d_idx.p, d_first.p, ctx->comp));
    host_prefault(idx_buffer, nc * (size_t)azim_num * sizeof(uint16_t));      // while the kernel runs
    HZB_CUDA(cudaStreamSynchronize(ctx->comp));
    const double t_trace = now_s() - t0;
    t0 = now_s();
    if (nc > 0) {
        if (host_is_pinned(idx_buffer)) HZB_CUDA(cudaMemcpy(idx_buffer, d_idx.p, nc * (size_t)azim_num * sizeof(uint16_t), cudaMemcpyDeviceToHost));
        else HZB_TRY(staged_d2h(idx_buffer, d_idx.p, nc * (size_t)azim_num * sizeof(uint16_t), nullptr));
        HZB_CUDA(cudaMemcpy(first_buffer, d_first.p, nc * sizeof(float), cudaMemcpyDeviceToHost));
    }
    HZB_CUDA(cudaGetLastError());
    HZB_TRY(read_counters(h->s, g_stats));
    g_stats.t_h2d += t_h2d_extra; g_stats.t_trace = t_trace; g_stats.t_d2h = now_s() - t0; g_stats.t_total = now_s() - t_start;
    return 0;
}

// additive: horizon + sky view factor in one call.  The integral (topo_param.pyx:412-460) runs on the
// device-resident horizon right behind the search, so the 2 GB-scale array is never uploaded again
// (examples/horizon/gridded_curved_DEM.py:104-144 calls the two back to back).
int hzb_horizon_gridded_svf(const float* vert_grid, int dem_dim_0, int dem_dim_1, const float* vec_norm,
                            const float* vec_north, int offset_0, int offset_1, float* hori_buffer, int dim_in_0,
                            int dim_in_1, int azim_num, float dist_search, float hori_acc, const char* ray_algorithm,
                            const char* geom_type, const float* vert_simp, int num_vert_simp,
                            const int32_t* tri_ind_simp, int num_tri_simp, float elev_ang_low_lim, const uint8_t* mask,
                            float hori_fill, float ray_org_elev, const float* vec_tilt, float* svf_buffer) {
    if (!vec_tilt || !svf_buffer) { set_error("null pointer argument"); return 1; }
    return horizon_gridded_host(vert_grid, dem_dim_0, dem_dim_1, vec_norm, vec_north, offset_0, offset_1, hori_buffer, dim_in_0,
                                dim_in_1, azim_num, dist_search, hori_acc, ray_algorithm, geom_type, vert_simp, num_vert_simp,
                                tri_ind_simp, num_tri_simp, elev_ang_low_lim, mask, hori_fill, ray_org_elev, 0, vec_tilt, svf_buffer);
}

int hzb_horizon_locations(const float* vert_grid, int dem_dim_0, int dem_dim_1, const float* coords,
                          const float* vec_norm, const float* vec_north, float* hori_buffer, float* hori_dist_buffer,
                          int num_loc, int azim_num, float dist_search, float hori_acc, const char* ray_algorithm,
                          const char* geom_type, float elev_ang_low_lim, const float* ray_org_elev, int hori_dist_out) {
    memset(&g_stats, 0, sizeof(g_stats));
    const double t_start = now_s();
    if (require_device()) return 1;
    const int alg = parse_algorithm(ray_algorithm);
    if (alg < 0) { set_error("invalid input argument for ray_algorithm"); return 1; }
    if (parse_geom_type(geom_type) < 0) { set_error("invalid input argument for geom_type"); return 1; }
    if (hori_dist_out && alg == 2) { set_error("horizon detection algorithm 'guess_constant' not implemented for horizon distance computation"); return 1; }
    if (num_loc < 0 || azim_num < 1 || !(hori_acc > 0.f)) { set_error("invalid dimensions"); return 1; }
    int dev = 0; cudaGetDevice(&dev);
    hzb_scene* h = hzb_scene_create(vert_grid, dem_dim_0, dem_dim_1, nullptr, 0, nullptr, 0, dev);  // no TIN (:848-852)
    if (!h) return 1;
    struct Guard { hzb_scene* h; ~Guard() { hzb_scene_destroy(h); } } guard{h};
    if (num_loc == 0) return 0;
    const size_t n = (size_t)num_loc, no = n * (size_t)azim_num;
    DevBuf<float> d_coords, d_norm, d_north, d_elev, d_hori, d_dist;
    HZB_TRY(d_coords.upload(coords, n * 3)); HZB_TRY(d_norm.upload(vec_norm, n * 3)); HZB_TRY(d_north.upload(vec_north, n * 3));
    HZB_TRY(d_elev.upload(ray_org_elev, n));
    HZB_TRY(d_hori.upload(hori_buffer, no));  // keeps the caller's pre-fill for skipped locations
    if (hori_dist_out) HZB_TRY(d_dist.upload(hori_dist_buffer, no));
    HorizonParams p{};
    HZB_TRY(scene_tables(h->s, azim_num, dist_search, hori_acc, elev_ang_low_lim, p, nullptr));
    p.algorithm = alg;
    LocationParams lp{d_coords.p, d_norm.p, d_north.p, d_elev.p, d_hori.p, d_dist.p, num_loc, hori_dist_out};
    double t0 = now_s();
    HZB_TRY(launch_horizon_locations(h->s, p, lp, nullptr));
    HZB_CUDA(cudaDeviceSynchronize());
    const double t_trace = now_s() - t0;
    HZB_CUDA(cudaMemcpy(hori_buffer, d_hori.p, no * sizeof(float), cudaMemcpyDeviceToHost));
    if (hori_dist_out) HZB_CUDA(cudaMemcpy(hori_dist_buffer, d_dist.p, no * sizeof(float), cudaMemcpyDeviceToHost));
    HZB_TRY(read_counters(h->s, g_stats));
    g_stats.t_trace = t_trace; g_stats.t_total = now_s() - t_start;
    return 0;
}

// ---------------------------------------------------------------- terrain
hzb_terrain* hzb_terrain_create(void) { return new hzb_terrain(); }
void hzb_terrain_destroy(hzb_terrain* t) { if (t) { if (t->ready) cudaSetDevice(t->s.device); t->release(); delete t; } }

int hzb_terrain_initialise(hzb_terrain* t, const float* vert_grid, int dem_dim_0, int dem_dim_1, int offset_0,
                           int offset_1, const float* vec_tilt, const float* vec_norm, int dim_in_0, int dim_in_1,
                           const float* surf_enl_fac, const float* elevation, const uint8_t* mask,
                           const char* geom_type, float sw_dir_cor_fill, float ang_max, int refrac_cor) {
    if (!t) { set_error("null terrain"); return 1; }
    if (require_device()) return 1;
    if (parse_geom_type(geom_type) < 0) { set_error("invalid input argument for geom_type"); return 1; }
    if (offset_0 < 0 || offset_1 < 0 || offset_0 + dim_in_0 > dem_dim_0 || offset_1 + dim_in_1 > dem_dim_1) {
        set_error("inner domain exceeds DEM dimensions"); return 1;
    }
    t->release();
    int dev = 0; cudaGetDevice(&dev);
    t->s.device = dev;
    HZB_TRY(scene_upload_and_build(t->s, vert_grid, dem_dim_0, dem_dim_1, nullptr, 0, nullptr, 0));
    const size_t nc = (size_t)dim_in_0 * dim_in_1;
    auto up = [&](auto** d, const auto* h, size_t n) -> int {
        *d = (std::remove_reference_t<decltype(*d)>)pool_alloc((n ? n : 1) * sizeof(**d));
        if (!*d) return 1;
        HZB_CUDA(cudaMemcpy(*d, h, n * sizeof(**d), cudaMemcpyHostToDevice));
        return 0;
    };
    HZB_TRY(up(&t->d_tilt, vec_tilt, nc * 3)); HZB_TRY(up(&t->d_norm, vec_norm, nc * 3));
    HZB_TRY(up(&t->d_enl, surf_enl_fac, nc)); HZB_TRY(up(&t->d_elev, elevation, nc)); HZB_TRY(up(&t->d_mask, mask, nc));
    TerrainParams& tp = t->tp;
    tp.vec_tilt = t->d_tilt; tp.vec_norm = t->d_norm; tp.surf_enl_fac = t->d_enl; tp.elevation = t->d_elev; tp.mask = t->d_mask;
    tp.offset_0 = offset_0; tp.offset_1 = offset_1; tp.dim_in_0 = dim_in_0; tp.dim_in_1 = dim_in_1;
    tp.sw_dir_cor_fill = sw_dir_cor_fill; tp.ang_max = ang_max; tp.refrac_cor = refrac_cor;
    tp.dot_prod_min = cosf((float)(((double)ang_max / 180.0) * M_PI));  // shadow_comp.cpp:498
    tp.t_ref = 283.15f; tp.p_ref = 101.0f; tp.lapse = 0.0065f;           // :349-354
    const float g = 9.81f, R_d = 287.0f;
    tp.expo = g / (R_d * tp.lapse);
    t->ready = true;
    return 0;
}

static int terrain_out(hzb_terrain* t, size_t bytes) {
    if (t->out_cap < bytes) {
        pool_free(t->d_out); t->d_out = nullptr; t->out_cap = 0;
        t->d_out = pool_alloc(bytes);
        if (!t->d_out) return 1;
        t->out_cap = bytes;
    }
    return 0;
}

int hzb_terrain_shadow_dev(hzb_terrain* t, const float* sun, uint8_t* d_out, void* stream) {
    if (!t || !t->ready) { set_error("terrain not initialised"); return 1; }
    HZB_CUDA(cudaSetDevice(t->s.device));
    return launch_shadow(t->s, t->tp, sun, d_out, (cudaStream_t)stream);
}
int hzb_terrain_sw_dir_cor_dev(hzb_terrain* t, const float* sun, float* d_out, void* stream) {
    if (!t || !t->ready) { set_error("terrain not initialised"); return 1; }
    HZB_CUDA(cudaSetDevice(t->s.device));
    return launch_sw_dir_cor(t->s, t->tp, sun, d_out, (cudaStream_t)stream);
}
}  // extern "C"

// n_sun positions: kernel k runs on one stream while the result of position k-1 is
// copied to the caller's array on another (D2H and its page faults hide behind compute).
template <typename T, typename Launch>
static int terrain_batch(hzb_terrain* t, const float* suns, int n_sun, T* out, Launch launch) {
    if (!t || !t->ready) { set_error("terrain not initialised"); return 1; }
    if (n_sun <= 0) return 0;
    const size_t nc = (size_t)t->tp.dim_in_0 * t->tp.dim_in_1;
    HZB_CUDA(cudaSetDevice(t->s.device));
    const int nbuf = n_sun > 1 ? 2 : 1;
    HZB_TRY(terrain_out(t, nc * sizeof(T) * nbuf));
    HZB_TRY(t->streams());
    cudaStream_t s_comp = t->s_comp, s_copy = t->s_copy;
    cudaEvent_t* done = t->done; cudaEvent_t* freed = t->freed;
    int rc = 0;
    for (int k = 0; k <= n_sun && rc == 0; ++k) {
        if (k < n_sun) {
            const int b = k % nbuf;
            if (k >= nbuf) cudaStreamWaitEvent(s_comp, freed[b], 0);       // buffer b was copied out
            rc = launch(t, suns + 3 * k, (T*)t->d_out + nc * b, s_comp);
            cudaEventRecord(done[b], s_comp);
        }
        if (k >= 1 && rc == 0) {
            const int b = (k - 1) % nbuf;
            cudaStreamWaitEvent(s_copy, done[b], 0);
            if (cudaMemcpyAsync(out + nc * (size_t)(k - 1), (T*)t->d_out + nc * b, nc * sizeof(T), cudaMemcpyDeviceToHost, s_copy) != cudaSuccess) rc = 1;
            cudaEventRecord(freed[b], s_copy);
            if (cudaStreamSynchronize(s_copy) != cudaSuccess) rc = 1;       // pageable destination: host-synchronous anyway
        }
    }
    cudaStreamSynchronize(s_comp);
    if (rc) { if (g_error.empty()) set_error("terrain batch failed"); cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) set_error(cudaGetErrorString(e)); return 1; }
    HZB_CUDA(cudaGetLastError());
    return 0;
}
extern "C" {

int hzb_terrain_shadow_batch(hzb_terrain* t, const float* suns, int n_sun, uint8_t* out) {
    return terrain_batch<uint8_t>(t, suns, n_sun, out, [](hzb_terrain* tt, const float* sun, uint8_t* d, cudaStream_t st) {
        return launch_shadow(tt->s, tt->tp, sun, d, st); });
}
int hzb_terrain_sw_dir_cor_batch(hzb_terrain* t, const float* suns, int n_sun, float* out) {
    return terrain_batch<float>(t, suns, n_sun, out, [](hzb_terrain* tt, const float* sun, float* d, cudaStream_t st) {
        return launch_sw_dir_cor(tt->s, tt->tp, sun, d, st); });
}
int hzb_terrain_shadow(hzb_terrain* t, const float* sun, uint8_t* out) { return hzb_terrain_shadow_batch(t, sun, 1, out); }
int hzb_terrain_sw_dir_cor(hzb_terrain* t, const float* sun, float* out) { return hzb_terrain_sw_dir_cor_batch(t, sun, 1, out); }
int hzb_terrain_stats(const hzb_terrain* t, hzb_stats* out) {
    if (!t || !t->ready || !out) { set_error("terrain not initialised"); return 1; }
    memset(out, 0, sizeof(*out));
    return read_counters(t->s, *out);
}

// ------------------------------------------------------ azimuthal integrals
int hzb_sky_view_factor_dev(const float* a, const float* h, const float* t, long long cells, int K, float* o, void* st) {
    return launch_svf(0, a, h, t, cells, K, o, (cudaStream_t)st);
}
int hzb_visible_sky_fraction_dev(const float* a, const float* h, const float* t, long long cells, int K, float* o, void* st) {
    return launch_svf(1, a, h, t, cells, K, o, (cudaStream_t)st);
}
int hzb_topographic_openness_dev(const float* a, const float* h, long long cells, int K, float* o, void* st) {
    return launch_svf(2, a, h, nullptr, cells, K, o, (cudaStream_t)st);
}
static int integral_host(int kind, const float* azim, const float* hori, const float* tilt, int ny, int nx, int K, float* out) {
    if (require_device()) return 1;
    if (ny < 0 || nx < 0 || K < 1) { set_error("invalid dimensions"); return 1; }
    const size_t nc = (size_t)ny * nx;
    if (nc == 0) return 0;
    // The horizon array (nc x K floats, 2 GB for cfg2) streams through two device chunks:
    // staged H2D of chunk i+1 overlaps the integral kernel of chunk i.
    const size_t chunk_cells = std::min(nc, std::max<size_t>(1, ((size_t)64 << 20) / ((size_t)K * sizeof(float))));
    DevBuf<float> d_a, d_t, d_o, d_h[2];
    HZB_TRY(d_a.upload(azim, K));
    if (kind != 2) HZB_TRY(d_t.upload(tilt, nc * 3));
    HZB_TRY(d_o.alloc(nc));
    HZB_TRY(d_h[0].alloc(chunk_cells * (size_t)K));
    if (chunk_cells < nc) HZB_TRY(d_h[1].alloc(chunk_cells * (size_t)K));
    cudaStream_t st = nullptr;
    HZB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } sguard{st};
    cudaEvent_t done[2] = {nullptr, nullptr};
    HZB_CUDA(cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming));
    HZB_CUDA(cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming));
    struct EvGuard { cudaEvent_t* e; ~EvGuard() { cudaEventDestroy(e[0]); cudaEventDestroy(e[1]); } } eguard{done};
    const bool in_pinned = host_is_pinned(hori);   // e.g. the array horizon_gridded returned: plain DMA
    int slot = 0;
    for (size_t c0 = 0; c0 < nc; c0 += chunk_cells, slot ^= 1) {
        const size_t cells = std::min(chunk_cells, nc - c0);
        if (c0 >= 2 * chunk_cells) HZB_CUDA(cudaEventSynchronize(done[slot]));   // the kernel two chunks back has read this buffer
        if (in_pinned) HZB_CUDA(cudaMemcpyAsync(d_h[slot].p, hori + c0 * (size_t)K, cells * (size_t)K * sizeof(float), cudaMemcpyHostToDevice, st));
        else HZB_TRY(staged_h2d(d_h[slot].p, hori + c0 * (size_t)K, cells * (size_t)K * sizeof(float), st));
        HZB_TRY(launch_svf(kind, d_a.p, d_h[slot].p, kind != 2 ? d_t.p + 3 * c0 : nullptr, (long long)cells, K, d_o.p + c0, st));
        HZB_CUDA(cudaEventRecord(done[slot], st));
    }
    HZB_CUDA(cudaStreamSynchronize(st));
    HZB_CUDA(cudaMemcpy(out, d_o.p, nc * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}
static int slope_host(int method, const float* x, const float* y, const float* z, const float* rot, int ny, int nx, int output_rot, float* out) {
    if (require_device()) return 1;
    if (ny < 0 || nx < 0) { set_error("invalid dimensions"); return 1; }
    const size_t n = (size_t)ny * nx;
    if (n == 0) return 0;
    DevBuf<float> dx, dy, dz, dr, d_o;
    HZB_TRY(dx.upload(x, n)); HZB_TRY(dy.upload(y, n)); HZB_TRY(dz.upload(z, n));
    if (rot) HZB_TRY(dr.upload(rot, n * 9));
    HZB_TRY(d_o.alloc(n * 3));
    HZB_TRY(launch_slope(method, dx.p, dy.p, dz.p, rot ? dr.p : nullptr, ny, nx, output_rot, d_o.p, nullptr));
    HZB_CUDA(cudaMemcpy(out, d_o.p, n * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}
int hzb_slope_plane_meth(const float* x, const float* y, const float* z, const float* rot, int ny, int nx, int output_rot, float* out) {
    return slope_host(0, x, y, z, rot, ny, nx, output_rot, out);
}
int hzb_slope_vector_meth(const float* x, const float* y, const float* z, const float* rot, int ny, int nx, int output_rot, float* out) {
    return slope_host(1, x, y, z, rot, ny, nx, output_rot, out);
}
int hzb_sky_view_factor(const float* azim, const float* hori, const float* tilt, int ny, int nx, int K, float* out) {
    return integral_host(0, azim, hori, tilt, ny, nx, K, out);
}
int hzb_visible_sky_fraction(const float* azim, const float* hori, const float* tilt, int ny, int nx, int K, float* out) {
    return integral_host(1, azim, hori, tilt, ny, nx, K, out);
}
int hzb_topographic_openness(const float* azim, const float* hori, int ny, int nx, int K, float* out) {
    return integral_host(2, azim, hori, nullptr, ny, nx, K, out);
}


// ---- coordinate preparation, host tier (transform.pyx / direction.pyx loops)
int hzb_lonlat2ecef(const double* lon, const double* lat, const float* h, long long n, const char* ellps,
                    double* x_ecef, double* y_ecef, double* z_ecef) {
    if (require_device()) return 1;
    const int el = parse_ellps(ellps);
    if (el < 0) { set_error("Unknown value for 'ellps'"); return 1; }
    if (n <= 0) return 0;
    DevBuf<double> a, b, x, y, z; DevBuf<float> c;
    HZB_TRY(up_big(a, lon, (size_t)n)); HZB_TRY(up_big(b, lat, (size_t)n)); HZB_TRY(up_big(c, h, (size_t)n));
    HZB_TRY(x.alloc((size_t)n)); HZB_TRY(y.alloc((size_t)n)); HZB_TRY(z.alloc((size_t)n));
    HZB_TRY(launch_lonlat2ecef(el, a.p, b.p, c.p, n, x.p, y.p, z.p, nullptr));
    HZB_TRY(down_big(x_ecef, x, (size_t)n)); HZB_TRY(down_big(y_ecef, y, (size_t)n)); HZB_TRY(down_big(z_ecef, z, (size_t)n));
    return 0;
}
int hzb_ecef2enu(const double* x_ecef, const double* y_ecef, const double* z_ecef, long long n, double x_ecef_or,
                 double y_ecef_or, double z_ecef_or, double lon_or, double lat_or, float* x_enu, float* y_enu, float* z_enu) {
    if (require_device()) return 1;
    if (n <= 0) return 0;
    DevBuf<double> a, b, c; DevBuf<float> x, y, z;
    HZB_TRY(up_big(a, x_ecef, (size_t)n)); HZB_TRY(up_big(b, y_ecef, (size_t)n)); HZB_TRY(up_big(c, z_ecef, (size_t)n));
    HZB_TRY(x.alloc((size_t)n)); HZB_TRY(y.alloc((size_t)n)); HZB_TRY(z.alloc((size_t)n));
    HZB_TRY(launch_ecef2enu(a.p, b.p, c.p, n, x_ecef_or, y_ecef_or, z_ecef_or, lon_or, lat_or, x.p, y.p, z.p, nullptr));
    HZB_TRY(down_big(x_enu, x, (size_t)n)); HZB_TRY(down_big(y_enu, y, (size_t)n)); HZB_TRY(down_big(z_enu, z, (size_t)n));
    return 0;
}
int hzb_ecef2enu_vector(const float* vec_ecef, long long n, double lon_or, double lat_or, float* vec_enu) {
    if (require_device()) return 1;
    if (n <= 0) return 0;
    DevBuf<float> v, o;
    HZB_TRY(up_big(v, vec_ecef, (size_t)n * 3)); HZB_TRY(o.alloc((size_t)n * 3));
    HZB_TRY(launch_ecef2enu_vector(v.p, n, lon_or, lat_or, o.p, nullptr));
    return down_big(vec_enu, o, (size_t)n * 3);
}
int hzb_surf_norm(const double* lon, const double* lat, long long n, float* vec_norm_ecef) {
    if (require_device()) return 1;
    if (n <= 0) return 0;
    DevBuf<double> a, b; DevBuf<float> o;
    HZB_TRY(up_big(a, lon, (size_t)n)); HZB_TRY(up_big(b, lat, (size_t)n)); HZB_TRY(o.alloc((size_t)n * 3));
    HZB_TRY(launch_surf_norm(a.p, b.p, n, o.p, nullptr));
    return down_big(vec_norm_ecef, o, (size_t)n * 3);
}
int hzb_north_dir(const double* x_ecef, const double* y_ecef, const double* z_ecef, const float* vec_norm_ecef, long long n,
                  const char* ellps, float* vec_north_ecef) {
    if (require_device()) return 1;
    const int el = parse_ellps(ellps);
    if (el < 0) { set_error("Unknown value for 'ellps'"); return 1; }
    if (n <= 0) return 0;
    DevBuf<double> a, b, c; DevBuf<float> v, o;
    HZB_TRY(up_big(a, x_ecef, (size_t)n)); HZB_TRY(up_big(b, y_ecef, (size_t)n)); HZB_TRY(up_big(c, z_ecef, (size_t)n));
    HZB_TRY(up_big(v, vec_norm_ecef, (size_t)n * 3)); HZB_TRY(o.alloc((size_t)n * 3));
    HZB_TRY(launch_north_dir(el, a.p, b.p, c.p, v.p, n, o.p, nullptr));
    return down_big(vec_north_ecef, o, (size_t)n * 3);
}
int hzb_wgs2swiss(const double* lon, const double* lat, const float* h_wgs, long long n, double* e, double* nn, float* h_ch) {
    if (require_device()) return 1;
    if (n <= 0) return 0;
    DevBuf<double> a, b, x, y; DevBuf<float> c, z;
    HZB_TRY(up_big(a, lon, (size_t)n)); HZB_TRY(up_big(b, lat, (size_t)n)); HZB_TRY(up_big(c, h_wgs, (size_t)n));
    HZB_TRY(x.alloc((size_t)n)); HZB_TRY(y.alloc((size_t)n)); HZB_TRY(z.alloc((size_t)n));
    HZB_TRY(launch_wgs2swiss(a.p, b.p, c.p, n, x.p, y.p, z.p, nullptr));
    HZB_TRY(down_big(e, x, (size_t)n)); HZB_TRY(down_big(nn, y, (size_t)n)); HZB_TRY(down_big(h_ch, z, (size_t)n));
    return 0;
}
int hzb_swiss2wgs(const double* e, const double* nn, const float* h_ch, long long n, double* lon, double* lat, float* h_wgs) {
    if (require_device()) return 1;
    if (n <= 0) return 0;
    DevBuf<double> a, b, x, y; DevBuf<float> c, z;
    HZB_TRY(up_big(a, e, (size_t)n)); HZB_TRY(up_big(b, nn, (size_t)n)); HZB_TRY(up_big(c, h_ch, (size_t)n));
    HZB_TRY(x.alloc((size_t)n)); HZB_TRY(y.alloc((size_t)n)); HZB_TRY(z.alloc((size_t)n));
    HZB_TRY(launch_swiss2wgs(a.p, b.p, c.p, n, x.p, y.p, z.p, nullptr));
    HZB_TRY(down_big(lon, x, (size_t)n)); HZB_TRY(down_big(lat, y, (size_t)n)); HZB_TRY(down_big(h_wgs, z, (size_t)n));
    return 0;
}
int hzb_rotation_matrix_glob2loc(const float* vec_north_enu, const float* vec_norm_enu, int ny, int nx, float* rot_mat) {
    if (require_device()) return 1;
    if (ny < 0 || nx < 0) { set_error("invalid dimensions"); return 1; }
    const size_t nc = (size_t)ny * nx, no = (size_t)(ny + 2) * (nx + 2) * 9;
    DevBuf<float> a, b, o;
    HZB_TRY(up_big(a, vec_north_enu, nc * 3)); HZB_TRY(up_big(b, vec_norm_enu, nc * 3)); HZB_TRY(o.alloc(no));
    HZB_TRY(launch_rotmat(a.p, b.p, ny, nx, o.p, nullptr));
    return down_big(rot_mat, o, no);
}
// additive, resident tier: the whole preparation chain fused, inputs and outputs in HBM
int hzb_prep_enu_dev(const double* d_lon, const double* d_lat, const float* d_elev, int ny, int nx, const char* ellps,
                     double x_ecef_or, double y_ecef_or, double z_ecef_or, double lon_or, double lat_or, int offset_0,
                     int offset_1, int dim_in_0, int dim_in_1, float* d_vert_grid, float* d_vec_norm, float* d_vec_north,
                     void* stream) {
    const int el = parse_ellps(ellps);
    if (el < 0) { set_error("Unknown value for 'ellps'"); return 1; }
    if (!d_lon || !d_lat || !d_elev || !d_vert_grid) { set_error("null pointer argument"); return 1; }
    if ((d_vec_norm == nullptr) != (d_vec_north == nullptr)) { set_error("vec_norm and vec_north go together"); return 1; }
    return launch_prep_enu(el, d_lon, d_lat, d_elev, ny, nx, x_ecef_or, y_ecef_or, z_ecef_or, lon_or, lat_or, offset_0, offset_1,
                           dim_in_0, dim_in_1, d_vert_grid, d_vec_norm, d_vec_north, (cudaStream_t)stream);
}
}  // extern "C"
