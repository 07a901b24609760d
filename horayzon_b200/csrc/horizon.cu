// horizon.cu -- horizon search kernels (B200, sm_100a).
//
// Replaces the TBB row loop and the three search algorithms of the reference
// (horizon_comp.cpp:302-498, 519-612, 739-800, 926-1070).  One lane owns one
// grid cell and walks its azimuth chain; warps pull 8x4-cell tiles from an
// atomic queue (persistent CTAs, grid = SM count x resident CTAs).  All
// arithmetic that decides a table index or a ray direction is written with
// explicitly rounded intrinsics in the reference's float/double mix, so the
// outputs are decision-exact against the CPU oracle.
//
//   k_horizon_wq6      production: search state machine + two-ray packet warp-queue
//                      traversal of the compressed 4-wide BVH (hzb_wq2.cuh)
//                      ("Queue order and azimuth segments" below: band first, the last tiles' chains split)
//   k_horizon_redo     fix-up pass behind every production launch: cells whose traversal stack was full,
//                      azimuth segments that started from a wrong chain index (normally neither)
//   k_horizon_gridded  reference-shaped per-lane kernel on the binary BVH: the second,
//                      structurally different implementation the full-size parity test
//                      compares the production kernel with (hzb_debug_option, test only)
//   k_loc_wq           arbitrary locations on the same traversal step (any-hit / closest-hit), one lane per
//                      (location, azimuth); k_loc_snap / k_loc_chain / k_loc_indep: per-lane binary-BVH forms
//                      (surface snap; second implementation for the parity tests)
#include "hzb_geom.cuh"
#include "hzb_wq.cuh"
#include "hzb_wq2.cuh"
#include "hzb_search.cuh"
#include "hzb_queue.cuh"
#include <math.h>
#include <string.h>
#include <algorithm>
#include <mutex>
#include <vector>

namespace hzb {

// ------------------------------------------------------------ host tables
static inline float deg2rad_f(float a) { return (float)(((double)a / 180.0) * M_PI); }  // horizon_comp.cpp:37-39

void HorizonTables::make(int azim_n, float dist_km, float acc_deg, float low_deg, bool fill) {
    azim_num = azim_n;
    acc = deg2rad_f(acc_deg);
    low = deg2rad_f(low_deg);
    up = deg2rad_f(89.98f);                      // horizon_comp.cpp:648
    dist = (float)((double)dist_km * 1000.0);    // :670
    step = (double)acc / 5.0;
    elev_num = (int)ceil((double)(up - low) / step) + 1;  // :721-722
    if (!fill) return;                           // scalars only (the arrays are cached on the device)
    azim_sin.resize(azim_n); azim_cos.resize(azim_n);
    for (int i = 0; i < azim_n; ++i) {           // :714-718: float angle, float sin/cos overloads
        const float ang = (float)((2 * M_PI) / azim_n * i);
        azim_sin[i] = sinf(ang); azim_cos[i] = cosf(ang);
    }
    elev_ang.resize(elev_num); elev_sin.resize(elev_num); elev_cos.resize(elev_num);
    for (int i = 0; i < elev_num; ++i) {         // :726-731: anchored at the upper limit
        const float ang = (float)((double)up - step * i);
        elev_ang[elev_num - i - 1] = ang;
        elev_sin[elev_num - i - 1] = sinf(ang);
        elev_cos[elev_num - i - 1] = cosf(ang);
    }
}

// Device tables of one parameter set: looked up in the scene's cache, uploaded once.
int scene_tables(Scene& s, int azim_num, float dist_km, float acc_deg, float low_deg, HorizonParams& p, cudaStream_t st) {
    HorizonTables T;
    const Scene::TableEntry* hit = nullptr;
    for (const Scene::TableEntry& e : s.tables)
        if (e.azim_num == azim_num && e.acc_deg == acc_deg && e.low_deg == low_deg && e.dist_km == dist_km) { hit = &e; break; }
    T.make(azim_num, dist_km, acc_deg, low_deg, hit == nullptr);
    if (T.elev_num < 2) { set_error("elevation table too small"); return 1; }
    if (!hit) {
        if (s.tables.size() >= 32) {   // bounded cache: drop everything once nothing can be in flight any more
            HZB_CUDA(cudaDeviceSynchronize());
            for (Scene::TableEntry& e : s.tables) pool_free(e.d);
            s.tables.clear();
        }
        const size_t need = (size_t)2 * T.azim_num + (size_t)3 * T.elev_num;
        std::vector<float> h(need);
        float* q = h.data();
        memcpy(q, T.azim_sin.data(), 4 * (size_t)T.azim_num); q += T.azim_num;
        memcpy(q, T.azim_cos.data(), 4 * (size_t)T.azim_num); q += T.azim_num;
        memcpy(q, T.elev_ang.data(), 4 * (size_t)T.elev_num); q += T.elev_num;
        memcpy(q, T.elev_sin.data(), 4 * (size_t)T.elev_num); q += T.elev_num;
        memcpy(q, T.elev_cos.data(), 4 * (size_t)T.elev_num);
        Scene::TableEntry e{azim_num, acc_deg, low_deg, dist_km, nullptr, T.elev_num};
        e.d = (float*)pool_alloc(need * sizeof(float));
        if (!e.d) return 1;
        // pageable source: the call returns once the data has left `h` (first use of a parameter set only)
        if (cudaMemcpyAsync(e.d, h.data(), need * sizeof(float), cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) { pool_free(e.d); set_error("table upload failed"); cudaGetLastError(); return 1; }
        s.tables.push_back(e);
        hit = &s.tables.back();
    }
    p.azim_sin = hit->d; p.azim_cos = p.azim_sin + azim_num;
    p.elev_ang = p.azim_cos + azim_num; p.elev_sin = p.elev_ang + T.elev_num; p.elev_cos = p.elev_sin + T.elev_num;
    p.azim_num = azim_num; p.elev_num = T.elev_num;
    p.acc = T.acc; p.low = T.low; p.up = T.up; p.dist = T.dist; p.step = T.step;
    return 0;
}

unsigned int* scene_tile_counter(Scene& s, cudaStream_t st) {
    unsigned int* c = s.d_tile_counter + (s.tile_slot++ % HZB_TILE_SLOTS);
    if (cudaMemsetAsync(c, 0, sizeof(unsigned int), st) != cudaSuccess) { set_error("work-queue reset failed"); cudaGetLastError(); return nullptr; }
    return c;
}

// ------------------------------------------------------------ device side
namespace {

struct Search : SearchTables {  // everything a lane needs to cast one ray of the search (tables: hzb_search.cuh)
    SceneView sv;
    unsigned int* overflow;
};

__device__ __forceinline__ Search make_search(const SceneView& sv, const HorizonParams& p, Counters* c) {
    Search s;
    s.sv = sv; s.azim_sin = p.azim_sin; s.azim_cos = p.azim_cos; s.elev_ang = p.elev_ang;
    s.elev_sin = p.elev_sin; s.elev_cos = p.elev_cos; s.azim_num = p.azim_num; s.elev_num = p.elev_num;
    s.acc = p.acc; s.low = p.low; s.up = p.up; s.dist = p.dist; s.step = p.step;
    s.overflow = reinterpret_cast<unsigned int*>(&c->stack_overflow);
    return s;
}

// any-hit cast (castRay_occluded1, horizon_comp.cpp:241-262)
__device__ __forceinline__ bool cast_any(const Search& s, const Frame& f, int ie, int k, LaneCounters& cnt) {
    cnt.rays++;
    float tfar = s.dist;
    return trace_bvh2<false>(s.sv, f.org, ray_dir(s, f, ie, k), tfar, cnt, s.overflow);
}
// closest-hit cast (castRay_intersect1, :268-292): dist = distance of the hit
__device__ __forceinline__ bool cast_closest(const Search& s, const Frame& f, int ie, int k, LaneCounters& cnt, float& dist) {
    cnt.rays++;
    float tfar = s.dist;
    const bool hit = trace_bvh2<true>(s.sv, f.org, ray_dir(s, f, ie, k), tfar, cnt, s.overflow);
    dist = tfar;
    return hit;
}

// bisection for one azimuth (horizon_comp.cpp:348-376); returns the final index
template <bool WD>
__device__ __forceinline__ int bisect(const Search& s, const Frame& f, int k, LaneCounters& cnt, float& mid, float& dist_hit) {
    float lim_up = s.up, lim_low = s.low;
    float samp = midpoint(lim_up, lim_low);
    int ie = index_of(s, samp);
    while (true) {
        const float ea = __ldg(s.elev_ang + ie);
        if (!(fmaxf(__fsub_rn(lim_up, ea), __fsub_rn(ea, lim_low)) > s.acc)) break;
        bool hit;
        if (WD) { float d; hit = cast_closest(s, f, ie, k, cnt, d); if (hit) dist_hit = d; }
        else hit = cast_any(s, f, ie, k, cnt);
        if (hit) lim_low = ea; else lim_up = ea;
        samp = midpoint(lim_up, lim_low);
        ie = index_of(s, samp);
    }
    mid = samp;
    return ie;
}

// output staging: four consecutive azimuths per 16-byte store when aligned
struct OutBuf {
    float* out; bool vec, writer; float b0, b1, b2, b3; long long sk;
    __device__ __forceinline__ void init(float* o, bool v, bool w = true, long long stride_k = 1) {
        out = o; vec = v && stride_k == 1; writer = w; b0 = b1 = b2 = b3 = 0.f; sk = stride_k;
    }
    __device__ __forceinline__ void put(int k, float v) {
        if (!vec) { if (writer) out[k * sk] = v; return; }
        b0 = b1; b1 = b2; b2 = b3; b3 = v;
        if ((k & 3) == 3 && writer) *reinterpret_cast<float4*>(out + (k - 3)) = make_float4(b0, b1, b2, b3);
    }
    __device__ __forceinline__ void put_idx(int k, int, float v) { put(k, v); }
    __device__ __forceinline__ void mark_redo() { out[0] = __uint_as_float(HZB_REDO_F32); }
    __device__ __forceinline__ void fill(int K, float v) { for (int k = 0; k < K; ++k) out[k * sk] = v; }
};

// Quantised output (scope row 8f-4): every guess_constant result but the first azimuth's IS a table entry
// (horizon_comp.cpp:490-494), so the kernel stores its 16-bit table index; the first azimuth's un-quantised
// bisection midpoint (:428) goes to a float per cell.  Lossless: elev_ang[index] is the float the other mode stores.
constexpr unsigned short HZB_Q_FIRST = 0xFFFFu;   // "take the cell's float" (azimuth 0, masked cells)
struct OutBufQ {
    unsigned short* idx; float* first; long long sk;
    __device__ __forceinline__ void init(unsigned short* q, float* f, long long stride_k) { idx = q; first = f; sk = stride_k; }
    __device__ __forceinline__ void put(int k, float v) { *first = v; idx[k * sk] = HZB_Q_FIRST; }     // k == 0 only (guess_constant)
    __device__ __forceinline__ void put_idx(int k, int ie, float) { idx[k * sk] = (unsigned short)ie; }
    __device__ __forceinline__ void mark_redo() { *first = __uint_as_float(HZB_REDO_F32); }
    __device__ __forceinline__ void fill(int K, float v) { *first = v; for (int k = 0; k < K; ++k) idx[k * sk] = HZB_Q_FIRST; }
};
template <bool Q> struct OutSel { typedef OutBuf type; };
template <> struct OutSel<true> { typedef OutBufQ type; };
template <bool Q>
__device__ __forceinline__ void out_init(typename OutSel<Q>::type& ob, const HorizonParams& p, size_t slot, size_t cell);
template <>
__device__ __forceinline__ void out_init<false>(OutBuf& ob, const HorizonParams& p, size_t slot, size_t) {
    ob.init(p.hori + slot * p.stride_c, false, true, p.stride_k);   // one 4-byte store per azimuth: L2 merges them long before the sector is evicted
}
template <>
__device__ __forceinline__ void out_init<true>(OutBufQ& ob, const HorizonParams& p, size_t slot, size_t) {
    ob.init(p.hori_q + slot * p.stride_c, p.hori_first + slot, p.stride_k);
}

// One cell, the azimuths [k_begin, k_end) (all of them by default).  ALG 0 discrete_sampling (:302-333), 1 binary_search
// (:339-381), 2 guess_constant (:387-498); a guess_constant range that does not start at azimuth 0 continues the
// chain from prev_az, the table index of azimuth k_begin - 1.  Termination rule: a hit
// at the top index counts as a miss, a miss at index 0 as a hit.
template <int ALG, typename OB>
__device__ void cell_search(const Search& s, const Frame& f, OB& ob, LaneCounters& cnt, int k_begin = 0, int k_end = -1, int prev_az = 0) {
    const int top = s.elev_num - 1;
    if (k_end < 0) k_end = s.azim_num;
    if (ALG == 0) {
        for (int k = k_begin; k < k_end; ++k) {
            int cur = 0, prev = 0; bool hit = true;
            while (hit) {
                prev = cur; cur = min(cur + 10, top);
                hit = cast_any(s, f, cur, k, cnt);
                if (cur == top) hit = false;
            }
            ob.put(k, midpoint(__ldg(s.elev_ang + prev), __ldg(s.elev_ang + cur)));
        }
    } else if (ALG == 1) {
        for (int k = k_begin; k < k_end; ++k) {
            float mid, dh = 0.f;
            bisect<false>(s, f, k, cnt, mid, dh);
            ob.put(k, mid);
        }
    } else {
        if (k_begin == 0) {
            float mid, dh = 0.f;
            prev_az = bisect<false>(s, f, 0, cnt, mid, dh);
            ob.put(0, mid);
            k_begin = 1;
        }
        for (int k = k_begin; k < k_end; ++k) {
            int cur = max(prev_az - 5, 0), prev = 0, count = 0; bool hit = true;
            while (hit) {
                prev = cur; cur = min(cur + 10, top);
                hit = cast_any(s, f, cur, k, cnt); ++count;
                if (cur == top) hit = false;
            }
            if (count <= 1) {
                cur = min(prev_az + 5, top); hit = false;
                while (!hit) {
                    prev = cur; cur = max(cur - 10, 0);
                    hit = cast_any(s, f, cur, k, cnt);
                    if (cur == 0) hit = true;
                }
            }
            const int ie = index_of(s, midpoint(__ldg(s.elev_ang + prev), __ldg(s.elev_ang + cur)));
            ob.put_idx(k, ie, __ldg(s.elev_ang + ie));
            prev_az = ie;
        }
    }
}

__device__ __forceinline__ void flush_counters(LaneCounters& cnt, unsigned int units, Counters* c) {
    unsigned int r = cnt.rays, n = cnt.nodes, p = cnt.prims, u = units;
    for (int o = 16; o > 0; o >>= 1) {
        r += __shfl_xor_sync(0xffffffffu, r, o); n += __shfl_xor_sync(0xffffffffu, n, o);
        p += __shfl_xor_sync(0xffffffffu, p, o); u += __shfl_xor_sync(0xffffffffu, u, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&c->rays, (unsigned long long)r); atomicAdd(&c->node_visits, (unsigned long long)n);
        atomicAdd(&c->prim_tests, (unsigned long long)p); atomicAdd(&c->units, (unsigned long long)u);
    }
    cnt.rays = cnt.nodes = cnt.prims = 0;
}

// A finished cell (or empty cell slot) of row block `blk`: its outputs are made visible system-wide, the
// block's counter goes up, and the lane that completes the block raises the block's flag in mapped host
// memory -- the host tier polls those flags and copies finished blocks while the kernel is still running.
__device__ __forceinline__ void publish_cell(const HorizonParams& p, int blk, unsigned int slots) {
    // device scope is enough for the cell's outputs: they are read by the copy engine (through L2) after the host has
    // seen the block's flag, and the flag is written behind a system-scope fence by the lane that saw every other
    // lane's (fenced) arrival
    __threadfence();
    if (atomicAdd(p.row_done + blk, 1u) == slots - 1u && p.row_flags) {
        __threadfence_system();
        p.row_flags[blk] = 1u;
    }
}

// 4-row blocks of this launch: those blocks b of the row range with b % blk_stride == blk_offset
__device__ __forceinline__ int local_blocks(const HorizonParams& p, int rows) {
    const int all = (rows + 3) >> 2;
    return all > p.blk_offset ? (all - p.blk_offset + p.blk_stride - 1) / p.blk_stride : 0;
}

// ---------------------------------------------------------------------------
// Queue order and azimuth segments.
//
// A cell's guess_constant search is a chain over the azimuths (azimuth k starts from the index of k-1), ~40 ms long on
// cfg2, and a launch ends when the last chain ends: with few cells per resident lane (the eighth of cfg2 one GPU of an
// 8-GPU run owns has 1.6) the lanes that happen to start a chain late set the time while the others idle.  The tail
// of the queue is therefore made of shorter tasks: the cells of the last q_tail tiles are split into SEG_COUNT
// azimuth SEGMENTS, each a queue entry of its own.  Segment 0 is the head of the chain; a later segment does not know
// the chain's index at its first azimuth and starts with the prelude of hzb_search.cuh, which finds it from the
// residue class the chain keeps (one bisection over the rungs of that class, ~8 single-ray casts; the class comes
// from the bisection of azimuth 0, which the lane that owns the head publishes -- a segment that starts before that
// repeats it, ~7 casts).  That value is an assumption -- it is wrong when the
// chain ran into the lower end of the elevation table on its way (horizon below the table's low limit, i.e. cells
// that look out over the DEM's edge: the index is clamped there and the residue changes) -- so every segment
// records it, and the fix-up kernel (k_horizon_redo, behind every launch) compares it with the index the preceding
// segment really produced; a segment that started from anything else is recomputed there.  The outputs are the
// sequential chain's in every case; the cast counters too (a recomputed segment's first count is taken back).
// To keep recomputation out of the picture, the tiles within the BAND along the DEM's edge where such clamping
// can happen (host: relief / tan(-low limit)) are never split; they are queued first, as whole chains.
// The azimuths of discrete_sampling / binary_search are independent: their segments need no prelude and no check.
//
// Queue: [band tiles: top block rows, bottom block rows, left / right tile columns of the rows between]
//        [interior tiles in row order, whole chains] [the last q_tail interior tiles x SEG_COUNT segments, segment-major]
// (rows still complete in order for the host tier's overlapped copy: the interior is the last part of every row).
// ---------------------------------------------------------------------------
__device__ __forceinline__ SegRecord* seg_record(const HorizonParams& p, int ci, int cj, int n) { return p.seg + seg_record_index(p, ci, cj, n); }

constexpr int HG_THREADS = 128;

template <int ALG>
__global__ void __launch_bounds__(HG_THREADS) k_horizon_gridded(SceneView sv, HorizonParams p, Counters* counters,
                                                                unsigned int* tile_counter) {
    const Search s = make_search(sv, p, counters);
    const int lane = threadIdx.x & 31;
    const int rows = p.row_end - p.row_begin;
    const int tiles_x = (p.dim_in_1 + 7) >> 3, tiles_y = local_blocks(p, rows);
    const unsigned int num_tiles = (unsigned int)tiles_x * (unsigned int)tiles_y;
    const bool vec = (p.azim_num & 3) == 0 && ((reinterpret_cast<size_t>(p.hori) & 15) == 0);
    LaneCounters cnt; cnt.rays = cnt.nodes = cnt.prims = 0;
    while (true) {
        unsigned int tile = 0;
        if (lane == 0) tile = atomicAdd(tile_counter, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= num_tiles) break;
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const int i = p.row_begin + (ty * p.blk_stride + p.blk_offset) * 4 + (lane >> 3), j = tx * 8 + (lane & 7);
        unsigned int units = 0;
        if (i < p.row_end && j < p.dim_in_1) {
            const size_t c = (size_t)i * p.dim_in_1 + j;
            float* out = p.hori + (p.packed ? (size_t)(ty * 4 + (lane >> 3)) * p.dim_in_1 + j : c) * p.stride_c;
            if (p.mask[c] == 1) {
                const F3 nrm = f3(p.vec_norm[3 * c], p.vec_norm[3 * c + 1], p.vec_norm[3 * c + 2]);
                const F3 nth = f3(p.vec_north[3 * c], p.vec_north[3 * c + 1], p.vec_north[3 * c + 2]);
                const float4 v = sv.vert4[(size_t)(i + p.offset_0) * sv.W + (j + p.offset_1)];
                const Frame f = make_frame(f3(v.x, v.y, v.z), nrm, nth, p.ray_org_elev);
                OutBuf ob; ob.init(out, vec, true, p.stride_k);
                cell_search<ALG>(s, f, ob, cnt);
                units = p.azim_num;
            } else {
                for (int k = 0; k < p.azim_num; ++k) out[k * p.stride_k] = p.hori_fill;  // horizon_comp.cpp:789-794
            }
        }
        flush_counters(cnt, units, counters);
    }
}


// Fix-up kernel, launched right behind every production launch on the same stream; one strided 4-byte read per
// cell (+ three records per split cell) when there is nothing to do, which is the rule.
//  * Cells the production kernel marked with HZB_REDO_F32 (its traversal stack was full) are recomputed with the
//    per-lane search on the binary BVH.
//  * Split cells ("Queue order and azimuth segments"): segment n >= 1 is valid if it started from the table index the
//    preceding segment ended with; otherwise (or after a full stack) its azimuths are recomputed here, in order, so
//    that the next segment is checked against the corrected value.
// Rewritten row blocks are flagged 2 for the host tier, which copies them again.
template <int ALG, bool Q>
__global__ void __launch_bounds__(HG_THREADS) k_horizon_redo(SceneView sv, HorizonParams p, Counters* counters) {
    const Search s = make_search(sv, p, counters);
    const int rows = p.row_end - p.row_begin;
    const long long slots = (long long)local_blocks(p, rows) * 4 * p.dim_in_1;
    LaneCounters cnt; cnt.rays = cnt.nodes = cnt.prims = 0;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < slots; g += (long long)gridDim.x * blockDim.x) {
        const int lr = (int)(g / p.dim_in_1), j = (int)(g - (long long)lr * p.dim_in_1);    // local (packed) row, column
        const int i = p.row_begin + ((lr >> 2) * p.blk_stride + p.blk_offset) * 4 + (lr & 3);
        if (i >= p.row_end) continue;
        const size_t c = (size_t)i * p.dim_in_1 + j;
        const size_t slot = p.packed ? (size_t)lr * p.dim_in_1 + j : c;
        const bool marked = __float_as_uint(Q ? p.hori_first[slot] : p.hori[slot * p.stride_c]) == HZB_REDO_F32;
        const int tt = tail_tile(p, lr >> 2, j >> 3);
        if (!marked && tt < 0) continue;
        if (p.mask[c] != 1) continue;
        bool rewritten = false;
        typename OutSel<Q>::type ob; out_init<Q>(ob, p, slot, c);
        Frame f; bool have_frame = false;
        auto frame = [&]() {
            if (have_frame) return;
            const F3 nrm = f3(p.vec_norm[3 * c], p.vec_norm[3 * c + 1], p.vec_norm[3 * c + 2]);
            const F3 nth = f3(p.vec_north[3 * c], p.vec_north[3 * c + 1], p.vec_north[3 * c + 2]);
            const float4 v = sv.vert4[(size_t)(i + p.offset_0) * sv.W + (j + p.offset_1)];
            f = make_frame(f3(v.x, v.y, v.z), nrm, nth, p.ray_org_elev);
            have_frame = true;
        };
        if (marked) { frame(); cell_search<ALG>(s, f, ob, cnt); rewritten = true; }
        else {
            for (int n = 1; n < SEG_COUNT; ++n) {
                const SegRecord rec = *seg_record(p, i, j, n);
                const int k0 = seg_begin(p, n), k1 = seg_begin(p, n + 1);
                bool redo = rec.guess == SEG_REDO || rec.guess == SEG_NONE;
                int prev_az = 0;
                if (ALG == 2) {   // the chain's index at azimuth k0 - 1 (a table entry: k0 >= 2)
                    if (Q) prev_az = (int)p.hori_q[slot * p.stride_c + (long long)(k0 - 1) * p.stride_k];
                    else {
                        const float v = p.hori[slot * p.stride_c + (long long)(k0 - 1) * p.stride_k];
                        prev_az = min(max(index_of(s, v), 1), s.elev_num - 2);
                        if (__ldg(s.elev_ang + prev_az) != v) prev_az += (__ldg(s.elev_ang + prev_az + 1) == v) ? 1 : -1;
                    }
                    if (rec.guess != prev_az) redo = true;
                }
                if (!redo) continue;
                frame();
                if (rec.guess >= 0) cnt.rays -= rec.casts;      // the task ran to its end from the wrong index: its count is taken back
                cell_search<ALG>(s, f, ob, cnt, k0, k1, prev_az);
                atomicAdd(&counters->segment_redos, 1ull);
                rewritten = true;
            }
        }
        if (rewritten && p.row_flags) p.row_flags[(i - p.row_begin) >> 2] = 2u;
    }
    if (cnt.rays | cnt.nodes | cnt.prims) {
        atomicAdd(&counters->rays, (unsigned long long)(long long)(int)cnt.rays);      // (may be negative: sign-extended, wraps correctly)
        atomicAdd(&counters->node_visits, (unsigned long long)cnt.nodes); atomicAdd(&counters->prim_tests, (unsigned long long)cnt.prims);
    }
}

// ===========================================================================
// k_horizon_wq6: production kernel.  Persistent warps pull 8x4-cell tiles from an atomic
// queue; every lane runs the search state machine of hzb_search.cuh for its cell on the
// two-ray packet traversal of hzb_wq2.cuh: the casts at prev+5 and prev-5 of a
// guess_constant azimuth share one traversal.
// ===========================================================================
template <int ALG, int MINB, bool Q>
__global__ void __launch_bounds__(WQ_BLOCK, MINB) k_horizon_wq6(SceneView sv, HorizonParams p, Counters* counters,
                                                                unsigned int* tile_counter, int refill_thr, int wait_thr,
                                                                int stack_lim) {
    __shared__ Wq2Shared sh;
    const Search s = make_search(sv, p, counters);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
    const unsigned int FULL = 0xffffffffu, lt_mask = (1u << lane) - 1u;
    const int tiles_x = p.q_tiles_x;
    LaneCounters cnt; cnt.rays = cnt.nodes = cnt.prims = 0;
    if (lane == 0) { sh.hit1[warp] = 0u; sh.hit2[warp] = 0u; }
    unsigned int pend_est = 0;
    __syncwarp();

    unsigned int cur_tile = 0; int next_cell = 32; bool more_tiles = true;   // cur_tile: tile index | task << 28 (0 whole chains, 1 + n: azimuth segment n)
    LaneSM m; m.phase = 0; m.k = 0; m.cur = m.prev = m.count = m.prev_az = 0;
    m.spec_ie = -1; m.spec_hit = false;
    typedef typename OutSel<Q>::type OB;
    OB ob;
    unsigned int my_cell = 0;   // (row << 16) | column of the lane's cell (+ segment number, see cell_seg): its frame is rebuilt at every ray set-up
    bool has_cell = false, have_result = false;
    Wq2Lane L; L.state = 0; L.hit1 = L.hit2 = false; L.node = WQ_NONE; L.sp = 0; L.pc = 0;
    L.A1x = L.A1y = L.A1z = L.B1x = L.B1y = L.B1z = 0.f; L.A2x = L.A2y = L.A2z = L.B2x = L.B2y = L.B2z = 0.f;
    L.selxy = 0x74107410u;

    while (true) {
        // (A) hand out cells to lanes that have none
        while (true) {
            const bool want = !has_cell;
            const unsigned int wmask = __ballot_sync(FULL, want);
            if (wmask == 0u) break;
            if (next_cell >= 32) {
                if (!more_tiles) break;
                unsigned int t = 0;
                if (lane == 0) t = atomicAdd(tile_counter, 1u);
                t = __shfl_sync(FULL, t, 0);
                if (t >= p.q_total) { more_tiles = false; break; }
                int qy, qx, task;
                queue_decode(p, t, qy, qx, task);
                cur_tile = (unsigned int)(qy * tiles_x + qx) | ((unsigned int)task << 28); next_cell = 0;   // (< 2^25 tiles: dims <= 32767)
            }
            const int mine = next_cell + __popc(wmask & lt_mask);
            next_cell += __popc(wmask);
            if (want && mine < 32) {
                const int cur_seg = (int)(cur_tile >> 28), tile = (int)(cur_tile & 0x0FFFFFFFu);
                const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
                const int ci = p.row_begin + (ty * p.blk_stride + p.blk_offset) * 4 + (mine >> 3), cj = tx * 8 + (mine & 7);
                bool done_now = true;
                if (ci < p.row_end && cj < p.dim_in_1) {
                    const size_t c = (size_t)ci * p.dim_in_1 + cj;
                    const size_t slot = p.packed ? (size_t)(ty * 4 + (mine >> 3)) * p.dim_in_1 + cj : c;
                    if (p.mask[c] == 1) {
                        my_cell = ((unsigned int)ci << 16) | (unsigned int)cj;    // dims <= 32767 (horizon.pyx:149-151)
                        int k0 = 0, k1 = p.azim_num;
                        if (cur_seg > 0) {          // azimuth segment n of a split cell
                            const int n = cur_seg - 1;
                            my_cell |= ((unsigned int)(n & 1) << 15) | ((unsigned int)(n & 2) << 30);
                            k0 = seg_begin(p, n); k1 = seg_begin(p, n + 1);
                            if (n > 0) {
                                SegRecord* r = seg_record(p, ci, cj, n);
                                r->guess = (ALG == 2) ? SEG_NONE : SEG_OK;
                                r->casts = 0u - cnt.rays;          // + the lane's count at the end of the task = the task's casts
                                atomicAdd(&counters->segment_tasks, 1ull);
                            }
                        }
                        out_init<Q>(ob, p, slot, c);
                        m.phase = 0; m.spec_ie = -1;
                        m.k = (ALG == 2) ? 0 : k0;      // guess_constant: azimuth 0, or the prelude of segment n >= 1 (sm_advance)
                        has_cell = true; have_result = false;
                        atomicAdd(&counters->units, (unsigned long long)(k1 - k0));      // (once per cell: no per-lane register for it)
                        done_now = false;
                    } else if (cur_seg <= 1) {
                        OB fo; out_init<Q>(fo, p, slot, c);
                        fo.fill(p.azim_num, p.hori_fill);  // horizon_comp.cpp:789-794
                    }
                }
                if (done_now && p.row_done) publish_cell(p, ty, row_slots(p, ty));
            }
        }
        // (B) lanes with a cell but no packet in flight advance their search; the ray set-up
        //     runs after the lanes have reconverged (sm_advance leaves through many exits)
        bool finished_cell = false, need_ray = false;
        int ie = 0, lo_ie = -1;
        if (has_cell && L.state == 0) {
            unsigned int extra = 0;
            const int seg_n = cell_seg(my_cell);
            // end of the lane's azimuths: the segment's end, or azim_num for a whole chain (the cell word has no room for
            // "segment 0 of a split cell": that is read off the cell's place in the queue)
            const bool split = seg_n > 0 || cell_is_split(p, cell_row(my_cell), cell_col(my_cell));
            const int k_end = split ? seg_begin(p, seg_n + 1) : p.azim_num;
            m.spec_hit = L.hit2;
            if (L.node == WQ_OVF) {   // the packet's stack was full: the cell (or segment) is left to the fix-up kernel (same results)
                L.node = WQ_NONE;
                if (seg_n > 0) seg_record(p, cell_row(my_cell), cell_col(my_cell), seg_n)->guess = SEG_REDO;
                else ob.mark_redo();
                atomicAdd(&counters->fallback_packets, 1ull);
                need_ray = false;
            } else {
                int seg_guess = -1, r0_known = -1;
                if (ALG == 2 && seg_n > 0 && !have_result && m.phase == 0)      // first step of the task: has the head of the chain reported yet?
                    r0_known = *(volatile int*)&seg_record(p, cell_row(my_cell), cell_col(my_cell), SEG_COUNT)->guess;
                need_ray = sm_advance<ALG, true, OB>(s, m, have_result, L.hit1, ob, ie, lo_ie, extra,
                                                     (ALG == 2 && seg_n > 0) ? seg_begin(p, seg_n) : 0, k_end, &seg_guess, r0_known);
                if (seg_guess >= 0 && split) seg_record(p, cell_row(my_cell), cell_col(my_cell), seg_n > 0 ? seg_n : SEG_COUNT)->guess = seg_guess;
                cnt.rays += extra + ((need_ray && m.phase < 5) ? 1u : 0u);      // prelude casts are not reference casts
                if (!need_ray && seg_n > 0) seg_record(p, cell_row(my_cell), cell_col(my_cell), seg_n)->casts += cnt.rays;
            }
            if (!need_ray) {
                has_cell = false; finished_cell = true;
                if (p.row_done) {   // (row_done is only used unsharded: block == local block)
                    const int lb = (cell_row(my_cell) - p.row_begin) >> 2;
                    publish_cell(p, lb, row_slots(p, lb));
                }
            }
        }
        __syncwarp();
        if (need_ray) {
            // the cell's frame (12 registers) is not kept across the traversal loop: three cached loads rebuild it
            const int ci = cell_row(my_cell), cj = cell_col(my_cell);
            const size_t c = (size_t)ci * p.dim_in_1 + cj;
            const F3 nrm = f3(__ldg(p.vec_norm + 3 * c), __ldg(p.vec_norm + 3 * c + 1), __ldg(p.vec_norm + 3 * c + 2));
            const F3 nth = f3(__ldg(p.vec_north + 3 * c), __ldg(p.vec_north + 3 * c + 1), __ldg(p.vec_north + 3 * c + 2));
            const float4 v = __ldg(sv.vert4 + (size_t)(ci + p.offset_0) * sv.W + (cj + p.offset_1));
            const Frame f = make_frame(f3(v.x, v.y, v.z), nrm, nth, p.ray_org_elev);
            const F3 D1 = ray_dir(s, f, ie, m.k);
            const F3 D2 = lo_ie >= 0 ? ray_dir(s, f, lo_ie, m.k) : D1;
            const bool two = wq2_start(sv, sh, warp, lane, L, f.org, D1, D2);
            m.spec_ie = (two && lo_ie >= 0) ? lo_ie : -1;
            have_result = true;
        }
        if (__any_sync(FULL, finished_cell) && (more_tiles || next_cell < 32)) continue;   // give them a new cell first
        const unsigned int cell_mask = __ballot_sync(FULL, has_cell);
        if (cell_mask == 0u) break;
        const int thr = min(refill_thr, __popc(cell_mask));
        __syncwarp();
        // (C) shared traversal loop
        while (__popc(wq2_step<true, false>(sv, sh, warp, lane, tid, L, pend_est, s.dist, wait_thr, cnt, stack_lim)) >= thr) {}
    }
    flush_counters(cnt, 0u, counters);
}

// ---- arbitrary locations (horizon_comp.cpp:828-1094)
__global__ void k_loc_snap(SceneView sv, LocationParams lp, float4* org_valid, Counters* counters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= lp.num_loc) return;
    LaneCounters cnt; cnt.rays = cnt.nodes = cnt.prims = 0;
    unsigned int* ovf = reinterpret_cast<unsigned int*>(&counters->stack_overflow);
    const F3 nrm = f3(lp.vec_norm[3 * i], lp.vec_norm[3 * i + 1], lp.vec_norm[3 * i + 2]);
    const F3 ini = f3(lp.coords[3 * i], lp.coords[3 * i + 1], lp.coords[3 * i + 2]);
    float dist = 100000.0f;                                           // :951-952
    bool hit = trace_bvh2<true>(sv, ini, nrm, dist, cnt, ovf);
    if (!hit) {                                                       // :953-957
        dist = 100000.0f;
        hit = trace_bvh2<true>(sv, ini, f3(-nrm.x, -nrm.y, -nrm.z), dist, cnt, ovf);
        dist = -dist;
    }
    const float lift = __fadd_rn(dist, lp.ray_org_elev[i]);           // :961-963
    org_valid[i] = make_float4(__fadd_rn(ini.x, __fmul_rn(nrm.x, lift)), __fadd_rn(ini.y, __fmul_rn(nrm.y, lift)),
                               __fadd_rn(ini.z, __fmul_rn(nrm.z, lift)), hit ? 1.f : 0.f);
}

__device__ __forceinline__ Frame loc_frame(const LocationParams& lp, const float4* org_valid, int i) {
    const F3 nrm = f3(lp.vec_norm[3 * i], lp.vec_norm[3 * i + 1], lp.vec_norm[3 * i + 2]);
    const F3 nth = f3(lp.vec_north[3 * i], lp.vec_north[3 * i + 1], lp.vec_north[3 * i + 2]);
    Frame f = make_frame(f3(0.f, 0.f, 0.f), nrm, nth, 0.f);
    const float4 o = org_valid[i];
    f.org = f3(o.x, o.y, o.z);
    return f;
}

// chained search (guess_constant): one lane per location
__global__ void k_loc_chain(SceneView sv, HorizonParams p, LocationParams lp, const float4* org_valid, Counters* counters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    LaneCounters cnt; cnt.rays = cnt.nodes = cnt.prims = 0;
    unsigned int units = 0;
    if (i < lp.num_loc && org_valid[i].w != 0.f) {
        const Search s = make_search(sv, p, counters);
        const Frame f = loc_frame(lp, org_valid, i);
        OutBuf ob; ob.init(lp.hori + (size_t)i * p.azim_num, false);
        cell_search<2>(s, f, ob, cnt);
        units = p.azim_num;
    }
    flush_counters(cnt, units, counters);
}

// independent azimuths (discrete_sampling / binary_search): one lane per
// (location, azimuth).  With distance output a miss-only azimuth writes -1 and
// k_loc_dist_fix carries the previous azimuth's distance forward, which is what
// the reference's function-scope dist_hit does (:526-527, :570-571).
template <int ALG, bool WD>
__global__ void k_loc_indep(SceneView sv, HorizonParams p, LocationParams lp, const float4* org_valid, Counters* counters) {
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    LaneCounters cnt; cnt.rays = cnt.nodes = cnt.prims = 0;
    unsigned int units = 0;
    const long long total = (long long)lp.num_loc * p.azim_num;
    if (g < total) {
        const int i = (int)(g / p.azim_num), k = (int)(g - (long long)i * p.azim_num);
        if (org_valid[i].w != 0.f) {
            const Search s = make_search(sv, p, counters);
            const Frame f = loc_frame(lp, org_valid, i);
            const int top = s.elev_num - 1;
            float res, dist_hit = -1.0f;
            if (ALG == 0) {
                int cur = 0, prev = 0; bool hit = true;
                while (hit) {
                    prev = cur; cur = min(cur + 10, top);
                    if (WD) { float d; hit = cast_closest(s, f, cur, k, cnt, d); if (hit) dist_hit = d; }
                    else hit = cast_any(s, f, cur, k, cnt);
                    if (cur == top) hit = false;
                }
                res = midpoint(__ldg(s.elev_ang + prev), __ldg(s.elev_ang + cur));
            } else {
                bisect<WD>(s, f, k, cnt, res, dist_hit);
            }
            lp.hori[g] = res;
            if (WD) lp.hori_dist[g] = dist_hit;
            units = 1;
        }
    }
    flush_counters(cnt, units, counters);
}
// ---------------------------------------------------------------------------
// k_loc_wq: horizon_locations on the production traversal (scope row 8f-2).  One lane per (location,
// azimuth) for the independent algorithms (discrete_sampling / binary_search; the reference's own use:
// 1440 azimuths, hori_dist_out, examples/horizon/locations_curved_DEM.py), one lane per location for the
// guess_constant chain; all lanes of a warp share the packet step of hzb_wq2.cuh in single-ray mode --
// any-hit, or closest-hit when the distance to the horizon is wanted (castRay_intersect1,
// horizon_comp.cpp:268-292, 519-612).  Persistent warps, grid-stride over 32-unit groups.
// ---------------------------------------------------------------------------
struct LocOut {   // one azimuth row of one location
    float* out;
    __device__ __forceinline__ void put(int k, float v) { out[k] = v; }
    __device__ __forceinline__ void put_idx(int k, int, float v) { out[k] = v; }
};

template <int ALG, bool WD>
__global__ void __launch_bounds__(WQ_BLOCK, 4) k_loc_wq(SceneView sv, HorizonParams p, LocationParams lp, const float4* org_valid,
                                                        Counters* counters, int stack_lim) {
    __shared__ Wq2Shared sh;
    const Search s = make_search(sv, p, counters);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
    const unsigned int FULL = 0xffffffffu;
    const long long total = (ALG == 2) ? (long long)lp.num_loc : (long long)lp.num_loc * p.azim_num;
    const long long warps_total = (long long)gridDim.x * WQ_NWARPS;
    LaneCounters cnt; cnt.rays = cnt.nodes = cnt.prims = 0;
    unsigned int units = 0;
    if (lane == 0) { sh.hit1[warp] = 0u; sh.hit2[warp] = 0u; }
    unsigned int pend_est = 0;
    __syncwarp();
    Wq2Lane L; L.state = 0; L.hit1 = L.hit2 = false; L.node = WQ_NONE; L.sp = 0; L.pc = 0; L.tfar = 0.f;
    L.A1x = L.A1y = L.A1z = L.B1x = L.B1y = L.B1z = 0.f; L.A2x = L.A2y = L.A2z = L.B2x = L.B2y = L.B2z = 0.f;
    L.selxy = 0x74107410u;

    for (long long base = ((long long)blockIdx.x * WQ_NWARPS + warp) * 32; base < total; base += warps_total * 32) {
        const long long g = base + lane;
        int loc = 0, k0 = 0;
        bool has_unit = g < total;
        if (has_unit) {
            if (ALG == 2) loc = (int)g; else { loc = (int)(g / p.azim_num); k0 = (int)(g - (long long)loc * p.azim_num); }
            if (org_valid[loc].w == 0.f) has_unit = false;          // the normal line missed the surface: output keeps NaN
        }
        Frame f; LocOut ob; ob.out = nullptr;
        LaneSM m; m.phase = 0; m.k = k0; m.cur = m.prev = m.count = m.prev_az = 0; m.spec_ie = -1; m.spec_hit = false;
        float dist_hit = -1.0f;          // -1: no cast of this azimuth hit (k_loc_dist_fix carries the previous azimuth's distance forward)
        bool have_result = false;
        if (has_unit) {
            f = loc_frame(lp, org_valid, loc);
            ob.out = lp.hori + (size_t)loc * p.azim_num;
            units += (ALG == 2) ? (unsigned int)p.azim_num : 1u;
        }
        while (true) {
            // lanes without a ray in flight consume the last result and ask the search for the next cast
            if (has_unit && L.state == 0) {
                bool hit = L.hit1;
                if (have_result && L.node == WQ_OVF) {     // full traversal stack: this cast is decided by the binary-BVH walker
                    L.node = WQ_NONE;
                    const float* r = &sh.ray[warp][0][lane];
                    float tf = s.dist;
                    hit = trace_bvh2<WD>(sv, f3(r[0], r[32], r[64]), f3(r[96], r[128], r[160]), tf, cnt,
                                         reinterpret_cast<unsigned int*>(&counters->stack_overflow));
                    L.tfar = tf;
                    atomicAdd(&counters->fallback_packets, 1ull);
                }
                if (WD && have_result && hit) dist_hit = L.tfar;                    // :546-553, :592-608
                int ie = 0, lo_ie = -1; unsigned int extra = 0;
                bool need = sm_advance<ALG, false, LocOut>(s, m, have_result, hit, ob, ie, lo_ie, extra);
                if (ALG != 2 && m.k != k0) need = false;        // this lane owns ONE azimuth: the search has moved on to the next
                if (need) {
                    const F3 D = ray_dir(s, f, ie, ALG == 2 ? m.k : k0);
                    wq2_start(sv, sh, warp, lane, L, f.org, D, D);
                    L.tfar = s.dist; sh.tbest[warp][lane] = __float_as_uint(s.dist);
                    have_result = true; cnt.rays++;
                } else {
                    if (WD) lp.hori_dist[(size_t)loc * p.azim_num + k0] = dist_hit;
                    has_unit = false;
                }
            }
            const unsigned int live = __ballot_sync(FULL, L.state != 0);
            if (live == 0u) break;
            const int thr = min(24, __popc(live));
            __syncwarp();
            while (__popc(wq2_step<false, false, WD>(sv, sh, warp, lane, tid, L, pend_est, s.dist, 2, cnt, stack_lim)) >= thr) {}
        }
    }
    flush_counters(cnt, units, counters);
}

__global__ void k_loc_dist_fix(LocationParams lp, int azim_num, const float4* org_valid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= lp.num_loc || org_valid[i].w == 0.f) return;
    float carry = 0.0f;
    float* d = lp.hori_dist + (size_t)i * azim_num;
    for (int k = 0; k < azim_num; ++k) {
        if (d[k] < 0.0f) d[k] = carry; else carry = d[k];
    }
}

}  // namespace

// Memory pool of the segment records, one per device: stream-ordered allocation without synchronisation, and -- unlike
// the device's default pool -- it keeps its few MB across synchronisation points, so a repeated launch never goes
// to the operating system (on some boxes that costs tens of milliseconds).
static std::mutex g_seg_mu;
static std::vector<cudaMemPool_t> g_seg_pools;
void seg_pool_trim() {      // hzb_trim(): give the idle blocks back
    std::lock_guard<std::mutex> lk(g_seg_mu);
    for (cudaMemPool_t mp : g_seg_pools) if (mp) cudaMemPoolTrimTo(mp, 0);
}
static cudaMemPool_t seg_pool(int device) {
    std::mutex& mu = g_seg_mu;
    std::vector<cudaMemPool_t>& pools = g_seg_pools;
    std::lock_guard<std::mutex> lk(mu);
    if (device < 0) return nullptr;
    if ((size_t)device >= pools.size()) pools.resize((size_t)device + 1, nullptr);
    if (!pools[device]) {
        cudaMemPoolProps props{};
        props.allocType = cudaMemAllocationTypePinned; props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice; props.location.id = device;
        cudaMemPool_t mp = nullptr;
        if (cudaMemPoolCreate(&mp, &props) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep);
        pools[device] = mp;
    }
    return pools[device];
}

// Queue layout of one launch (see "Queue order and azimuth segments"): band, interior, split tail.
static void plan_queue(const Scene& s, HorizonParams& p, int grid_ctas) {
    const DebugOptions& o = debug_options();
    const int rows = p.row_end - p.row_begin;
    const int all = (rows + 3) >> 2;
    const int tiles_y = all > p.blk_offset ? (all - p.blk_offset + p.blk_stride - 1) / p.blk_stride : 0;
    const int tiles_x = (p.dim_in_1 + 7) >> 3;
    p.seg_count = 1; p.q_by0 = 0; p.q_by1 = tiles_y; p.q_bx = 0; p.q_tail = 0; p.seg = nullptr;
    queue_sections(p, tiles_x, tiles_y);
    if (o.tail_segments == 1 || o.horizon_kernel != 0 || tiles_y <= 0 || tiles_x <= 0) return;
    const bool forced = o.tail_segments == SEG_COUNT;
    if (p.azim_num < 4 * SEG_COUNT) return;                 // every segment starts behind azimuth 1 and has a few azimuths
    if (p.elev_num < 32) return;                            // the prelude bisects rungs 10 table entries apart: not worth it (or defined) on a tiny table
    const long long warps = (long long)grid_ctas * WQ_NWARPS;
    if (!forced) {
        if (p.azim_num < 64) return;
        // the tail matters when a lane owns few cells; beyond ~40 cells per lane it is below 1 % of the launch
        if ((long long)tiles_x * tiles_y > 40 * warps) return;
    }
    // Band: cells that may see below the table's low limit over the DEM's edge -- closer to the edge than relief / tan(-low).
    int band_x = 0, band_y = 0;      // DEM cells
    if (o.tail_band >= 0) band_x = band_y = o.tail_band;
    else if (p.algorithm == 2) {
        if (!(p.low < -0.01f)) return;                      // low limit near or above the horizontal: chains clamp anywhere
        const double reach = (double)(s.hi[2] - s.lo[2]) / tan(-(double)p.low);
        // a cell could also lose sight of all terrain above the low limit INSIDE the DEM if the search distance were
        // shorter than that reach (a slope falling steeper than the low limit all the way to dist_search): no split then
        if (!(reach < (double)p.dist)) return;
        const double dx = (double)(s.hi[0] - s.lo[0]) / std::max(1, s.W - 1), dy = (double)(s.hi[1] - s.lo[1]) / std::max(1, s.H - 1);
        if (!(dx > 0.0) || !(dy > 0.0)) return;
        band_x = (int)std::min(1e9, ceil(reach / dx)); band_y = (int)std::min(1e9, ceil(reach / dy));
    }
    // inner-domain rows / columns outside the band: [r0, r1) x [c0, c1)
    const int r0 = std::max(0, band_y - p.offset_0), r1 = std::min(p.dim_in_0, s.H - band_y - p.offset_0);
    const int c0 = std::max(0, band_x - p.offset_1), c1 = std::min(p.dim_in_1, s.W - band_x - p.offset_1);
    // whole 4-row blocks of this launch / whole 8-column tiles inside it
    // (a partial last block / tile only holds cells beyond the domain's end: it counts as inside when the band does not cut there)
    const int gb0 = std::max(0, (r0 - p.row_begin + 3) >> 2), gb1 = r1 >= p.row_end ? all : std::max(0, std::min(all, (r1 - p.row_begin) >> 2));
    auto local_below = [&](int gb) { return gb > p.blk_offset ? std::min(tiles_y, (gb - p.blk_offset + p.blk_stride - 1) / p.blk_stride) : 0; };   // local blocks with global index < gb
    const int by0 = local_below(gb0), by1 = std::max(by0, local_below(gb1));
    const int bx = std::max((c0 + 7) >> 3, c1 >= p.dim_in_1 ? 0 : tiles_x - (std::max(c1, 0) >> 3));
    if (by1 <= by0 || 2 * bx >= tiles_x) return;            // no interior
    const long long interior = (long long)(by1 - by0) * (tiles_x - 2 * bx);
    // two tiles per resident warp: the split part outlasts the last whole chains (measured on cfg2: one tile per warp
    // leaves the end to chance, more than two only adds preludes)
    long long tail = o.tail_tiles >= 0 ? o.tail_tiles : 2 * warps;
    tail = std::min(tail, interior);
    if (tail <= 0) return;
    p.seg_count = SEG_COUNT; p.q_by0 = by0; p.q_by1 = by1; p.q_bx = bx; p.q_tail = (unsigned int)tail;
    queue_sections(p, tiles_x, tiles_y);
}

// host-only view of plan_queue for the CPU suite (hzb_plan_queue, api.cu)
void plan_queue_host(const Scene& s, HorizonParams& p, int grid_ctas) { plan_queue(s, p, grid_ctas); }

int launch_horizon_gridded(Scene& s, const HorizonParams& p_in, cudaStream_t st) {
    HorizonParams p = p_in;
    if (p.row_end <= p.row_begin || p.dim_in_1 <= 0) return 0;
    unsigned int* tile_counter = scene_tile_counter(s, st);
    if (!tile_counter) return 1;
    const SceneView sv = s.view();
    const DebugOptions& o = debug_options();
    if (p.hori_q && (p.algorithm != 2 || p.elev_num > 65534)) { set_error("quantised output needs ray_algorithm guess_constant and at most 65534 table entries"); return 1; }
    p.seg_count = 1; p.q_by0 = 0; p.q_by1 = 0; p.q_bx = 0; p.q_tail = 0; p.seg = nullptr;     // (plan_queue fills them for the production kernel)
    if (o.horizon_kernel == 1 && !p.hori_q) {   // reference-shaped per-lane kernel on the binary BVH (second implementation for the parity tests)
        const int grid = sm_count() * 8;
        switch (p.algorithm) {
            case 0: k_horizon_gridded<0><<<grid, HG_THREADS, 0, st>>>(sv, p, s.d_counters, tile_counter); break;
            case 1: k_horizon_gridded<1><<<grid, HG_THREADS, 0, st>>>(sv, p, s.d_counters, tile_counter); break;
            default: k_horizon_gridded<2><<<grid, HG_THREADS, 0, st>>>(sv, p, s.d_counters, tile_counter); break;
        }
    } else {
        // a warp leaves the traversal loop to refill when fewer than w_refill lanes hold a packet; pending
        // candidates are flushed when w_wait lanes wait on theirs (tuned on B200)
        const int w_refill = o.wrefill, w_wait = o.wwait, stack_lim = std::max(1, std::min(o.stack_limit, WQ_STACK_N));
#ifndef HZB_MB
#define HZB_MB 6
#endif
        constexpr int MB = HZB_MB;   // resident CTAs per SM (80 registers, 26 KB shared memory): the walk is bound by per-warp latency, a sixth CTA is worth 2-7 %
        const int cps = (o.ctas_per_sm >= 1 && o.ctas_per_sm <= MB) ? o.ctas_per_sm : MB;
        const int grid = sm_count() * cps;
        plan_queue(s, p, grid);
        if (p.seg_count > 1) {
            // records of the split cells: stream-ordered allocation (no synchronisation, the block returns to the
            // device's pool behind the fix-up kernel), initialised to SEG_NONE
            const size_t bytes = (size_t)p.q_tail * 32 * SEG_COUNT * sizeof(SegRecord);
            cudaMemPool_t pool = seg_pool(s.device);
            if (!pool || cudaMallocFromPoolAsync((void**)&p.seg, bytes, pool, st) != cudaSuccess) {   // no records, no split cells
                cudaGetLastError();
                p.seg = nullptr; p.seg_count = 1; p.q_tail = 0; queue_sections(p, p.q_tiles_x, p.q_tiles_y);
            }
            else HZB_CUDA(cudaMemsetAsync(p.seg, 0xFF, bytes, st));
        }
        if (p.hori_q) {   // quantised output: guess_constant only (checked by the caller)
            k_horizon_wq6<2, MB, true><<<grid, WQ_BLOCK, 0, st>>>(sv, p, s.d_counters, tile_counter, w_refill, w_wait, stack_lim);
        } else switch (p.algorithm) {
            case 0: k_horizon_wq6<0, MB, false><<<grid, WQ_BLOCK, 0, st>>>(sv, p, s.d_counters, tile_counter, w_refill, w_wait, stack_lim); break;
            case 1: k_horizon_wq6<1, MB, false><<<grid, WQ_BLOCK, 0, st>>>(sv, p, s.d_counters, tile_counter, w_refill, w_wait, stack_lim); break;
            default: k_horizon_wq6<2, MB, false><<<grid, WQ_BLOCK, 0, st>>>(sv, p, s.d_counters, tile_counter, w_refill, w_wait, stack_lim); break;
        }
        // cells whose traversal stack was full (none in practice) and segments that started from a wrong index are recomputed
        const long long slots = (long long)((p.row_end - p.row_begin + 3) / 4) * 4 * p.dim_in_1;
        const int rgrid = (int)std::min<long long>((slots + HG_THREADS - 1) / HG_THREADS, (long long)sm_count() * 8);
        if (p.hori_q) k_horizon_redo<2, true><<<rgrid, HG_THREADS, 0, st>>>(sv, p, s.d_counters);
        else switch (p.algorithm) {
            case 0: k_horizon_redo<0, false><<<rgrid, HG_THREADS, 0, st>>>(sv, p, s.d_counters); break;
            case 1: k_horizon_redo<1, false><<<rgrid, HG_THREADS, 0, st>>>(sv, p, s.d_counters); break;
            default: k_horizon_redo<2, false><<<rgrid, HG_THREADS, 0, st>>>(sv, p, s.d_counters); break;
        }
        if (p.seg) cudaFreeAsync(p.seg, st);
    }
    HZB_CUDA(cudaGetLastError());
    return 0;
}

int launch_horizon_locations(Scene& s, const HorizonParams& p, const LocationParams& lp, cudaStream_t st) {
    if (lp.num_loc <= 0) return 0;
    float4* d_org = (float4*)pool_alloc((size_t)lp.num_loc * sizeof(float4));
    if (!d_org) return 1;
    struct OrgGuard { float4* p; ~OrgGuard() { pool_free(p); } } org_guard{d_org};
    const SceneView sv = s.view();
    const int nb_loc = (lp.num_loc + 63) / 64;
    k_loc_snap<<<nb_loc, 64, 0, st>>>(sv, lp, d_org, s.d_counters);      // surface snap: two closest-hit rays per location (:951-957)
    const long long total = (long long)lp.num_loc * p.azim_num;
    if (debug_options().horizon_kernel == 1) {
        // second implementation (parity tests): per-lane searches on the binary BVH
        if (p.algorithm == 2) {
            k_loc_chain<<<(lp.num_loc + 31) / 32, 32, 0, st>>>(sv, p, lp, d_org, s.d_counters);
        } else {
            const int nb = (int)((total + 127) / 128);
            if (p.algorithm == 0) {
                if (lp.hori_dist_out) k_loc_indep<0, true><<<nb, 128, 0, st>>>(sv, p, lp, d_org, s.d_counters);
                else k_loc_indep<0, false><<<nb, 128, 0, st>>>(sv, p, lp, d_org, s.d_counters);
            } else {
                if (lp.hori_dist_out) k_loc_indep<1, true><<<nb, 128, 0, st>>>(sv, p, lp, d_org, s.d_counters);
                else k_loc_indep<1, false><<<nb, 128, 0, st>>>(sv, p, lp, d_org, s.d_counters);
            }
        }
    } else {
        const long long units = p.algorithm == 2 ? (long long)lp.num_loc : total;
        const int grid = (int)std::max<long long>(1, std::min<long long>((units + WQ_BLOCK - 1) / WQ_BLOCK, (long long)sm_count() * 4));
        const int stack_lim = std::max(1, std::min(debug_options().stack_limit, WQ_STACK_N));
#define HZB_LOC(ALG_, WD_) k_loc_wq<ALG_, WD_><<<grid, WQ_BLOCK, 0, st>>>(sv, p, lp, d_org, s.d_counters, stack_lim)
        if (p.algorithm == 2) HZB_LOC(2, false);
        else if (p.algorithm == 0) { if (lp.hori_dist_out) HZB_LOC(0, true); else HZB_LOC(0, false); }
        else { if (lp.hori_dist_out) HZB_LOC(1, true); else HZB_LOC(1, false); }
#undef HZB_LOC
    }
    if (lp.hori_dist_out && p.algorithm != 2) k_loc_dist_fix<<<nb_loc, 64, 0, st>>>(lp, p.azim_num, d_org);
    HZB_CUDA(cudaGetLastError());
    HZB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // namespace hzb
