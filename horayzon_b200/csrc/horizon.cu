// horizon.cu -- horizon search kernels (B200, sm_100a).
//
// Replaces the TBB row loop and the three search algorithms of the reference
// (horizon_comp.cpp:302-498, 519-612, 739-800, 926-1070).  One lane owns one
// grid cell and walks its azimuth chain; warps pull 8x4-cell tiles from an
// atomic queue (persistent CTAs, grid = SM count x resident CTAs).  All
// arithmetic that decides a table index or a ray direction is written with
// explicitly rounded intrinsics in the reference's float/double mix, so the
// outputs are decision-exact against the CPU oracle.
#include "hzb_geom.cuh"
#include <math.h>
#include <string.h>
#include <stdlib.h>

namespace hzb {

// ------------------------------------------------------------ host tables
static inline float deg2rad_f(float a) { return (float)(((double)a / 180.0) * M_PI); }  // horizon_comp.cpp:37-39

void HorizonTables::make(int azim_n, float dist_km, float acc_deg, float low_deg) {
    azim_num = azim_n;
    acc = deg2rad_f(acc_deg);
    low = deg2rad_f(low_deg);
    up = deg2rad_f(89.98f);                      // horizon_comp.cpp:648
    dist = (float)((double)dist_km * 1000.0);    // :670
    step = (double)acc / 5.0;
    azim_sin.resize(azim_n); azim_cos.resize(azim_n);
    for (int i = 0; i < azim_n; ++i) {           // :714-718: float angle, float sin/cos overloads
        const float ang = (float)((2 * M_PI) / azim_n * i);
        azim_sin[i] = sinf(ang); azim_cos[i] = cosf(ang);
    }
    elev_num = (int)ceil((double)(up - low) / step) + 1;  // :721-722
    elev_ang.resize(elev_num); elev_sin.resize(elev_num); elev_cos.resize(elev_num);
    for (int i = 0; i < elev_num; ++i) {         // :726-731: anchored at the upper limit
        const float ang = (float)((double)up - step * i);
        elev_ang[elev_num - i - 1] = ang;
        elev_sin[elev_num - i - 1] = sinf(ang);
        elev_cos[elev_num - i - 1] = cosf(ang);
    }
}

int upload_tables(Scene& s, const HorizonTables& T, HorizonParams& p, cudaStream_t st) {
    const size_t need = (size_t)2 * T.azim_num + (size_t)3 * T.elev_num;
    if (need > s.tables_cap) {
        if (s.d_tables) cudaFree(s.d_tables);
        s.d_tables = nullptr; s.tables_cap = 0;
        HZB_CUDA(cudaMalloc((void**)&s.d_tables, need * sizeof(float)));
        s.tables_cap = need;
    }
    std::vector<float> h(need);
    float* q = h.data();
    memcpy(q, T.azim_sin.data(), 4 * (size_t)T.azim_num); q += T.azim_num;
    memcpy(q, T.azim_cos.data(), 4 * (size_t)T.azim_num); q += T.azim_num;
    memcpy(q, T.elev_ang.data(), 4 * (size_t)T.elev_num); q += T.elev_num;
    memcpy(q, T.elev_sin.data(), 4 * (size_t)T.elev_num); q += T.elev_num;
    memcpy(q, T.elev_cos.data(), 4 * (size_t)T.elev_num);
    // synchronous copy from pageable memory: the host vector may die afterwards
    HZB_CUDA(cudaMemcpyAsync(s.d_tables, h.data(), need * sizeof(float), cudaMemcpyHostToDevice, st));
    HZB_CUDA(cudaStreamSynchronize(st));
    p.azim_sin = s.d_tables; p.azim_cos = p.azim_sin + T.azim_num;
    p.elev_ang = p.azim_cos + T.azim_num; p.elev_sin = p.elev_ang + T.elev_num; p.elev_cos = p.elev_sin + T.elev_num;
    p.azim_num = T.azim_num; p.elev_num = T.elev_num;
    p.acc = T.acc; p.low = T.low; p.up = T.up; p.dist = T.dist; p.step = T.step;
    return 0;
}

// ------------------------------------------------------------ device side
namespace {

struct Frame {  // per-cell local frame: columns east, north, norm (horizon_comp.cpp:773-779)
    F3 org;
    float m00, m01, m02, m10, m11, m12, m20, m21, m22;
};

__device__ __forceinline__ Frame make_frame(F3 vert, F3 norm, F3 north, float lift) {
    Frame f;
    f.org = f3(__fadd_rn(vert.x, __fmul_rn(norm.x, lift)), __fadd_rn(vert.y, __fmul_rn(norm.y, lift)),
               __fadd_rn(vert.z, __fmul_rn(norm.z, lift)));
    const float ex = __fsub_rn(__fmul_rn(north.y, norm.z), __fmul_rn(north.z, norm.y));
    const float ey = __fsub_rn(__fmul_rn(north.z, norm.x), __fmul_rn(north.x, norm.z));
    const float ez = __fsub_rn(__fmul_rn(north.x, norm.y), __fmul_rn(north.y, norm.x));
    f.m00 = ex; f.m01 = north.x; f.m02 = norm.x;
    f.m10 = ey; f.m11 = north.y; f.m12 = norm.y;
    f.m20 = ez; f.m21 = north.z; f.m22 = norm.z;
    return f;
}

struct Search {  // everything a lane needs to cast one ray of the search
    SceneView sv;
    const float* __restrict__ azim_sin; const float* __restrict__ azim_cos;
    const float* __restrict__ elev_ang; const float* __restrict__ elev_sin; const float* __restrict__ elev_cos;
    int azim_num, elev_num;
    float acc, low, up, dist; double step;
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

__device__ __forceinline__ F3 ray_dir(const Search& s, const Frame& f, int ie, int k) {
    const float ec = __ldg(s.elev_cos + ie), es = __ldg(s.elev_sin + ie);
    const float r0 = __fmul_rn(ec, __ldg(s.azim_sin + k)), r1 = __fmul_rn(ec, __ldg(s.azim_cos + k)), r2 = es;
    return f3(__fadd_rn(__fadd_rn(__fmul_rn(f.m00, r0), __fmul_rn(f.m01, r1)), __fmul_rn(f.m02, r2)),
              __fadd_rn(__fadd_rn(__fmul_rn(f.m10, r0), __fmul_rn(f.m11, r1)), __fmul_rn(f.m12, r2)),
              __fadd_rn(__fadd_rn(__fmul_rn(f.m20, r0), __fmul_rn(f.m21, r1)), __fmul_rn(f.m22, r2)));
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

__device__ __forceinline__ int index_of(const Search& s, float elev) {  // (int)roundf((elev-low)/(acc/5.0))
    const double q = __ddiv_rn((double)__fsub_rn(elev, s.low), s.step);
    return (int)roundf(__double2float_rn(q));
}
__device__ __forceinline__ float midpoint(float a, float b) { return __fmul_rn(__fadd_rn(a, b), 0.5f); }

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
    float* out; bool vec, writer; float b0, b1, b2, b3;
    __device__ __forceinline__ void init(float* o, bool v, bool w = true) { out = o; vec = v; writer = w; b0 = b1 = b2 = b3 = 0.f; }
    __device__ __forceinline__ void put(int k, float v) {
        if (!vec) { if (writer) out[k] = v; return; }
        b0 = b1; b1 = b2; b2 = b3; b3 = v;
        if ((k & 3) == 3 && writer) *reinterpret_cast<float4*>(out + (k - 3)) = make_float4(b0, b1, b2, b3);
    }
};

// One cell, all azimuths.  ALG 0 discrete_sampling (:302-333), 1 binary_search
// (:339-381), 2 guess_constant (:387-498).  Termination rule: a hit
// at the top index counts as a miss, a miss at index 0 as a hit.
template <int ALG>
__device__ void cell_search(const Search& s, const Frame& f, OutBuf& ob, LaneCounters& cnt) {
    const int top = s.elev_num - 1;
    if (ALG == 0) {
        for (int k = 0; k < s.azim_num; ++k) {
            int cur = 0, prev = 0; bool hit = true;
            while (hit) {
                prev = cur; cur = min(cur + 10, top);
                hit = cast_any(s, f, cur, k, cnt);
                if (cur == top) hit = false;
            }
            ob.put(k, midpoint(__ldg(s.elev_ang + prev), __ldg(s.elev_ang + cur)));
        }
    } else if (ALG == 1) {
        for (int k = 0; k < s.azim_num; ++k) {
            float mid, dh = 0.f;
            bisect<false>(s, f, k, cnt, mid, dh);
            ob.put(k, mid);
        }
    } else {
        float mid, dh = 0.f;
        int prev_az = bisect<false>(s, f, 0, cnt, mid, dh);
        ob.put(0, mid);
        for (int k = 1; k < s.azim_num; ++k) {
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
            ob.put(k, __ldg(s.elev_ang + ie));
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

constexpr int HG_THREADS = 128;

template <int ALG>
__global__ void __launch_bounds__(HG_THREADS) k_horizon_gridded(SceneView sv, HorizonParams p, Counters* counters,
                                                                unsigned int* tile_counter) {
    const Search s = make_search(sv, p, counters);
    const int lane = threadIdx.x & 31;
    const int rows = p.row_end - p.row_begin;
    const int tiles_x = (p.dim_in_1 + 7) >> 3, tiles_y = (rows + 3) >> 2;
    const unsigned int num_tiles = (unsigned int)tiles_x * (unsigned int)tiles_y;
    const bool vec = (p.azim_num & 3) == 0 && ((reinterpret_cast<size_t>(p.hori) & 15) == 0);
    LaneCounters cnt; cnt.rays = cnt.nodes = cnt.prims = 0;
    while (true) {
        unsigned int tile = 0;
        if (lane == 0) tile = atomicAdd(tile_counter, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= num_tiles) break;
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const int i = p.row_begin + ty * 4 + (lane >> 3), j = tx * 8 + (lane & 7);
        unsigned int units = 0;
        if (i < p.row_end && j < p.dim_in_1) {
            const size_t c = (size_t)i * p.dim_in_1 + j;
            float* out = p.hori + c * p.azim_num;
            if (p.mask[c] == 1) {
                const F3 nrm = f3(p.vec_norm[3 * c], p.vec_norm[3 * c + 1], p.vec_norm[3 * c + 2]);
                const F3 nth = f3(p.vec_north[3 * c], p.vec_north[3 * c + 1], p.vec_north[3 * c + 2]);
                const float4 v = sv.vert4[(size_t)(i + p.offset_0) * sv.W + (j + p.offset_1)];
                const Frame f = make_frame(f3(v.x, v.y, v.z), nrm, nth, p.ray_org_elev);
                OutBuf ob; ob.init(out, vec);
                cell_search<ALG>(s, f, ob, cnt);
                units = p.azim_num;
            } else {
                for (int k = 0; k < p.azim_num; ++k) out[k] = p.hori_fill;  // horizon_comp.cpp:789-794
            }
        }
        flush_counters(cnt, units, counters);
    }
}


// ===========================================================================
// Persistent state-machine kernel (the production path for gridded domains).
//
// The simple kernel above inlines the traversal at three call sites and tests
// leaves inside the node loop, so lanes of a warp sit in different copies of
// the code (ncu r01: 4.9 of 32 lanes active per instruction).  Here every lane
// owns one cell and runs the search as an explicit state machine; there is ONE
// traversal loop per warp in which all lanes step through BVH nodes together,
// each on its own ray:
//   * refill: the loop is left only when fewer than `thr` lanes still have a
//     ray in flight (warp ballot); lanes whose ray finished then advance their
//     state machine (next table index / next azimuth / cell done) and re-enter
//     with a fresh ray while the others resume where they stopped;
//   * postponed leaves: primitives found during node steps are queued per lane
//     (4 registers) and tested in a separate leaf step that runs only when
//     enough lanes have queued work or a lane cannot continue otherwise.
// The hit/miss DECISIONS are those of cell_search<ALG> above, bit for bit.
// ===========================================================================
constexpr int NODE_NONE = 0x7fffffff;
constexpr uint32_t G_NONE = 0xFFFFFFFFu;
constexpr int PEND_MAX = 4;

struct LaneSM {
    // search state (horizon_comp.cpp:387-498 unrolled into states)
    int phase;       // 0 idle/no cell, 1 bisect, 2 upward, 3 downward, 4 discrete
    int k, cur, prev, count, prev_az;
    float lim_up, lim_low, samp;
};

template <int ALG>
__device__ __forceinline__ bool sm_begin_azimuth(const Search& s, LaneSM& m, int& cast_ie) {
    // returns true if a cast is required (cast_ie set), false if the azimuth needs none
    const int top = s.elev_num - 1;
    if (ALG == 0) {
        m.phase = 4; m.prev = 0; m.cur = min(10, top); cast_ie = m.cur; return true;
    } else if (ALG == 1 || m.k == 0) {
        m.phase = 1; m.lim_up = s.up; m.lim_low = s.low;
        m.samp = midpoint(m.lim_up, m.lim_low);
        m.cur = index_of(s, m.samp);
        const float ea = __ldg(s.elev_ang + m.cur);
        if (fmaxf(__fsub_rn(m.lim_up, ea), __fsub_rn(ea, m.lim_low)) > s.acc) { cast_ie = m.cur; return true; }
        return false;
    } else {
        m.phase = 2; m.count = 0;
        m.prev = max(m.prev_az - 5, 0); m.cur = min(m.prev + 10, top); cast_ie = m.cur; return true;
    }
}

// Consume the result of the last cast (if any) and move on until the next cast
// is known or the cell is finished.  Returns true with cast_ie set when a ray
// must be traced; false when the cell is complete.
template <int ALG>
__device__ __forceinline__ bool sm_advance(const Search& s, LaneSM& m, bool have_result, bool hit, OutBuf& ob, int& cast_ie) {
    const int top = s.elev_num - 1;
    while (true) {
        if (!have_result) {  // start of an azimuth
            if (sm_begin_azimuth<ALG>(s, m, cast_ie)) return true;
            // bisect needed no cast at all: fall through to "azimuth finished" with phase 1
            hit = false; have_result = true;
            // (emulate loop exit below)
            goto bisect_done;
        }
        if (m.phase == 1) {
            {
                const float ea = __ldg(s.elev_ang + m.cur);
                if (hit) m.lim_low = ea; else m.lim_up = ea;
                m.samp = midpoint(m.lim_up, m.lim_low);
                m.cur = index_of(s, m.samp);
                const float ea2 = __ldg(s.elev_ang + m.cur);
                if (fmaxf(__fsub_rn(m.lim_up, ea2), __fsub_rn(ea2, m.lim_low)) > s.acc) { cast_ie = m.cur; return true; }
            }
        bisect_done:
            ob.put(m.k, m.samp);          // un-quantised midpoint (:377, :428)
            m.prev_az = m.cur;            // seeds the chain (:429)
        } else if (m.phase == 2) {
            m.count++;
            if (m.cur == top) hit = false;            // termination rule
            if (hit) { m.prev = m.cur; m.cur = min(m.cur + 10, top); cast_ie = m.cur; return true; }
            if (m.count <= 1) {                       // first upward cast missed: search downwards (:471-488)
                m.phase = 3;
                m.prev = min(m.prev_az + 5, top); m.cur = max(m.prev - 10, 0); cast_ie = m.cur; return true;
            }
            const int ie = index_of(s, midpoint(__ldg(s.elev_ang + m.prev), __ldg(s.elev_ang + m.cur)));
            ob.put(m.k, __ldg(s.elev_ang + ie)); m.prev_az = ie;
        } else if (m.phase == 3) {
            if (m.cur == 0) hit = true;               // termination rule
            if (!hit) { m.prev = m.cur; m.cur = max(m.cur - 10, 0); cast_ie = m.cur; return true; }
            const int ie = index_of(s, midpoint(__ldg(s.elev_ang + m.prev), __ldg(s.elev_ang + m.cur)));
            ob.put(m.k, __ldg(s.elev_ang + ie)); m.prev_az = ie;
        } else {  // phase 4: discrete sampling (:309-331)
            if (m.cur == top) hit = false;
            if (hit) { m.prev = m.cur; m.cur = min(m.cur + 10, top); cast_ie = m.cur; return true; }
            ob.put(m.k, midpoint(__ldg(s.elev_ang + m.prev), __ldg(s.elev_ang + m.cur)));
        }
        // azimuth finished
        m.k++;
        if (m.k >= s.azim_num) { m.phase = 0; return false; }
        have_result = false;
    }
}

constexpr int SM_THREADS = 128;
#ifndef SM_MINB
#define SM_MINB 4
#endif

template <int ALG>
__global__ void __launch_bounds__(SM_THREADS, SM_MINB) k_horizon_sm(SceneView sv, HorizonParams p, Counters* counters,
                                                              unsigned int* tile_counter, int refill_thr, int leaf_thr) {
    const Search s = make_search(sv, p, counters);
    const int lane = threadIdx.x & 31;
    const int rows = p.row_end - p.row_begin;
    const int tiles_x = (p.dim_in_1 + 7) >> 3, tiles_y = (rows + 3) >> 2;
    const unsigned int num_tiles = (unsigned int)tiles_x * (unsigned int)tiles_y;
    const bool vec = (p.azim_num & 3) == 0 && ((reinterpret_cast<size_t>(p.hori) & 15) == 0);
    LaneCounters cnt; cnt.rays = cnt.nodes = cnt.prims = 0;
    int stack[HZB_STACK2];

    while (true) {
        unsigned int tile = 0;
        if (lane == 0) tile = atomicAdd(tile_counter, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= num_tiles) break;
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const int ci = p.row_begin + ty * 4 + (lane >> 3), cj = tx * 8 + (lane & 7);

        // ---- per-lane cell set-up
        LaneSM m; m.phase = 0; m.k = 0; m.cur = m.prev = m.count = m.prev_az = 0; m.lim_up = m.lim_low = m.samp = 0.f;
        Frame f; OutBuf ob; ob.init(nullptr, false);
        bool has_cell = false;
        unsigned int units = 0;
        if (ci < p.row_end && cj < p.dim_in_1) {
            const size_t c = (size_t)ci * p.dim_in_1 + cj;
            float* out = p.hori + c * p.azim_num;
            if (p.mask[c] == 1) {
                const F3 nrm = f3(p.vec_norm[3 * c], p.vec_norm[3 * c + 1], p.vec_norm[3 * c + 2]);
                const F3 nth = f3(p.vec_north[3 * c], p.vec_north[3 * c + 1], p.vec_north[3 * c + 2]);
                const float4 v = sv.vert4[(size_t)(ci + p.offset_0) * sv.W + (cj + p.offset_1)];
                f = make_frame(f3(v.x, v.y, v.z), nrm, nth, p.ray_org_elev);
                ob.init(out, vec);
                has_cell = true; units = p.azim_num;
            } else {
                for (int k = 0; k < p.azim_num; ++k) out[k] = p.hori_fill;  // horizon_comp.cpp:789-794
            }
        }

        // ---- ray state
        bool ray_active = false, ray_hit = false, have_result = false;
        F3 D = f3(0.f, 0.f, 1.f); RayInv inv = make_inv(D);
        int node = NODE_NONE, sp = 0, npend = 0;
        unsigned int pq0 = 0, pq1 = 0, pq2 = 0, pq3 = 0;

        while (true) {
            // (1) refill: lanes with a cell but no ray in flight advance their state machine
            if (has_cell && !ray_active) {
                int ie;
                if (sm_advance<ALG>(s, m, have_result, ray_hit, ob, ie)) {
                    D = ray_dir(s, f, ie, m.k); inv = make_inv(D);
                    node = 0; sp = 0; npend = 0; ray_active = true; cnt.rays++;
                } else has_cell = false;
            }
            const unsigned int cell_mask = __ballot_sync(0xffffffffu, has_cell);
            if (cell_mask == 0u) break;
            const int thr = min(refill_thr, __popc(cell_mask));

            // (2) shared traversal loop: every lane steps its own ray
            while (true) {
                if (ray_active && node != NODE_NONE && npend <= PEND_MAX - 2) {
                    const float4* np = reinterpret_cast<const float4*>(sv.nodes2 + node);
                    const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
                    cnt.nodes++;
                    const float lo0[3] = {n0.x, n0.y, n0.z}, hi0[3] = {n0.w, n1.x, n1.y};
                    const float lo1[3] = {n1.z, n1.w, n2.x}, hi1[3] = {n2.y, n2.z, n2.w};
                    const int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
                    float tn0, tn1;
                    bool h0 = slab(lo0, hi0, f.org, inv, s.dist, tn0);
                    bool h1 = slab(lo1, hi1, f.org, inv, s.dist, tn1);
                    if (h0 && c0 < 0) { pq3 = pq2; pq2 = pq1; pq1 = pq0; pq0 = (unsigned int)(~c0); npend++; h0 = false; }
                    if (h1 && c1 < 0) { pq3 = pq2; pq2 = pq1; pq1 = pq0; pq0 = (unsigned int)(~c1); npend++; h1 = false; }
                    if (h0 && h1) {
                        int nearc = c0, farc = c1;
                        if (tn1 < tn0) { nearc = c1; farc = c0; }
                        if (sp < HZB_STACK2) stack[sp++] = farc; else atomicAdd(s.overflow, 1u);
                        node = nearc;
                    } else if (h0) node = c0;
                    else if (h1) node = c1;
                    else node = (sp > 0) ? stack[--sp] : NODE_NONE;
                }
                // leaf step: only when worthwhile or unavoidable
                const bool pending = ray_active && npend > 0;
                const bool must = pending && (npend > PEND_MAX - 2 || node == NODE_NONE);
                const unsigned int pend_mask = __ballot_sync(0xffffffffu, pending);
                if (__any_sync(0xffffffffu, must) || __popc(pend_mask) >= leaf_thr) {
                    if (pending) {
                        const unsigned int prim = pq0; pq0 = pq1; pq1 = pq2; pq2 = pq3; npend--;
                        cnt.prims++;
                        float tfar = s.dist;
                        if (prim_hit<false>(sv, prim, f.org, D, tfar)) { ray_active = false; ray_hit = true; have_result = true; }
                    }
                }
                if (ray_active && node == NODE_NONE && npend == 0) { ray_active = false; ray_hit = false; have_result = true; }
                if (__popc(__ballot_sync(0xffffffffu, ray_active)) < thr) break;
            }
        }
        flush_counters(cnt, units, counters);
    }
}


// ===========================================================================
// Warp-queue kernel (production path for gridded domains).
//
// Same per-lane state machine and refill as k_horizon_sm, plus:
//   * leaf primitives found during node steps go into a ring buffer SHARED BY
//     THE WARP (slot = ballot rank), tagged with the owning lane; when 32 are
//     queued, all 32 lanes test one candidate each against the owner's ray
//     (ray parameters are mirrored in shared memory), so the ray/triangle code
//     runs at full SIMT width no matter which lanes produced the work.  A ray
//     is only retired once its queued candidates have been tested (FIFO
//     sequence numbers), so no candidate can outlive its ray;
//   * the traversal stack lives in shared memory ([entry][thread], conflict
//     free) and push / pop are predicated, not branched;
//   * 6 CTAs per SM.
// Decisions are unchanged: every candidate is tested with tri_hit exactly as
// in the simple kernel, only the order and the lane doing the test differ.
// ===========================================================================
constexpr int WQ_THREADS = 128;
constexpr int WQ_WARPS = WQ_THREADS / 32;
constexpr int WQ_STACK = 48;
constexpr int WQ_RING = 128;

struct WqShared {
    int stack[WQ_STACK][WQ_THREADS];
    uint32_t ring_prim[WQ_WARPS][WQ_RING];
    uint32_t ring_owner[WQ_WARPS][WQ_RING];
    float ray[WQ_WARPS][6][32];     // O.xyz, D.xyz per lane
    unsigned int hitmask[WQ_WARPS];
};

template <int ALG>
__global__ void __launch_bounds__(WQ_THREADS, 6) k_horizon_wq(SceneView sv, HorizonParams p, Counters* counters,
                                                              unsigned int* tile_counter, int refill_thr, int wait_thr) {
    __shared__ WqShared sh;
    const Search s = make_search(sv, p, counters);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
    const unsigned int FULL = 0xffffffffu, lt_mask = (1u << lane) - 1u;
    const int rows = p.row_end - p.row_begin;
    const int tiles_x = (p.dim_in_1 + 7) >> 3, tiles_y = (rows + 3) >> 2;
    const unsigned int num_tiles = (unsigned int)tiles_x * (unsigned int)tiles_y;
    const bool vec = (p.azim_num & 3) == 0 && ((reinterpret_cast<size_t>(p.hori) & 15) == 0);
    LaneCounters cnt; cnt.rays = cnt.nodes = cnt.prims = 0;
    uint32_t* ring_prim = sh.ring_prim[warp];
    uint32_t* ring_owner = sh.ring_owner[warp];
    if (lane == 0) sh.hitmask[warp] = 0u;
    unsigned int pushed = 0, tested = 0;   // warp-uniform FIFO sequence numbers
    __syncwarp();

    while (true) {
        unsigned int tile = 0;
        if (lane == 0) tile = atomicAdd(tile_counter, 1u);
        tile = __shfl_sync(FULL, tile, 0);
        if (tile >= num_tiles) break;
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const int ci = p.row_begin + ty * 4 + (lane >> 3), cj = tx * 8 + (lane & 7);

        LaneSM m; m.phase = 0; m.k = 0; m.cur = m.prev = m.count = m.prev_az = 0; m.lim_up = m.lim_low = m.samp = 0.f;
        Frame f; OutBuf ob; ob.init(nullptr, false);
        bool has_cell = false;
        unsigned int units = 0;
        if (ci < p.row_end && cj < p.dim_in_1) {
            const size_t c = (size_t)ci * p.dim_in_1 + cj;
            float* out = p.hori + c * p.azim_num;
            if (p.mask[c] == 1) {
                const F3 nrm = f3(p.vec_norm[3 * c], p.vec_norm[3 * c + 1], p.vec_norm[3 * c + 2]);
                const F3 nth = f3(p.vec_north[3 * c], p.vec_north[3 * c + 1], p.vec_north[3 * c + 2]);
                const float4 v = sv.vert4[(size_t)(ci + p.offset_0) * sv.W + (cj + p.offset_1)];
                f = make_frame(f3(v.x, v.y, v.z), nrm, nth, p.ray_org_elev);
                ob.init(out, vec);
                has_cell = true; units = p.azim_num;
                sh.ray[warp][0][lane] = f.org.x; sh.ray[warp][1][lane] = f.org.y; sh.ray[warp][2][lane] = f.org.z;
            } else {
                for (int k = 0; k < p.azim_num; ++k) out[k] = p.hori_fill;  // horizon_comp.cpp:789-794
            }
        }

        bool ray_active = false, ray_hit = false, have_result = false;
        RayInv inv; inv.ix = inv.iy = inv.iz = 1.f;
        int node = NODE_NONE, sp = 0;
        unsigned int my_last = 0; bool have_queued = false;   // sequence number of this ray's newest candidate

        while (true) {
            // (1) refill
            if (has_cell && !ray_active) {
                int ie;
                if (sm_advance<ALG>(s, m, have_result, ray_hit, ob, ie)) {
                    const F3 D = ray_dir(s, f, ie, m.k);
                    inv = make_inv(D);
                    sh.ray[warp][3][lane] = D.x; sh.ray[warp][4][lane] = D.y; sh.ray[warp][5][lane] = D.z;
                    node = 0; sp = 0; ray_active = true; ray_hit = false; have_queued = false; cnt.rays++;
                } else has_cell = false;
            }
            const unsigned int cell_mask = __ballot_sync(FULL, has_cell);
            if (cell_mask == 0u) break;
            const int thr = min(refill_thr, __popc(cell_mask));
            __syncwarp();

            // (2) traversal
            while (true) {
                // ---- node step (predicated)
                const bool do_node = ray_active && !ray_hit && node != NODE_NONE;
                bool l0 = false, l1 = false; int c0 = 0, c1 = 0;
                if (do_node) {
                    const float4* np = reinterpret_cast<const float4*>(sv.nodes2 + node);
                    const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
                    cnt.nodes++;
                    const float lo0[3] = {n0.x, n0.y, n0.z}, hi0[3] = {n0.w, n1.x, n1.y};
                    const float lo1[3] = {n1.z, n1.w, n2.x}, hi1[3] = {n2.y, n2.z, n2.w};
                    c0 = __float_as_int(n3.x); c1 = __float_as_int(n3.y);
                    float tn0, tn1;
                    bool h0 = slab(lo0, hi0, f.org, inv, s.dist, tn0);
                    bool h1 = slab(lo1, hi1, f.org, inv, s.dist, tn1);
                    l0 = h0 && c0 < 0; l1 = h1 && c1 < 0;
                    h0 = h0 && c0 >= 0; h1 = h1 && c1 >= 0;
                    const bool swap = tn1 < tn0;
                    const int nearc = swap ? c1 : c0, farc = swap ? c0 : c1;
                    const bool both = h0 && h1, none = !h0 && !h1;
                    if (both) { if (sp < WQ_STACK) sh.stack[sp][tid] = farc; else atomicAdd(s.overflow, 1u); }
                    int popped = NODE_NONE;
                    if (none && sp > 0) popped = sh.stack[sp - 1][tid];
                    sp += both ? 1 : 0; sp -= (none && sp > 0) ? 1 : 0;
                    node = both ? nearc : (h0 ? c0 : (h1 ? c1 : popped));
                }
                // ---- enqueue leaf candidates in the warp ring (ballot-ranked slots)
                {
                    const unsigned int m0 = __ballot_sync(FULL, l0), m1 = __ballot_sync(FULL, l1);
                    const unsigned int n0c = __popc(m0);
                    if (l0) { const unsigned int q = pushed + __popc(m0 & lt_mask); ring_prim[q & (WQ_RING - 1)] = (unsigned int)(~c0); ring_owner[q & (WQ_RING - 1)] = lane; my_last = q; have_queued = true; }
                    if (l1) { const unsigned int q = pushed + n0c + __popc(m1 & lt_mask); ring_prim[q & (WQ_RING - 1)] = (unsigned int)(~c1); ring_owner[q & (WQ_RING - 1)] = lane; my_last = q; have_queued = true; }
                    pushed += n0c + __popc(m1);
                }
                // ---- leaf batches: 32 candidates at a time, any lane tests any ray's candidate
                {
                    const bool drained = !have_queued || (int)(tested - my_last) > 0;
                    const bool waiting = ray_active && (ray_hit || node == NODE_NONE) && !drained;
                    const unsigned int wmask = __ballot_sync(FULL, waiting);
                    const unsigned int trav = __ballot_sync(FULL, ray_active && !ray_hit && node != NODE_NONE);
                    unsigned int avail = pushed - tested;
                    bool flush = avail > 0u && (__popc(wmask) >= wait_thr || trav == 0u);
                    __syncwarp();
                    while (avail >= 32u || flush) {
                        const unsigned int nb = min(avail, 32u);
                        bool hit = false; unsigned int owner = 0;
                        if ((unsigned int)lane < nb) {
                            const unsigned int q = (tested + lane) & (WQ_RING - 1);
                            const unsigned int prim = ring_prim[q]; owner = ring_owner[q];
                            const F3 O = f3(sh.ray[warp][0][owner], sh.ray[warp][1][owner], sh.ray[warp][2][owner]);
                            const F3 D = f3(sh.ray[warp][3][owner], sh.ray[warp][4][owner], sh.ray[warp][5][owner]);
                            float tfar = s.dist;
                            hit = prim_hit<false>(sv, prim, O, D, tfar);
                            cnt.prims++;
                        }
                        if (hit) atomicOr(&sh.hitmask[warp], 1u << owner);
                        tested += nb; avail -= nb; flush = false;
                        __syncwarp();
                    }
                    const unsigned int hm = sh.hitmask[warp];
                    if ((hm >> lane) & 1u) ray_hit = true;
                    __syncwarp();
                    if (hm != 0u && lane == 0) sh.hitmask[warp] = 0u;
                }
                // ---- retire rays whose traversal is over and whose candidates are all tested
                {
                    const bool drained = !have_queued || (int)(tested - my_last) > 0;
                    if (ray_active && (ray_hit || node == NODE_NONE) && drained) { ray_active = false; have_result = true; }
                }
                if (__popc(__ballot_sync(FULL, ray_active)) < thr) break;
            }
        }
        flush_counters(cnt, units, counters);
    }
}

// ===========================================================================
// Warp-queue kernel on the compressed 4-wide BVH (k_horizon_wq4).
// Identical scheduling to k_horizon_wq; the node step fetches one 64-byte
// Bvh4Node (four 128-bit loads) and tests four 16-bit-quantised child boxes:
// half the loads and half the dependent steps per ray of the binary tree.
// ===========================================================================
constexpr int W4_STACK = 40;
constexpr int W4_RING = 256;

struct Wq4Shared {
    uint32_t stack[W4_STACK][WQ_THREADS];
    uint2 ring[WQ_WARPS][W4_RING];   // (primitive, owner lane)
    float ray[WQ_WARPS][6][32];      // O.xyz, D.xyz per lane
    unsigned int hitmask[WQ_WARPS];
};

template <int ALG>
__global__ void __launch_bounds__(WQ_THREADS, 6) k_horizon_wq4(SceneView sv, HorizonParams p, Counters* counters,
                                                               unsigned int* tile_counter, int refill_thr, int wait_thr) {
    __shared__ Wq4Shared sh;
    const Search s = make_search(sv, p, counters);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
    const unsigned int FULL = 0xffffffffu, lt_mask = (1u << lane) - 1u;
    const int rows = p.row_end - p.row_begin;
    const int tiles_x = (p.dim_in_1 + 7) >> 3, tiles_y = (rows + 3) >> 2;
    const unsigned int num_tiles = (unsigned int)tiles_x * (unsigned int)tiles_y;
    const bool vec = (p.azim_num & 3) == 0 && ((reinterpret_cast<size_t>(p.hori) & 15) == 0);
    LaneCounters cnt; cnt.rays = cnt.nodes = cnt.prims = 0;
    uint2* ring = sh.ring[warp];
    if (lane == 0) sh.hitmask[warp] = 0u;
    unsigned int pushed = 0, tested = 0;   // warp-uniform FIFO sequence numbers
    __syncwarp();

    while (true) {
        unsigned int tile = 0;
        if (lane == 0) tile = atomicAdd(tile_counter, 1u);
        tile = __shfl_sync(FULL, tile, 0);
        if (tile >= num_tiles) break;
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const int ci = p.row_begin + ty * 4 + (lane >> 3), cj = tx * 8 + (lane & 7);

        LaneSM m; m.phase = 0; m.k = 0; m.cur = m.prev = m.count = m.prev_az = 0; m.lim_up = m.lim_low = m.samp = 0.f;
        Frame f; OutBuf ob; ob.init(nullptr, false);
        bool has_cell = false;
        unsigned int units = 0;
        if (ci < p.row_end && cj < p.dim_in_1) {
            const size_t c = (size_t)ci * p.dim_in_1 + cj;
            float* out = p.hori + c * p.azim_num;
            if (p.mask[c] == 1) {
                const F3 nrm = f3(p.vec_norm[3 * c], p.vec_norm[3 * c + 1], p.vec_norm[3 * c + 2]);
                const F3 nth = f3(p.vec_north[3 * c], p.vec_north[3 * c + 1], p.vec_north[3 * c + 2]);
                const float4 v = sv.vert4[(size_t)(ci + p.offset_0) * sv.W + (cj + p.offset_1)];
                f = make_frame(f3(v.x, v.y, v.z), nrm, nth, p.ray_org_elev);
                ob.init(out, vec);
                has_cell = true; units = p.azim_num;
                sh.ray[warp][0][lane] = f.org.x; sh.ray[warp][1][lane] = f.org.y; sh.ray[warp][2][lane] = f.org.z;
            } else {
                for (int k = 0; k < p.azim_num; ++k) out[k] = p.hori_fill;  // horizon_comp.cpp:789-794
            }
        }

        bool ray_active = false, ray_hit = false, have_result = false;
        float Ax = 0.f, Ay = 0.f, Az = 0.f, Bx = 0.f, By = 0.f, Bz = 0.f;
        unsigned int selnx = 0x7410u, selny = 0x7410u, selnz = 0x7410u;
        uint32_t node = G_NONE; int sp = 0;
        unsigned int my_last = 0; bool have_queued = false;

        while (true) {
            // (1) refill
            if (has_cell && !ray_active) {
                int ie;
                if (sm_advance<ALG>(s, m, have_result, ray_hit, ob, ie)) {
                    const F3 D = ray_dir(s, f, ie, m.k);
                    const RayInv inv = make_inv(D);
                    Ax = sv.qstep[0] * inv.ix; Ay = sv.qstep[1] * inv.iy; Az = sv.qstep[2] * inv.iz;
                    Bx = (sv.qorg[0] - f.org.x) * inv.ix; By = (sv.qorg[1] - f.org.y) * inv.iy; Bz = (sv.qorg[2] - f.org.z) * inv.iz;
                    selnx = inv.ix >= 0.f ? 0x7410u : 0x7432u;
                    selny = inv.iy >= 0.f ? 0x7410u : 0x7432u;
                    selnz = inv.iz >= 0.f ? 0x7410u : 0x7432u;
                    sh.ray[warp][3][lane] = D.x; sh.ray[warp][4][lane] = D.y; sh.ray[warp][5][lane] = D.z;
                    node = 0u; sp = 0; ray_active = true; ray_hit = false; have_queued = false; cnt.rays++;
                } else has_cell = false;
            }
            const unsigned int cell_mask = __ballot_sync(FULL, has_cell);
            if (cell_mask == 0u) break;
            const int thr = min(refill_thr, __popc(cell_mask));
            __syncwarp();

            // (2) traversal
            while (true) {
                const bool do_node = ray_active && !ray_hit && node != G_NONE;
                bool lf0 = false, lf1 = false, lf2 = false, lf3 = false;
                uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
                if (do_node) {
                    const uint4* np = reinterpret_cast<const uint4*>(sv.nodes4 + node);
                    const uint4 r0 = __ldg(np), r1 = __ldg(np + 1), r2 = __ldg(np + 2), r3 = __ldg(np + 3);
                    cnt.nodes++;
                    float t0, t1, t2, t3; bool h0, h1, h2, h3;
                    wide_child_test(r0, selnx, selny, selnz, Ax, Ay, Az, Bx, By, Bz, s.dist, t0, h0);
                    wide_child_test(r1, selnx, selny, selnz, Ax, Ay, Az, Bx, By, Bz, s.dist, t1, h1);
                    wide_child_test(r2, selnx, selny, selnz, Ax, Ay, Az, Bx, By, Bz, s.dist, t2, h2);
                    wide_child_test(r3, selnx, selny, selnz, Ax, Ay, Az, Bx, By, Bz, s.dist, t3, h3);
                    w0 = r0.w; w1 = r1.w; w2 = r2.w; w3 = r3.w;
                    lf0 = h0 && (w0 & WIDE_LEAF); lf1 = h1 && (w1 & WIDE_LEAF); lf2 = h2 && (w2 & WIDE_LEAF); lf3 = h3 && (w3 & WIDE_LEAF);
                    const bool i0 = h0 && !(w0 & WIDE_LEAF), i1 = h1 && !(w1 & WIDE_LEAF), i2 = h2 && !(w2 & WIDE_LEAF), i3 = h3 && !(w3 & WIDE_LEAF);
                    const float k0 = i0 ? t0 : INFINITY, k1 = i1 ? t1 : INFINITY, k2 = i2 ? t2 : INFINITY, k3 = i3 ? t3 : INFINITY;
                    const float kmin = fminf(fminf(k0, k1), fminf(k2, k3));
                    const bool any_int = i0 || i1 || i2 || i3;
                    // nearest internal child first; the others go to the stack
                    const int idx = (i0 && k0 == kmin) ? 0 : ((i1 && k1 == kmin) ? 1 : ((i2 && k2 == kmin) ? 2 : 3));
                    const uint32_t nearest = idx == 0 ? w0 : (idx == 1 ? w1 : (idx == 2 ? w2 : w3));
                    if (sp + 3 > W4_STACK) { if (any_int) atomicAdd(s.overflow, 1u); }
                    else {
                        if (i0 && idx != 0) { sh.stack[sp][tid] = w0; ++sp; }
                        if (i1 && idx != 1) { sh.stack[sp][tid] = w1; ++sp; }
                        if (i2 && idx != 2) { sh.stack[sp][tid] = w2; ++sp; }
                        if (i3 && idx != 3) { sh.stack[sp][tid] = w3; ++sp; }
                    }
                    if (any_int) node = nearest;
                    else if (sp > 0) { --sp; node = sh.stack[sp][tid]; }
                    else node = G_NONE;
                }
                // ---- enqueue leaf candidates in the warp ring (ballot-ranked slots)
                {
                    const unsigned int m0 = __ballot_sync(FULL, lf0), m1 = __ballot_sync(FULL, lf1);
                    const unsigned int m2 = __ballot_sync(FULL, lf2), m3 = __ballot_sync(FULL, lf3);
                    unsigned int base = pushed;
                    if (lf0) { const unsigned int q = base + __popc(m0 & lt_mask); ring[q & (W4_RING - 1)] = make_uint2(w0 & 0x7FFFFFFFu, lane); my_last = q; have_queued = true; }
                    base += __popc(m0);
                    if (lf1) { const unsigned int q = base + __popc(m1 & lt_mask); ring[q & (W4_RING - 1)] = make_uint2(w1 & 0x7FFFFFFFu, lane); my_last = q; have_queued = true; }
                    base += __popc(m1);
                    if (lf2) { const unsigned int q = base + __popc(m2 & lt_mask); ring[q & (W4_RING - 1)] = make_uint2(w2 & 0x7FFFFFFFu, lane); my_last = q; have_queued = true; }
                    base += __popc(m2);
                    if (lf3) { const unsigned int q = base + __popc(m3 & lt_mask); ring[q & (W4_RING - 1)] = make_uint2(w3 & 0x7FFFFFFFu, lane); my_last = q; have_queued = true; }
                    pushed = base + __popc(m3);
                }
                // ---- leaf batches: 32 candidates at a time, any lane tests any ray's candidate
                {
                    const bool drained = !have_queued || (int)(tested - my_last) > 0;
                    const bool waiting = ray_active && (ray_hit || node == G_NONE) && !drained;
                    const unsigned int wmask = __ballot_sync(FULL, waiting);
                    const unsigned int trav = __ballot_sync(FULL, ray_active && !ray_hit && node != G_NONE);
                    unsigned int avail = pushed - tested;
                    bool flush = avail > 0u && (__popc(wmask) >= wait_thr || trav == 0u);
                    __syncwarp();
                    while (avail >= 32u || flush) {
                        const unsigned int nb = min(avail, 32u);
                        bool hit = false; unsigned int owner = 0;
                        if ((unsigned int)lane < nb) {
                            const uint2 e = ring[(tested + lane) & (W4_RING - 1)];
                            owner = e.y;
                            const F3 O = f3(sh.ray[warp][0][owner], sh.ray[warp][1][owner], sh.ray[warp][2][owner]);
                            const F3 D = f3(sh.ray[warp][3][owner], sh.ray[warp][4][owner], sh.ray[warp][5][owner]);
                            float tfar = s.dist;
                            hit = prim_hit<false>(sv, e.x, O, D, tfar);
                            cnt.prims++;
                        }
                        if (hit) atomicOr(&sh.hitmask[warp], 1u << owner);
                        tested += nb; avail -= nb; flush = false;
                        __syncwarp();
                    }
                    const unsigned int hm = sh.hitmask[warp];
                    if ((hm >> lane) & 1u) ray_hit = true;
                    __syncwarp();
                    if (hm != 0u && lane == 0) sh.hitmask[warp] = 0u;
                }
                {
                    const bool drained = !have_queued || (int)(tested - my_last) > 0;
                    if (ray_active && (ray_hit || node == G_NONE) && drained) { ray_active = false; have_result = true; }
                }
                if (__popc(__ballot_sync(FULL, ray_active)) < thr) break;
            }
        }
        flush_counters(cnt, units, counters);
        if (p.row_done) {  // publish: this tile's outputs are complete and visible
            __threadfence_system();
            __syncwarp();
            if (lane == 0) atomicAdd(p.row_done + ty, 1u);
        }
    }
}

// ===========================================================================
// Group kernel (production path): 4 lanes cooperate on one ray over the 4-wide
// quantised BVH.  Lane `sub` of a group owns child `sub` of the current node:
// one 16-byte load, six plane decodes (PRMT + FADD), six FFMA, ballot.  The
// traversal stack and the queue of pending leaf primitives of each ray live in
// shared memory and are shared by the group's lanes; leaf primitives are tested
// two quads (four triangles) at a time, one triangle per lane.  Eight rays per
// warp: versus one ray per lane this cuts L1 tag traffic 4x (ncu r01 showed the
// per-lane kernel bound by L1 wavefronts), shortens the dependent-load chain
// per ray and keeps the four lanes of a ray in lock step by construction.
// Refill and postponed leaf tests work as in k_horizon_sm, at group level.
// ===========================================================================
constexpr int GRP = 4;
constexpr int GK_THREADS = 128;
constexpr int GK_GROUPS = GK_THREADS / GRP;
constexpr int GSTACK = 64;
constexpr int GQUEUE = 12;

struct GroupShared {
    uint32_t stack[GK_GROUPS][GSTACK + 1];
    uint32_t queue[GK_GROUPS][GQUEUE + 1];
    float frame[GK_GROUPS][13];
};

template <int ALG>
__global__ void __launch_bounds__(GK_THREADS, 6) k_horizon_grp(SceneView sv, HorizonParams p, Counters* counters,
                                                               unsigned int* tile_counter, int refill_groups, int leaf_thr) {
    __shared__ GroupShared sh;
    const Search s = make_search(sv, p, counters);
    const int lane = threadIdx.x & 31, sub = lane & 3, gshift = lane & ~3;
    const int grp = threadIdx.x >> 2;
    const unsigned int gmask = 0xFu << gshift, below = (1u << sub) - 1u;
    const unsigned int FULL = 0xffffffffu;
    const int rows = p.row_end - p.row_begin;
    const int tiles_x = (p.dim_in_1 + 7) >> 3, tiles_y = (rows + 3) >> 2;
    const unsigned int num_tiles = (unsigned int)tiles_x * (unsigned int)tiles_y;
    const bool vec = (p.azim_num & 3) == 0 && ((reinterpret_cast<size_t>(p.hori) & 15) == 0);
    const int thr_lanes_cfg = refill_groups * GRP;
    LaneCounters cnt; cnt.rays = cnt.nodes = cnt.prims = 0;
    unsigned int units = 0;
    uint32_t* stack = sh.stack[grp];
    uint32_t* queue = sh.queue[grp];
    float* fr = sh.frame[grp];

    while (true) {
        unsigned int tile = 0;
        if (lane == 0) tile = atomicAdd(tile_counter, 1u);
        tile = __shfl_sync(FULL, tile, 0);
        if (tile >= num_tiles) break;
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        int next_cell = 0;  // warp-uniform: next unassigned cell of the 8x4 tile

        LaneSM m; m.phase = 0; m.k = 0; m.cur = m.prev = m.count = m.prev_az = 0; m.lim_up = m.lim_low = m.samp = 0.f;
        OutBuf ob; ob.init(nullptr, false, false);
        bool has_cell = false, ray_active = false, ray_hit = false, have_result = false;
        F3 O = f3(0.f, 0.f, 0.f), D = f3(0.f, 0.f, 1.f);
        float Ax = 0.f, Ay = 0.f, Az = 0.f, Bx = 0.f, By = 0.f, Bz = 0.f;
        unsigned int selnx = 0x7410u, selny = 0x7410u, selnz = 0x7410u;
        uint32_t cur = G_NONE; int sp = 0, qn = 0;

        while (true) {
            // (0) hand out cells of the tile to groups that have none
            while (true) {
                const bool want = !has_cell;
                const unsigned int wmask = __ballot_sync(FULL, want && sub == 0);
                if (wmask == 0u || next_cell >= 32) break;
                const int my_cell = next_cell + __popc(wmask & ((1u << gshift) - 1u));
                next_cell += __popc(wmask);
                if (want && my_cell < 32) {
                    const int ci = p.row_begin + ty * 4 + (my_cell >> 3), cj = tx * 8 + (my_cell & 7);
                    if (ci < p.row_end && cj < p.dim_in_1) {
                        const size_t c = (size_t)ci * p.dim_in_1 + cj;
                        float* out = p.hori + c * p.azim_num;
                        if (p.mask[c] == 1) {
                            const F3 nrm = f3(p.vec_norm[3 * c], p.vec_norm[3 * c + 1], p.vec_norm[3 * c + 2]);
                            const F3 nth = f3(p.vec_north[3 * c], p.vec_north[3 * c + 1], p.vec_north[3 * c + 2]);
                            const float4 v = sv.vert4[(size_t)(ci + p.offset_0) * sv.W + (cj + p.offset_1)];
                            const Frame f = make_frame(f3(v.x, v.y, v.z), nrm, nth, p.ray_org_elev);
                            O = f.org;
                            if (sub == 0) {
                                fr[0] = f.m00; fr[1] = f.m01; fr[2] = f.m02; fr[3] = f.m10; fr[4] = f.m11; fr[5] = f.m12;
                                fr[6] = f.m20; fr[7] = f.m21; fr[8] = f.m22;
                            }
                            ob.init(out, vec, sub == 0);
                            m.phase = 0; m.k = 0;
                            has_cell = true; have_result = false;
                            units += (sub == 0) ? p.azim_num : 0;
                        } else {
                            for (int k = sub; k < p.azim_num; k += GRP) out[k] = p.hori_fill;  // horizon_comp.cpp:789-794
                        }
                    }
                }
                __syncwarp(FULL);
            }
            // (1) refill: groups with a cell but no ray in flight advance their state machine
            if (has_cell && !ray_active) {
                int ie;
                if (sm_advance<ALG>(s, m, have_result, ray_hit, ob, ie)) {
                    Frame f; f.org = O;
                    f.m00 = fr[0]; f.m01 = fr[1]; f.m02 = fr[2]; f.m10 = fr[3]; f.m11 = fr[4]; f.m12 = fr[5];
                    f.m20 = fr[6]; f.m21 = fr[7]; f.m22 = fr[8];
                    D = ray_dir(s, f, ie, m.k);
                    const RayInv inv = make_inv(D);
                    Ax = sv.qstep[0] * inv.ix; Ay = sv.qstep[1] * inv.iy; Az = sv.qstep[2] * inv.iz;
                    Bx = (sv.qorg[0] - O.x) * inv.ix; By = (sv.qorg[1] - O.y) * inv.iy; Bz = (sv.qorg[2] - O.z) * inv.iz;
                    selnx = inv.ix >= 0.f ? 0x7410u : 0x7432u;
                    selny = inv.iy >= 0.f ? 0x7410u : 0x7432u;
                    selnz = inv.iz >= 0.f ? 0x7410u : 0x7432u;
                    cur = 0u; sp = 0; qn = 0; ray_active = true; cnt.rays += (sub == 0);
                } else has_cell = false;
            }
            const unsigned int cell_mask = __ballot_sync(FULL, has_cell);
            if (cell_mask == 0u) { if (next_cell >= 32) break; else continue; }
            const int thr = min(thr_lanes_cfg, __popc(cell_mask));

            // (2) traversal: every group steps its own ray.  All warp collectives use the
            // full mask and are executed by every lane (sub-warp masks make the compiler
            // serialise the vote per group); inactive groups carry neutral data instead.
            while (true) {
                {
                    const bool do_node = ray_active && cur != G_NONE && qn <= GQUEUE - GRP;
                    uint4 rec = make_uint4(0x0000FFFFu, 0x0000FFFFu, 0x0000FFFFu, WIDE_EMPTY);
                    if (do_node) rec = __ldg(reinterpret_cast<const uint4*>(sv.nodes4 + cur) + sub);
                    const float M = 8388608.0f;
                    const float qnx = __uint_as_float(__byte_perm(rec.x, 0x4B000000u, selnx)) - M;
                    const float qfx = __uint_as_float(__byte_perm(rec.x, 0x4B000000u, selnx ^ 0x0022u)) - M;
                    const float qny = __uint_as_float(__byte_perm(rec.y, 0x4B000000u, selny)) - M;
                    const float qfy = __uint_as_float(__byte_perm(rec.y, 0x4B000000u, selny ^ 0x0022u)) - M;
                    const float qnz = __uint_as_float(__byte_perm(rec.z, 0x4B000000u, selnz)) - M;
                    const float qfz = __uint_as_float(__byte_perm(rec.z, 0x4B000000u, selnz ^ 0x0022u)) - M;
                    const float tmin = fmaxf(fmaxf(fmaf(qnx, Ax, Bx), fmaf(qny, Ay, By)), fmaxf(fmaf(qnz, Az, Bz), 0.0f));
                    const float tmax = fminf(fminf(fmaf(qfx, Ax, Bx), fmaf(qfy, Ay, By)), fminf(fmaf(qfz, Az, Bz), s.dist));
                    const bool hit = (tmin <= tmax * 1.000001f) && (rec.w != WIDE_EMPTY);
                    const bool leaf = (rec.w & WIDE_LEAF) != 0u;
                    const bool hint = hit && !leaf;
                    const unsigned int lb = (__ballot_sync(FULL, hit && leaf) >> gshift) & 0xFu;
                    const unsigned int ib = (__ballot_sync(FULL, hint) >> gshift) & 0xFu;
                    if (hit && leaf) queue[qn + __popc(lb & below)] = rec.w & 0x7FFFFFFFu;
                    qn += __popc(lb);
                    const float key = hint ? tmin : INFINITY;
                    const float k1 = fminf(key, __shfl_xor_sync(FULL, key, 1));
                    const float k2 = fminf(k1, __shfl_xor_sync(FULL, k1, 2));
                    const unsigned int nb = (__ballot_sync(FULL, hint && key == k2) >> gshift) & 0xFu;
                    const int nsub = nb ? (__ffs(nb) - 1) : 0;
                    const uint32_t near_ref = __shfl_sync(FULL, rec.w, gshift + nsub);
                    if (do_node) {
                        if (ib) {
                            const unsigned int others = ib & ~(1u << nsub);
                            const int npush = __popc(others);
                            if (sp + npush > GSTACK) { if (sub == 0) atomicAdd(s.overflow, 1u); }
                            else {
                                if ((others >> sub) & 1u) stack[sp + __popc(others & below)] = rec.w;
                                sp += npush;
                            }
                            cur = near_ref;
                        } else {
                            if (sp > 0) { --sp; cur = stack[sp]; } else cur = G_NONE;
                        }
                        cnt.nodes += (sub == 0);
                    }
                    __syncwarp();
                }
                // leaf step: two queued primitives per group, one triangle per lane
                const bool pending = ray_active && qn > 0;
                const bool must = pending && (qn > GQUEUE - GRP || cur == G_NONE);
                const unsigned int pend_mask = __ballot_sync(FULL, pending);
                if (__any_sync(FULL, must) || __popc(pend_mask) >= leaf_thr) {
                    const int nb = pending ? min(2, qn) : 0;
                    const int pr = sub >> 1, tri = sub & 1;
                    const bool have = pr < nb;
                    const uint32_t prim = have ? queue[qn - 1 - pr] : 0xFFFFFFFFu;
                    const bool isquad = have && prim < sv.num_quads;
                    F3 top = f3(0.f, 0.f, 0.f), bot = top;
                    if (isquad) {
                        const uint32_t wq = (uint32_t)(sv.W - 1);
                        const uint32_t qi = prim / wq, qj = prim - qi * wq;
                        const float4* r0 = sv.vert4 + (size_t)qi * sv.W + qj + tri;
                        top = ld_vert(r0); bot = ld_vert(r0 + sv.W);
                    }
                    // lane tri=0 holds (p00, p10), lane tri=1 holds (p01, p11): swap the shared diagonal
                    const F3 snd = tri ? top : bot;
                    F3 rcv;
                    rcv.x = __shfl_xor_sync(FULL, snd.x, 1); rcv.y = __shfl_xor_sync(FULL, snd.y, 1);
                    rcv.z = __shfl_xor_sync(FULL, snd.z, 1);
                    bool hit = false; float t;
                    if (isquad) {
                        hit = tri ? tri_hit(bot, rcv, top, O, D, s.dist, t)    // (p11, p10, p01)
                                  : tri_hit(top, rcv, bot, O, D, s.dist, t);   // (p00, p01, p10)
                    } else if (have && tri == 0) {
                        const float4* q = sv.tin4 + 3 * (size_t)(prim - sv.num_quads);
                        hit = tri_hit(ld_vert(q), ld_vert(q + 1), ld_vert(q + 2), O, D, s.dist, t);
                    }
                    const unsigned int anyhit = (__ballot_sync(FULL, hit) >> gshift) & 0xFu;
                    qn -= nb; cnt.prims += (sub == 0) ? nb : 0;
                    if (anyhit) { ray_active = false; ray_hit = true; have_result = true; }
                }
                if (ray_active && cur == G_NONE && qn == 0) { ray_active = false; ray_hit = false; have_result = true; }
                if (__popc(__ballot_sync(FULL, ray_active)) < thr) break;
            }
        }
    }
    flush_counters(cnt, units, counters);
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

int launch_horizon_gridded(Scene& s, const HorizonParams& p, cudaStream_t st) {
    if (p.row_end <= p.row_begin || p.dim_in_1 <= 0) return 0;
    HZB_CUDA(cudaMemsetAsync(s.d_tile_counter, 0, sizeof(unsigned int), st));
    const SceneView sv = s.view();
    static const char* kern_env = getenv("HZB_KERNEL");      // "simple" selects the reference-shaped kernel
    static const int refill_thr = getenv("HZB_REFILL") ? atoi(getenv("HZB_REFILL")) : 24;
    static const int leaf_thr = getenv("HZB_LEAF") ? atoi(getenv("HZB_LEAF")) : 12;
    if (kern_env && !strcmp(kern_env, "simple")) {
        const int grid = sm_count() * 8;
        switch (p.algorithm) {
            case 0: k_horizon_gridded<0><<<grid, HG_THREADS, 0, st>>>(sv, p, s.d_counters, s.d_tile_counter); break;
            case 1: k_horizon_gridded<1><<<grid, HG_THREADS, 0, st>>>(sv, p, s.d_counters, s.d_tile_counter); break;
            default: k_horizon_gridded<2><<<grid, HG_THREADS, 0, st>>>(sv, p, s.d_counters, s.d_tile_counter); break;
        }
    } else if (kern_env && !strcmp(kern_env, "sm")) {
        const int grid = sm_count() * SM_MINB;
        switch (p.algorithm) {
            case 0: k_horizon_sm<0><<<grid, SM_THREADS, 0, st>>>(sv, p, s.d_counters, s.d_tile_counter, refill_thr, leaf_thr); break;
            case 1: k_horizon_sm<1><<<grid, SM_THREADS, 0, st>>>(sv, p, s.d_counters, s.d_tile_counter, refill_thr, leaf_thr); break;
            default: k_horizon_sm<2><<<grid, SM_THREADS, 0, st>>>(sv, p, s.d_counters, s.d_tile_counter, refill_thr, leaf_thr); break;
        }
    } else if (!kern_env || !strcmp(kern_env, "wq4")) {
        static const int w_refill = getenv("HZB_WREFILL") ? atoi(getenv("HZB_WREFILL")) : 24;
        static const int w_wait = getenv("HZB_WWAIT") ? atoi(getenv("HZB_WWAIT")) : 6;
        const int grid = sm_count() * 6;
        switch (p.algorithm) {
            case 0: k_horizon_wq4<0><<<grid, WQ_THREADS, 0, st>>>(sv, p, s.d_counters, s.d_tile_counter, w_refill, w_wait); break;
            case 1: k_horizon_wq4<1><<<grid, WQ_THREADS, 0, st>>>(sv, p, s.d_counters, s.d_tile_counter, w_refill, w_wait); break;
            default: k_horizon_wq4<2><<<grid, WQ_THREADS, 0, st>>>(sv, p, s.d_counters, s.d_tile_counter, w_refill, w_wait); break;
        }
    } else if (!strcmp(kern_env, "wq")) {
        static const int w_refill = getenv("HZB_WREFILL") ? atoi(getenv("HZB_WREFILL")) : 24;
        static const int w_wait = getenv("HZB_WWAIT") ? atoi(getenv("HZB_WWAIT")) : 6;
        const int grid = sm_count() * 6;
        switch (p.algorithm) {
            case 0: k_horizon_wq<0><<<grid, WQ_THREADS, 0, st>>>(sv, p, s.d_counters, s.d_tile_counter, w_refill, w_wait); break;
            case 1: k_horizon_wq<1><<<grid, WQ_THREADS, 0, st>>>(sv, p, s.d_counters, s.d_tile_counter, w_refill, w_wait); break;
            default: k_horizon_wq<2><<<grid, WQ_THREADS, 0, st>>>(sv, p, s.d_counters, s.d_tile_counter, w_refill, w_wait); break;
        }
    } else {
        static const int g_refill = getenv("HZB_GREFILL") ? atoi(getenv("HZB_GREFILL")) : 6;   // groups (of 8)
        static const int g_leaf = getenv("HZB_GLEAF") ? atoi(getenv("HZB_GLEAF")) : 16;        // lanes (of 32)
        static const int g_ctas = getenv("HZB_GCTAS") ? atoi(getenv("HZB_GCTAS")) : 6;
        const int grid = sm_count() * g_ctas;
        switch (p.algorithm) {
            case 0: k_horizon_grp<0><<<grid, GK_THREADS, 0, st>>>(sv, p, s.d_counters, s.d_tile_counter, g_refill, g_leaf); break;
            case 1: k_horizon_grp<1><<<grid, GK_THREADS, 0, st>>>(sv, p, s.d_counters, s.d_tile_counter, g_refill, g_leaf); break;
            default: k_horizon_grp<2><<<grid, GK_THREADS, 0, st>>>(sv, p, s.d_counters, s.d_tile_counter, g_refill, g_leaf); break;
        }
    }
    HZB_CUDA(cudaGetLastError());
    return 0;
}

int launch_horizon_locations(Scene& s, const HorizonParams& p, const LocationParams& lp, cudaStream_t st) {
    if (lp.num_loc <= 0) return 0;
    float4* d_org = nullptr;
    HZB_CUDA(cudaMalloc((void**)&d_org, (size_t)lp.num_loc * sizeof(float4)));
    const SceneView sv = s.view();
    const int nb_loc = (lp.num_loc + 63) / 64;
    k_loc_snap<<<nb_loc, 64, 0, st>>>(sv, lp, d_org, s.d_counters);
    if (p.algorithm == 2) {
        k_loc_chain<<<(lp.num_loc + 31) / 32, 32, 0, st>>>(sv, p, lp, d_org, s.d_counters);
    } else {
        const long long total = (long long)lp.num_loc * p.azim_num;
        const int nb = (int)((total + 127) / 128);
        if (p.algorithm == 0) {
            if (lp.hori_dist_out) k_loc_indep<0, true><<<nb, 128, 0, st>>>(sv, p, lp, d_org, s.d_counters);
            else k_loc_indep<0, false><<<nb, 128, 0, st>>>(sv, p, lp, d_org, s.d_counters);
        } else {
            if (lp.hori_dist_out) k_loc_indep<1, true><<<nb, 128, 0, st>>>(sv, p, lp, d_org, s.d_counters);
            else k_loc_indep<1, false><<<nb, 128, 0, st>>>(sv, p, lp, d_org, s.d_counters);
        }
        if (lp.hori_dist_out) k_loc_dist_fix<<<nb_loc, 64, 0, st>>>(lp, p.azim_num, d_org);
    }
    HZB_CUDA(cudaGetLastError());
    HZB_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_org);
    return 0;
}

}  // namespace hzb
