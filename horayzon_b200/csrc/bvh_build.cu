// bvh_build.cu -- on-device scene construction.
//
// Replaces Embree's rtcNewScene / rtcCommitScene as called by the reference
// (horizon_comp.cpp:101-231, shadow_comp.cpp:198-298) with a B200-resident
// build: vertices -> float4, one primitive per grid quad (two triangles, split
// along (i,j+1)-(i+1,j)) plus one per TIN triangle, 63-bit Morton codes of the
// primitive box centres, hand-written LSD radix sort (8-bit digits), Karras
// binary radix tree, bottom-up box fit with conservative padding, and collapse
// into an 8-wide quantised BVH laid out breadth-first.
#include "hzb_common.cuh"
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include <chrono>

namespace hzb {

namespace {

// ------------------------------------------------------------------ helpers
__device__ __forceinline__ unsigned int f2ord(float f) {  // order-preserving float -> uint
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(unsigned int u) {
    unsigned int v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(v);
#else
    float f; memcpy(&f, &v, 4); return f;
#endif
}

struct PrimGeom {
    const float4* vert4; const float4* tin4; int W; uint32_t num_quads;
};

__device__ __forceinline__ void prim_box(const PrimGeom& g, uint32_t prim, float* lo, float* hi) {
    float4 a, b, c;
    if (prim < g.num_quads) {
        const uint32_t wq = (uint32_t)(g.W - 1);
        const uint32_t i = prim / wq, j = prim - i * wq;
        const float4* r0 = g.vert4 + (size_t)i * g.W + j;
        const float4* r1 = r0 + g.W;
        a = r0[0]; b = r0[1]; c = r1[0];
        const float4 d = r1[1];
        lo[0] = fminf(fminf(a.x, b.x), fminf(c.x, d.x)); hi[0] = fmaxf(fmaxf(a.x, b.x), fmaxf(c.x, d.x));
        lo[1] = fminf(fminf(a.y, b.y), fminf(c.y, d.y)); hi[1] = fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y));
        lo[2] = fminf(fminf(a.z, b.z), fminf(c.z, d.z)); hi[2] = fmaxf(fmaxf(a.z, b.z), fmaxf(c.z, d.z));
    } else {
        const float4* q = g.tin4 + 3 * (size_t)(prim - g.num_quads);
        a = q[0]; b = q[1]; c = q[2];
        lo[0] = fminf(fminf(a.x, b.x), c.x); hi[0] = fmaxf(fmaxf(a.x, b.x), c.x);
        lo[1] = fminf(fminf(a.y, b.y), c.y); hi[1] = fmaxf(fmaxf(a.y, b.y), c.y);
        lo[2] = fminf(fminf(a.z, b.z), c.z); hi[2] = fmaxf(fmaxf(a.z, b.z), c.z);
    }
}

// ------------------------------------------------------- vertex conversion
__global__ void k_vert_to_float4(const float* __restrict__ v3, float4* __restrict__ v4, size_t n,
                                 unsigned int* __restrict__ bounds /*[6] ord-encoded lo xyz, hi xyz*/) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float x = v3[3 * i], y = v3[3 * i + 1], z = v3[3 * i + 2];
        v4[i] = make_float4(x, y, z, 0.f);
        lo[0] = fminf(lo[0], x); hi[0] = fmaxf(hi[0], x);
        lo[1] = fminf(lo[1], y); hi[1] = fmaxf(hi[1], y);
        lo[2] = fminf(lo[2], z); hi[2] = fmaxf(hi[2], z);
    }
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
        for (int a = 0; a < 3; ++a) {
            atomicMin(&bounds[a], f2ord(lo[a]));
            atomicMax(&bounds[3 + a], f2ord(hi[a]));
        }
    }
}

__global__ void k_tin_gather(const float* __restrict__ vs, const int32_t* __restrict__ idx, float4* __restrict__ tin4,
                             uint32_t num_tin, unsigned int* __restrict__ bounds) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_tin) return;
    for (int c = 0; c < 3; ++c) {
        const float* q = vs + 3 * (size_t)idx[3 * t + c];
        tin4[3 * (size_t)t + c] = make_float4(q[0], q[1], q[2], 0.f);
        atomicMin(&bounds[0], f2ord(q[0])); atomicMax(&bounds[3], f2ord(q[0]));
        atomicMin(&bounds[1], f2ord(q[1])); atomicMax(&bounds[4], f2ord(q[1]));
        atomicMin(&bounds[2], f2ord(q[2])); atomicMax(&bounds[5], f2ord(q[2]));
    }
}

// ----------------------------------------------------------- Morton codes
__device__ __forceinline__ unsigned long long expand21(unsigned int v) {
    unsigned long long x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__device__ __forceinline__ unsigned long long expand31(unsigned int v) {   // spread 31 bits to even positions
    unsigned long long x = v & 0x7fffffffu;
    x = (x | x << 16) & 0x0000ffff0000ffffull;
    x = (x | x << 8) & 0x00ff00ff00ff00ffull;
    x = (x | x << 4) & 0x0f0f0f0f0f0f0f0full;
    x = (x | x << 2) & 0x3333333333333333ull;
    x = (x | x << 1) & 0x5555555555555555ull;
    return x;
}

// mode 0: 3-D Morton code (21 bits per axis) of the box centre in the scene cube.
// mode 1: 2-D Morton code (31 bits for x and y): for height fields the radix tree becomes a
//         quadtree in the horizontal plane, which collapses into fuller 4-wide nodes.
// mode 2: grid quads are keyed on their integer cell indices (i, j): the quadtree is aligned
//         with the grid (2^k x 2^k blocks of cells), every 4-wide node is full.  Used when the
//         scene has no TIN (TIN triangles have no cell index; then mode 1 is used for all).
__global__ void k_morton(PrimGeom g, uint32_t n, float lox, float loy, float loz, float inv_extent, int mode,
                         unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    float lo[3], hi[3];
    prim_box(g, p, lo, hi);
    if (mode == 2 && p < g.num_quads) {   // grid quad: key on the integer cell indices -> aligned quadtree
        const uint32_t wq = (uint32_t)(g.W - 1);
        const uint32_t qi = p / wq, qj = p - qi * wq;
        keys[p] = (expand31(qj) << 1) | expand31(qi);
        vals[p] = p;
        return;
    }
    if (mode >= 1) {
        const double s2 = 2147483648.0 * (double)inv_extent;  // 2^31 / square edge
        const double fx2 = (0.5 * ((double)lo[0] + hi[0]) - lox) * s2, fy2 = (0.5 * ((double)lo[1] + hi[1]) - loy) * s2;
        const unsigned int qx2 = (unsigned int)fmin(fmax(fx2, 0.0), 2147483647.0);
        const unsigned int qy2 = (unsigned int)fmin(fmax(fy2, 0.0), 2147483647.0);
        keys[p] = (expand31(qx2) << 1) | expand31(qy2);
        vals[p] = p;
        return;
    }
    const float s = 2097152.0f * inv_extent;  // 2^21 / cube edge
    const float fx = (0.5f * (lo[0] + hi[0]) - lox) * s;
    const float fy = (0.5f * (lo[1] + hi[1]) - loy) * s;
    const float fz = (0.5f * (lo[2] + hi[2]) - loz) * s;
    const unsigned int qx = (unsigned int)fminf(fmaxf(fx, 0.f), 2097151.f);
    const unsigned int qy = (unsigned int)fminf(fmaxf(fy, 0.f), 2097151.f);
    const unsigned int qz = (unsigned int)fminf(fmaxf(fz, 0.f), 2097151.f);
    keys[p] = (expand21(qx) << 2) | (expand21(qy) << 1) | expand21(qz);
    vals[p] = p;
}

// --------------------------------------------------------------- radix sort
// LSD radix sort, 8-bit digits, 64-bit keys + 32-bit values.  Per pass:
// per-tile digit histogram -> exclusive scan of the [digit][tile] matrix ->
// stable scatter (warp match + cross-warp prefix).
constexpr int RS_THREADS = 256;
constexpr int RS_ROUNDS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const unsigned long long* __restrict__ keys, uint32_t n,
                                                        int shift, uint32_t* __restrict__ hist, uint32_t ntiles) {
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * RS_TILE;
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const size_t idx = base + (size_t)r * RS_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&sh[(unsigned int)(keys[idx] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * ntiles + blockIdx.x] = sh[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const unsigned long long* __restrict__ kin,
                                                           const uint32_t* __restrict__ vin,
                                                           unsigned long long* __restrict__ kout,
                                                           uint32_t* __restrict__ vout, uint32_t n, int shift,
                                                           const uint32_t* __restrict__ offs, uint32_t ntiles) {
    __shared__ uint32_t running[256];
    __shared__ uint32_t wcnt[RS_THREADS / 32][256];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    running[tid] = offs[(size_t)tid * ntiles + blockIdx.x];
#pragma unroll
    for (int w = 0; w < RS_THREADS / 32; ++w) wcnt[w][tid] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * RS_TILE;
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const size_t idx = base + (size_t)r * RS_THREADS + tid;
        const bool valid = idx < n;
        unsigned long long key = 0; uint32_t val = 0; unsigned int digit = 256u;
        if (valid) { key = kin[idx]; val = vin[idx]; digit = (unsigned int)(key >> shift) & 255u; }
        const unsigned int peers = __match_any_sync(0xffffffffu, digit);
        const unsigned int rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) wcnt[warp][digit] = __popc(peers);
        __syncthreads();
        if (valid) {
            uint32_t pos = running[digit] + rank;
            for (int w = 0; w < warp; ++w) pos += wcnt[w][digit];
            kout[pos] = key; vout[pos] = val;
        }
        __syncthreads();
        uint32_t tot = 0;
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; ++w) { tot += wcnt[w][tid]; wcnt[w][tid] = 0; }
        running[tid] += tot;
        __syncthreads();
    }
}

// generic exclusive scan of uint32 (reduce / scan sums / apply), 4096 per block
constexpr int SC_THREADS = 256, SC_PER = 16, SC_TILE = SC_THREADS * SC_PER;

__global__ void __launch_bounds__(SC_THREADS) k_scan_reduce(const uint32_t* __restrict__ in, size_t m, uint32_t* __restrict__ sums) {
    __shared__ uint32_t sh[SC_THREADS / 32];
    const size_t base = (size_t)blockIdx.x * SC_TILE + (size_t)threadIdx.x * SC_PER;
    uint32_t s = 0;
    for (int k = 0; k < SC_PER; ++k) if (base + k < m) s += in[base + k];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < SC_THREADS / 32; ++w) t += sh[w]; sums[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(1024) k_scan_sums(uint32_t* sums, uint32_t nb) {  // single block, in place exclusive
    __shared__ uint32_t sh[1024];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = (i < nb) ? sums[i] : 0u;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            uint32_t t = (threadIdx.x >= (unsigned)o) ? sh[threadIdx.x - o] : 0u;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nb) sums[i] = carry + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += sh[1023];
        __syncthreads();
    }
}
__global__ void __launch_bounds__(SC_THREADS) k_scan_apply(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                           size_t m, const uint32_t* __restrict__ sums) {
    __shared__ uint32_t sh[SC_THREADS];
    const size_t base = (size_t)blockIdx.x * SC_TILE + (size_t)threadIdx.x * SC_PER;
    uint32_t v[SC_PER]; uint32_t s = 0;
    for (int k = 0; k < SC_PER; ++k) { v[k] = (base + k < m) ? in[base + k] : 0u; s += v[k]; }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < SC_THREADS; o <<= 1) {
        uint32_t t = (threadIdx.x >= (unsigned)o) ? sh[threadIdx.x - o] : 0u;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    uint32_t run = sums[blockIdx.x] + sh[threadIdx.x] - s;
    for (int k = 0; k < SC_PER; ++k) { if (base + k < m) out[base + k] = run; run += v[k]; }
}

// ------------------------------------------------------- Karras radix tree
__device__ __forceinline__ int delta(const unsigned long long* __restrict__ keys, long long n, long long i, long long j) {
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((unsigned int)(i ^ j));
    return __clzll((long long)(a ^ b));
}

// parent links: value = (parent index << 1) | side (0 left, 1 right)
__global__ void k_karras(const unsigned long long* __restrict__ keys, uint32_t n, Bvh2Node* __restrict__ nodes,
                         uint32_t* __restrict__ node_parent, uint32_t* __restrict__ leaf_parent) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)n - 1) return;
    const long long N = n;
    const int d = (delta(keys, N, i, i + 1) - delta(keys, N, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, N, i, i - d);
    long long lmax = 2;
    while (delta(keys, N, i, i + lmax * d) > dmin) lmax <<= 1;
    long long l = 0;
    for (long long t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, N, i, i + (l + t) * d) > dmin) l += t;
    const long long j = i + l * d;
    const int dnode = delta(keys, N, i, j);
    long long s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, N, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const long long gamma = i + s * d + (d < 0 ? -1 : 0);
    const long long lo = i < j ? i : j, hi = i < j ? j : i;
    const uint32_t me = (uint32_t)i;
    if (lo == gamma) { nodes[i].c0 = ~(int)gamma; leaf_parent[gamma] = (me << 1); }
    else { nodes[i].c0 = (int)gamma; node_parent[gamma] = (me << 1); }
    if (hi == gamma + 1) { nodes[i].c1 = ~(int)(gamma + 1); leaf_parent[gamma + 1] = (me << 1) | 1u; }
    else { nodes[i].c1 = (int)(gamma + 1); node_parent[gamma + 1] = (me << 1) | 1u; }
    if (i == 0) node_parent[0] = 0xffffffffu;
}

// bottom-up box fit; leaves are padded, unions inherit the padding
__global__ void k_refit(PrimGeom g, const uint32_t* __restrict__ sorted_prims, uint32_t n, float pad,
                        Bvh2Node* nodes, const uint32_t* __restrict__ node_parent,
                        const uint32_t* __restrict__ leaf_parent, unsigned int* visit, float* root_box) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float lo[3], hi[3];
    const uint32_t prim = sorted_prims[i];
    prim_box(g, prim, lo, hi);
    for (int a = 0; a < 3; ++a) { lo[a] -= pad; hi[a] += pad; }
    if (n == 1) {  // degenerate scene: node 0 = (leaf, empty)
        Bvh2Node nd;
        for (int a = 0; a < 3; ++a) { nd.lo0[a] = lo[a]; nd.hi0[a] = hi[a]; nd.lo1[a] = INFINITY; nd.hi1[a] = -INFINITY; }
        nd.c0 = ~(int)prim; nd.c1 = ~(int)prim; nd.pad0 = nd.pad1 = 0;
        nodes[0] = nd;
        for (int a = 0; a < 3; ++a) { root_box[a] = lo[a]; root_box[3 + a] = hi[a]; }
        return;
    }
    uint32_t link = leaf_parent[i];
    int code = ~(int)prim;
    int height = 0;   // height of the subtree carried upwards (leaf = 0); stored per child in pad0 / pad1
    while (true) {
        const uint32_t p = link >> 1, side = link & 1u;
        Bvh2Node* nd = nodes + p;
        float* blo = side ? nd->lo1 : nd->lo0;
        float* bhi = side ? nd->hi1 : nd->hi0;
        for (int a = 0; a < 3; ++a) { blo[a] = lo[a]; bhi[a] = hi[a]; }
        if (side) { nd->c1 = code; nd->pad1 = height; } else { nd->c0 = code; nd->pad0 = height; }
        __threadfence();
        if (atomicAdd(&visit[p], 1u) == 0u) return;
        const volatile float* slo = side ? nd->lo0 : nd->lo1;
        const volatile float* shi = side ? nd->hi0 : nd->hi1;
        for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], slo[a]); hi[a] = fmaxf(hi[a], shi[a]); }
        const volatile int* sh = side ? &nd->pad0 : &nd->pad1;
        height = max(height, *sh) + 1;
        code = (int)p;
        link = node_parent[p];
        if (link == 0xffffffffu) {
            for (int a = 0; a < 3; ++a) { root_box[a] = lo[a]; root_box[3 + a] = hi[a]; }
            return;
        }
    }
}

template <typename T>
int dalloc(T** p, size_t n) {
    *p = (T*)pool_alloc(std::max<size_t>(n, 1) * sizeof(T));     // pooled: a repeated call allocates nothing
    return *p ? 0 : 1;
}

int exclusive_scan_u32(uint32_t* d_in, uint32_t* d_out, size_t m, uint32_t* d_sums, cudaStream_t st) {
    const uint32_t nb = (uint32_t)((m + SC_TILE - 1) / SC_TILE);
    k_scan_reduce<<<nb, SC_THREADS, 0, st>>>(d_in, m, d_sums);
    k_scan_sums<<<1, 1024, 0, st>>>(d_sums, nb);
    k_scan_apply<<<nb, SC_THREADS, 0, st>>>(d_in, d_out, m, d_sums);
    HZB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace


void scene_free(Scene& s) {
    pool_free(s.d_vert4); pool_free(s.d_tin4); pool_free(s.d_nodes2); pool_free(s.d_nodes4);
    pool_free(s.d_prim_ids); pool_free(s.d_counters); pool_free(s.d_tile_counter);
    for (Scene::TableEntry& e : s.tables) pool_free(e.d);
    s.tables.clear();
    s.d_vert4 = nullptr; s.d_tin4 = nullptr; s.d_nodes2 = nullptr; s.d_nodes4 = nullptr; s.d_prim_ids = nullptr;
    s.d_counters = nullptr; s.d_tile_counter = nullptr;
}

int scene_upload_and_build(Scene& s, const float* vert_grid, int H, int W, const float* vert_simp,
                           int num_vert_simp, const int32_t* tri_ind_simp, int num_tri_simp) {
    if (H < 2 || W < 2) { set_error("DEM must be at least 2 x 2 vertices"); return 1; }
    cudaStream_t st = 0;
    s.H = H; s.W = W;
    s.num_quads = (uint32_t)(H - 1) * (uint32_t)(W - 1);
    s.num_tin = (num_vert_simp >= 3 && num_tri_simp > 0) ? (uint32_t)num_tri_simp : 0u;  // horizon_comp.cpp:199
    s.num_prims = s.num_quads + s.num_tin;
    const size_t nv = (size_t)H * W;
    const uint32_t n = s.num_prims;

    // build temporaries: released on every exit path
    struct Tmp {
        std::vector<void**> slots;
        void own(void** p) { slots.push_back(p); }
        ~Tmp() { for (void** p : slots) if (*p) { pool_free(*p); *p = nullptr; } }
    } tmp;
#define HZB_TMP(ptr) tmp.own((void**)&(ptr))

    // ---- H2D
    double t0 = now_s();
    float* d_v3 = nullptr; float* d_vs = nullptr; int32_t* d_ti = nullptr; unsigned int* d_bounds = nullptr;
    HZB_TMP(d_v3); HZB_TMP(d_vs); HZB_TMP(d_ti); HZB_TMP(d_bounds);
    HZB_TRY(dalloc(&d_v3, nv * 3));
    HZB_TRY(dalloc(&s.d_vert4, nv));
    HZB_TRY(dalloc(&d_bounds, 6));
    HZB_TRY(dalloc(&s.d_counters, 1));
    HZB_TRY(dalloc(&s.d_tile_counter, HZB_TILE_SLOTS));
    HZB_CUDA(cudaMemsetAsync(s.d_counters, 0, sizeof(Counters), st));
    HZB_CUDA(cudaMemcpyAsync(d_v3, vert_grid, nv * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    if (s.num_tin) {
        HZB_TRY(dalloc(&d_vs, (size_t)num_vert_simp * 3));
        HZB_TRY(dalloc(&d_ti, (size_t)num_tri_simp * 3));
        HZB_TRY(dalloc(&s.d_tin4, (size_t)s.num_tin * 3));
        HZB_CUDA(cudaMemcpyAsync(d_vs, vert_simp, (size_t)num_vert_simp * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
        HZB_CUDA(cudaMemcpyAsync(d_ti, tri_ind_simp, (size_t)num_tri_simp * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    }
    HZB_CUDA(cudaStreamSynchronize(st));
    s.t_h2d = now_s() - t0;

    // ---- build
    t0 = now_s();
    const unsigned int init_bounds[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    HZB_CUDA(cudaMemcpyAsync(d_bounds, init_bounds, sizeof(init_bounds), cudaMemcpyHostToDevice, st));
    k_vert_to_float4<<<sm_count() * 8, 256, 0, st>>>(d_v3, s.d_vert4, nv, d_bounds);
    if (s.num_tin) k_tin_gather<<<(s.num_tin + 255) / 256, 256, 0, st>>>(d_vs, d_ti, s.d_tin4, s.num_tin, d_bounds);
    unsigned int hb[6];
    HZB_CUDA(cudaMemcpyAsync(hb, d_bounds, sizeof(hb), cudaMemcpyDeviceToHost, st));
    HZB_CUDA(cudaStreamSynchronize(st));
    float extent = 0.f, scale = 0.f;
    for (int a = 0; a < 3; ++a) {
        s.lo[a] = ord2f(hb[a]); s.hi[a] = ord2f(hb[3 + a]);
        extent = std::max(extent, s.hi[a] - s.lo[a]);
        scale = std::max(scale, std::max(std::max(fabsf(s.lo[a]), fabsf(s.hi[a])), s.hi[a] - s.lo[a]));
    }
    if (!(extent > 0.f) || !std::isfinite(scale)) { set_error("degenerate or non-finite DEM vertices"); return 1; }
    // conservative padding: a few tens of ulps of the scene scale
    s.pad = scale * 4.0e-6f;

    PrimGeom g{s.d_vert4, s.d_tin4, W, s.num_quads};
    unsigned long long *d_k0 = nullptr, *d_k1 = nullptr; uint32_t *d_v0 = nullptr, *d_v1 = nullptr;
    HZB_TMP(d_k0); HZB_TMP(d_k1); HZB_TMP(d_v0); HZB_TMP(d_v1);
    HZB_TRY(dalloc(&d_k0, n)); HZB_TRY(dalloc(&d_k1, n)); HZB_TRY(dalloc(&d_v0, n)); HZB_TRY(dalloc(&d_v1, n));
    int morton_mode = getenv("HZB_MORTON") ? atoi(getenv("HZB_MORTON")) : 2;
    if (morton_mode == 2 && s.num_tin > 0) morton_mode = 1;
    const float extent_key = morton_mode >= 1 ? std::max(s.hi[0] - s.lo[0], s.hi[1] - s.lo[1]) : extent;
    k_morton<<<(n + 255) / 256, 256, 0, st>>>(g, n, s.lo[0], s.lo[1], s.lo[2], 1.0f / std::max(extent_key, 1e-20f), morton_mode, d_k0, d_v0);
    HZB_CUDA(cudaGetLastError());

    const uint32_t ntiles = (n + RS_TILE - 1) / RS_TILE;
    const size_t m = (size_t)256 * ntiles;
    uint32_t *d_hist = nullptr, *d_offs = nullptr, *d_sums = nullptr;
    HZB_TMP(d_hist); HZB_TMP(d_offs); HZB_TMP(d_sums);
    HZB_TRY(dalloc(&d_hist, m)); HZB_TRY(dalloc(&d_offs, m)); HZB_TRY(dalloc(&d_sums, (m + SC_TILE - 1) / SC_TILE + 1));
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 8 * pass;
        k_rs_hist<<<ntiles, RS_THREADS, 0, st>>>(d_k0, n, shift, d_hist, ntiles);
        HZB_TRY(exclusive_scan_u32(d_hist, d_offs, m, d_sums, st));
        k_rs_scatter<<<ntiles, RS_THREADS, 0, st>>>(d_k0, d_v0, d_k1, d_v1, n, shift, d_offs, ntiles);
        std::swap(d_k0, d_k1); std::swap(d_v0, d_v1);
    }
    HZB_CUDA(cudaGetLastError());
    // sorted: d_k0 (keys), d_v0 (primitive ids)
    s.d_prim_ids = d_v0; d_v0 = nullptr;

    const uint32_t n_int = n > 1 ? n - 1 : 1;
    uint32_t *d_np = nullptr, *d_lp = nullptr; unsigned int* d_visit = nullptr; float* d_root = nullptr;
    HZB_TMP(d_np); HZB_TMP(d_lp); HZB_TMP(d_visit); HZB_TMP(d_root);
    HZB_TRY(dalloc(&s.d_nodes2, n_int)); HZB_TRY(dalloc(&d_np, n_int)); HZB_TRY(dalloc(&d_lp, n));
    HZB_TRY(dalloc(&d_visit, n_int)); HZB_TRY(dalloc(&d_root, 6));
    HZB_CUDA(cudaMemsetAsync(d_visit, 0, (size_t)n_int * sizeof(unsigned int), st));
    HZB_CUDA(cudaMemsetAsync(s.d_nodes2, 0, (size_t)n_int * sizeof(Bvh2Node), st));
    if (n > 1) k_karras<<<(n - 1 + 255) / 256, 256, 0, st>>>(d_k0, n, s.d_nodes2, d_np, d_lp);
    k_refit<<<(n + 255) / 256, 256, 0, st>>>(g, s.d_prim_ids, n, s.pad, s.d_nodes2, d_np, d_lp, d_visit, d_root);
    HZB_CUDA(cudaGetLastError());
    HZB_CUDA(cudaStreamSynchronize(st));
    HZB_TRY(build_wide_bvh(s, st));
    s.bvh_bytes = (size_t)s.num_nodes4 * sizeof(Bvh4Node);

#undef HZB_TMP
    HZB_CUDA(cudaStreamSynchronize(st));
    s.t_build = now_s() - t0;
    return 0;   // ~Tmp frees the build temporaries
}

}  // namespace hzb
