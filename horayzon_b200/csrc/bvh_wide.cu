// bvh_wide.cu -- collapse of the binary LBVH into the compressed 4-wide BVH.
//
// Level by level from the root (so the node array is in breadth-first order and
// the top levels are one contiguous block that the traversal kernel stages into
// shared memory with a TMA bulk copy): every wide node starts from one binary
// node, then twice replaces its largest-area internal child by that child's two
// children.  Child boxes (already padded at build time) are quantised
// conservatively to 16 bits on a scene-global grid.
#include "hzb_common.cuh"
#include <math.h>
#include <algorithm>

namespace hzb {
namespace {

struct QGrid { double org[3], inv_step[3]; };

__device__ __forceinline__ float box_area(const float* lo, const float* hi) {
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return dx * dy + dy * dz + dz * dx;
}

__device__ __forceinline__ uint32_t quant_pair(float lo, float hi, double org, double inv_step) {
    // one extra quantum on each side: the traversal folds the 2^23 decode bias into the ray
    // constant (hzb_wq2.cuh), which costs up to half a quantum of accuracy
    double ql = floor(((double)lo - org) * inv_step) - 1.0;
    double qh = ceil(((double)hi - org) * inv_step) + 1.0;
    ql = fmin(fmax(ql, 0.0), 65535.0);
    qh = fmin(fmax(qh, 0.0), 65535.0);
    return (uint32_t)ql | ((uint32_t)qh << 16);
}

__global__ void k_collapse4(const Bvh2Node* __restrict__ nodes2, const uint32_t* __restrict__ frontier, uint32_t count,
                            uint32_t out_base, Bvh4Node* __restrict__ nodes4, uint32_t* __restrict__ next_frontier,
                            unsigned int* __restrict__ next_count, uint32_t next_base, QGrid g) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    int code[4], hgt[4]; float lo[4][3], hi[4][3];
    int n = 2;
    {
        const Bvh2Node nd = nodes2[frontier[t]];
        code[0] = nd.c0; code[1] = nd.c1; hgt[0] = nd.pad0; hgt[1] = nd.pad1;
        for (int a = 0; a < 3; ++a) { lo[0][a] = nd.lo0[a]; hi[0][a] = nd.hi0[a]; lo[1][a] = nd.lo1[a]; hi[1][a] = nd.hi1[a]; }
        if (nd.c0 == nd.c1 && nd.c0 < 0) n = 1;  // degenerate single-primitive scene
    }
    // Absorb up to two binary nodes.  Only children of ODD height are absorbed, so that
    // wide nodes sit at even heights above the leaves: the bottom wide level then holds
    // four leaves per node instead of stranding two-leaf nodes whenever a subtree has an
    // odd number of binary levels (1200 x 1200 quads: 2.6 -> ~4 children per node).
    // Among the candidates the largest box goes first.
    for (int round = 0; round < 2 && n < 4; ++round) {
        int best = -1; float best_area = -1.f;
        for (int k = 0; k < n; ++k)
            if (code[k] >= 0 && (hgt[k] & 1)) { const float ar = box_area(lo[k], hi[k]); if (ar > best_area) { best_area = ar; best = k; } }
        if (best < 0) break;
        const Bvh2Node nd = nodes2[code[best]];
        code[best] = nd.c0; code[n] = nd.c1; hgt[best] = nd.pad0; hgt[n] = nd.pad1;
        for (int a = 0; a < 3; ++a) { lo[best][a] = nd.lo0[a]; hi[best][a] = nd.hi0[a]; lo[n][a] = nd.lo1[a]; hi[n][a] = nd.hi1[a]; }
        ++n;
    }
    int n_int = 0;
    for (int k = 0; k < n; ++k) n_int += (code[k] >= 0);
    uint32_t slot = 0;
    if (n_int) slot = atomicAdd(next_count, (unsigned int)n_int);
    Bvh4Node out;
    for (int k = 0; k < 4; ++k) {
        WideChild c;
        if (k < n) {
            c.qx = quant_pair(lo[k][0], hi[k][0], g.org[0], g.inv_step[0]);
            c.qy = quant_pair(lo[k][1], hi[k][1], g.org[1], g.inv_step[1]);
            c.qz = quant_pair(lo[k][2], hi[k][2], g.org[2], g.inv_step[2]);
            if (code[k] >= 0) { next_frontier[slot] = (uint32_t)code[k]; c.ref = next_base + slot; ++slot; }
            else c.ref = WIDE_LEAF | (uint32_t)(~code[k]);
        } else { c.qx = c.qy = c.qz = 0x0000FFFFu; c.ref = WIDE_EMPTY; }  // lo = 65535, hi = 0: never hit
        out.c[k] = c;
    }
    nodes4[out_base + t] = out;
}

}  // namespace

int build_wide_bvh(Scene& s, cudaStream_t st) {
    const uint32_t n = s.num_prims;
    const uint32_t cap = n > 1 ? n - 1 : 1;  // a wide node absorbs >= 1 binary node
    // quantisation grid: scene box (plus padding and three steps of margin) over 65535 steps
    QGrid g;
    for (int a = 0; a < 3; ++a) {
        const double lo = (double)s.lo[a] - 2.0 * (double)s.pad, hi = (double)s.hi[a] + 2.0 * (double)s.pad;
        const double step = std::max(hi - lo, 1e-6) / 65529.0;
        s.qstep[a] = (float)step;
        s.qorg[a] = (float)(lo - 3.0 * step);
        // the traversal decodes with the float values: quantise against exactly those
        g.org[a] = (double)s.qorg[a];
        g.inv_step[a] = 1.0 / (double)s.qstep[a];
    }
    uint32_t *d_f0 = nullptr, *d_f1 = nullptr; unsigned int* d_cnt = nullptr;
    s.d_nodes4 = (Bvh4Node*)pool_alloc((size_t)cap * sizeof(Bvh4Node));
    d_f0 = (uint32_t*)pool_alloc((size_t)cap * sizeof(uint32_t));
    d_f1 = (uint32_t*)pool_alloc((size_t)cap * sizeof(uint32_t));
    d_cnt = (unsigned int*)pool_alloc(sizeof(unsigned int));
    struct TmpGuard { uint32_t*& a; uint32_t*& b; unsigned int*& c; ~TmpGuard() { pool_free(a); pool_free(b); pool_free(c); a = b = nullptr; c = nullptr; } } tguard{d_f0, d_f1, d_cnt};
    if (!s.d_nodes4 || !d_f0 || !d_f1 || !d_cnt) return 1;
    const uint32_t root = 0;
    HZB_CUDA(cudaMemcpyAsync(d_f0, &root, sizeof(root), cudaMemcpyHostToDevice, st));
    uint32_t count = 1, base = 0;
    int levels = 0;
    while (count > 0) {
        if ((size_t)base + count > cap) { set_error("wide BVH overflow (internal error)"); return 1; }
        HZB_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned int), st));
        k_collapse4<<<(count + 127) / 128, 128, 0, st>>>(s.d_nodes2, d_f0, count, base, s.d_nodes4, d_f1, d_cnt, base + count, g);
        unsigned int next = 0;
        HZB_CUDA(cudaMemcpyAsync(&next, d_cnt, sizeof(next), cudaMemcpyDeviceToHost, st));
        HZB_CUDA(cudaStreamSynchronize(st));
        base += count; count = next;
        std::swap(d_f0, d_f1);
        if (++levels > 4096) { set_error("wide BVH too deep"); return 1; }
    }
    s.num_nodes4 = base;
    // the node array was sized for the worst case (one wide node per binary node); a quadtree-shaped BVH needs a third
    // of that: give the rest back when it is worth a copy (576 M quads: 36.9 GB -> 12.3 GB)
    if ((size_t)cap > (size_t)base + base / 2 && (size_t)base * sizeof(Bvh4Node) > ((size_t)256 << 20)) {
        Bvh4Node* exact = (Bvh4Node*)pool_alloc((size_t)base * sizeof(Bvh4Node));
        if (exact) {
            HZB_CUDA(cudaMemcpyAsync(exact, s.d_nodes4, (size_t)base * sizeof(Bvh4Node), cudaMemcpyDeviceToDevice, st));
            HZB_CUDA(cudaStreamSynchronize(st));
            pool_free(s.d_nodes4);
            s.d_nodes4 = exact;
        }
    }
    HZB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace hzb
