// hzb_wq.cuh -- FIRST-GENERATION warp-queue traversal step (compressed 4-wide BVH, one ray per
// lane), superseded in production by hzb_wq2.cuh; still selectable (HZB_KERNEL=wq5,
// HZB_SHADOW_KERNEL=wq1) and parity-tested.  Shared constants (block size, stack depth) live here.
//
// One call = one iteration of the warp's traversal loop:
//   1. node step, executed by ALL lanes (lanes without a traversing ray read the
//      root and are masked), so every warp collective below sits in converged
//      code: one 64-byte Bvh4Node = four 128-bit loads, four quantised slab tests,
//      nearest hit child next, the other hit children onto the shared-memory stack
//      ([entry][thread]: conflict free), no branches;
//   2. leaf hits -> warp ring buffer, slot = ballot rank, entry = (primitive, owner);
//   3. when 32 candidates are queued (or lanes are waiting on theirs) every lane
//      tests one candidate against its OWNER's ray (mirrored in shared memory);
//   4. a ray retires when its traversal is over (or it was hit) and the FIFO has
//      passed its last candidate.
// Decisions are those of tri_hit on exactly the primitives whose (conservative)
// boxes the ray meets; only order and executing lane differ from a scalar loop.
#pragma once
#include "hzb_geom.cuh"

namespace hzb {

constexpr int WQ_BLOCK = 128;                 // threads per CTA
constexpr int WQ_NWARPS = WQ_BLOCK / 32;
constexpr int WQ_STACK_N = 36;
constexpr int WQ_RING_N = 256;
constexpr uint32_t WQ_NONE = 0xFFFFFFFFu;

constexpr int WQ_TOP_NODES = 85;              // levels 0..3 of a full 4-ary tree (optional smem copy)

struct WqShared {
    uint32_t stack[WQ_STACK_N][WQ_BLOCK];
    uint2 ring[WQ_NWARPS][WQ_RING_N];         // (primitive, owner lane)
    float ray[WQ_NWARPS][6][32];              // O.xyz, D.xyz per lane
    unsigned int hitmask[WQ_NWARPS];
};

struct WqLane {          // per-lane ray state
    int state;           // 0 no ray, 1 traversing, 2 traversal over (or hit), waiting for its candidates
    bool hit, queued;
    uint32_t node; int sp;
    unsigned int my_last;
    float Ax, Ay, Az, Bx, By, Bz;
    unsigned int selnx, selny, selnz;
};

struct WqWarp { unsigned int pushed, tested; };   // warp-uniform FIFO sequence numbers

__device__ __forceinline__ void wq_start_ray(const SceneView& sv, WqShared& sh, int warp, int lane, WqLane& L, F3 O, F3 D) {
    const RayInv inv = make_inv(D);
    L.Ax = sv.qstep[0] * inv.ix; L.Ay = sv.qstep[1] * inv.iy; L.Az = sv.qstep[2] * inv.iz;
    L.Bx = (sv.qorg[0] - O.x) * inv.ix; L.By = (sv.qorg[1] - O.y) * inv.iy; L.Bz = (sv.qorg[2] - O.z) * inv.iz;
    L.selnx = inv.ix >= 0.f ? 0x7410u : 0x7432u;
    L.selny = inv.iy >= 0.f ? 0x7410u : 0x7432u;
    L.selnz = inv.iz >= 0.f ? 0x7410u : 0x7432u;
    sh.ray[warp][0][lane] = O.x; sh.ray[warp][1][lane] = O.y; sh.ray[warp][2][lane] = O.z;
    sh.ray[warp][3][lane] = D.x; sh.ray[warp][4][lane] = D.y; sh.ray[warp][5][lane] = D.z;
    L.node = 0u; L.sp = 0; L.state = 1; L.hit = false; L.queued = false;
}

// Optional: stage the first `nodes` entries of the breadth-first node array (the top
// levels, walked by every ray) into shared memory with one TMA bulk copy
// (cp.async.bulk + mbarrier).  Measured on B200: 3 % SLOWER than leaving them to L1
// (generic loads replace LDG.CONSTANT), so it is off by default (HZB_TOPSMEM=1).
__device__ __forceinline__ void wq_tma_stage_top(uint4* dst, const Bvh4Node* src, unsigned int nodes, unsigned long long* mbar) {
    const unsigned int bytes = nodes * (unsigned int)sizeof(Bvh4Node);
    const unsigned int mbar_s = (unsigned int)__cvta_generic_to_shared(mbar);
    const unsigned int dst_s = (unsigned int)__cvta_generic_to_shared(dst);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_s), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst_s), "l"(src), "r"(bytes), "r"(mbar_s) : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "HZB_WQ_TMA_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
        "@p bra HZB_WQ_TMA_DONE;\n\t"
        "bra HZB_WQ_TMA_WAIT;\n\t"
        "HZB_WQ_TMA_DONE:\n\t}" ::"r"(mbar_s) : "memory");
}

// returns the ballot of lanes that still own a ray after this step
template <bool TOPS>
__device__ __forceinline__ unsigned int wq_step(const SceneView& sv, WqShared& sh, const uint4* top_nodes, const unsigned int n_top,
                                                const int warp, const int lane, const int tid,
                                                WqLane& L, WqWarp& W, const float tfar, const int wait_thr,
                                                LaneCounters& cnt, unsigned int* overflow) {
    const unsigned int FULL = 0xffffffffu, lt_mask = (1u << lane) - 1u;
    uint2* ring = sh.ring[warp];
    // ---- 1. node step (all lanes)
    const bool trav = L.state == 1;
    const uint32_t nidx = trav ? L.node : 0u;
    uint4 r0, r1, r2, r3;
    if (TOPS) {   // generic loads: shared memory for the staged top levels, global otherwise
        const uint4* np = (nidx < n_top) ? (top_nodes + 4u * nidx) : reinterpret_cast<const uint4*>(sv.nodes4 + nidx);
        r0 = np[0]; r1 = np[1]; r2 = np[2]; r3 = np[3];
    } else {
        const uint4* np = reinterpret_cast<const uint4*>(sv.nodes4 + nidx);
        r0 = __ldg(np); r1 = __ldg(np + 1); r2 = __ldg(np + 2); r3 = __ldg(np + 3);
    }
    float t0, t1, t2, t3; bool h0, h1, h2, h3;
    wide_child_test(r0, L.selnx, L.selny, L.selnz, L.Ax, L.Ay, L.Az, L.Bx, L.By, L.Bz, tfar, t0, h0);
    wide_child_test(r1, L.selnx, L.selny, L.selnz, L.Ax, L.Ay, L.Az, L.Bx, L.By, L.Bz, tfar, t1, h1);
    wide_child_test(r2, L.selnx, L.selny, L.selnz, L.Ax, L.Ay, L.Az, L.Bx, L.By, L.Bz, tfar, t2, h2);
    wide_child_test(r3, L.selnx, L.selny, L.selnz, L.Ax, L.Ay, L.Az, L.Bx, L.By, L.Bz, tfar, t3, h3);
    cnt.nodes += trav ? 1u : 0u;
    const unsigned int hm = trav ? ((h0 ? 1u : 0u) | (h1 ? 2u : 0u) | (h2 ? 4u : 0u) | (h3 ? 8u : 0u)) : 0u;
    const unsigned int lfm = (r0.w >> 31) | ((r1.w >> 31) << 1) | ((r2.w >> 31) << 2) | ((r3.w >> 31) << 3);
    const unsigned int lm = hm & lfm;
    const unsigned int im = hm & ~lfm;
    {   // nearest internal child next, the rest onto the stack
        const float k0 = (im & 1u) ? t0 : INFINITY, k1 = (im & 2u) ? t1 : INFINITY;
        const float k2 = (im & 4u) ? t2 : INFINITY, k3 = (im & 8u) ? t3 : INFINITY;
        const float kmin = fminf(fminf(k0, k1), fminf(k2, k3));
        const unsigned int eq = ((k0 == kmin) ? 1u : 0u) | ((k1 == kmin) ? 2u : 0u) | ((k2 == kmin) ? 4u : 0u) | ((k3 == kmin) ? 8u : 0u);
        const unsigned int first = (im & eq) & (0u - (im & eq));     // lowest set bit (0 if no internal hit)
        const uint32_t nearest = (first & 1u) ? r0.w : ((first & 2u) ? r1.w : ((first & 4u) ? r2.w : r3.w));
        const unsigned int others = im & ~first;
        int sp = L.sp;
        if (sp + 3 > WQ_STACK_N) { if (others) atomicAdd(overflow, 1u); }
        else {
            if (others & 1u) { sh.stack[sp][tid] = r0.w; ++sp; }
            if (others & 2u) { sh.stack[sp][tid] = r1.w; ++sp; }
            if (others & 4u) { sh.stack[sp][tid] = r2.w; ++sp; }
            if (others & 8u) { sh.stack[sp][tid] = r3.w; ++sp; }
        }
        if (trav) {
            uint32_t next = nearest;
            if (!im) {
                next = WQ_NONE;
                if (sp > 0) { --sp; next = sh.stack[sp][tid]; }
            }
            L.sp = sp; L.node = next;
            if (next == WQ_NONE) L.state = 2;
        }
    }
    // ---- 2. leaf hits -> warp ring: slot = exclusive prefix of the per-lane hit counts
    //         (count <= 4: three ballots, one per bit of the count), straight-line code
    {
        const unsigned int nl = __popc(lm);
        const unsigned int c0 = __ballot_sync(FULL, nl & 1u), c1 = __ballot_sync(FULL, nl & 2u), c2 = __ballot_sync(FULL, nl & 4u);
        unsigned int q = W.pushed + __popc(c0 & lt_mask) + 2u * __popc(c1 & lt_mask) + 4u * __popc(c2 & lt_mask);
        if (lm & 1u) { ring[q & (WQ_RING_N - 1)] = make_uint2(r0.w & 0x7FFFFFFFu, (unsigned int)lane); ++q; }
        if (lm & 2u) { ring[q & (WQ_RING_N - 1)] = make_uint2(r1.w & 0x7FFFFFFFu, (unsigned int)lane); ++q; }
        if (lm & 4u) { ring[q & (WQ_RING_N - 1)] = make_uint2(r2.w & 0x7FFFFFFFu, (unsigned int)lane); ++q; }
        if (lm & 8u) { ring[q & (WQ_RING_N - 1)] = make_uint2(r3.w & 0x7FFFFFFFu, (unsigned int)lane); ++q; }
        if (lm) { L.my_last = q - 1u; L.queued = true; }
        W.pushed += __popc(c0) + 2u * __popc(c1) + 4u * __popc(c2);
    }
    // ---- 3. leaf batches
    {
        const bool drained = !L.queued || (int)(W.tested - L.my_last) > 0;
        const unsigned int b_wait = __ballot_sync(FULL, L.state == 2 && !drained);
        const unsigned int b_trav = __ballot_sync(FULL, L.state == 1);
        unsigned int avail = W.pushed - W.tested;
        const bool flush = avail > 0u && (__popc(b_wait) >= wait_thr || b_trav == 0u);
        if (avail >= 32u || flush) {
            __syncwarp();
            do {
                const unsigned int nb = min(avail, 32u);
                bool hit = false; unsigned int owner = 0;
                if ((unsigned int)lane < nb) {
                    const uint2 e = ring[(W.tested + lane) & (WQ_RING_N - 1)];
                    owner = e.y;
                    const F3 O = f3(sh.ray[warp][0][owner], sh.ray[warp][1][owner], sh.ray[warp][2][owner]);
                    const F3 D = f3(sh.ray[warp][3][owner], sh.ray[warp][4][owner], sh.ray[warp][5][owner]);
                    float tf = tfar;
                    hit = prim_hit<false>(sv, e.x, O, D, tf);
                    cnt.prims++;
                }
                if (hit) atomicOr(&sh.hitmask[warp], 1u << owner);
                W.tested += nb; avail -= nb;
            } while (avail >= 32u);
            __syncwarp();
            const unsigned int hmask = sh.hitmask[warp];
            if ((hmask >> lane) & 1u) { L.hit = true; if (L.state == 1) L.state = 2; }
            __syncwarp();
            if (hmask != 0u && lane == 0) sh.hitmask[warp] = 0u;
        }
    }
    // ---- 4. retire
    {
        const bool drained = !L.queued || (int)(W.tested - L.my_last) > 0;
        if (L.state == 2 && drained) L.state = 0;
    }
    return __ballot_sync(FULL, L.state != 0);
}

}  // namespace hzb
