// hzb_wq2.cuh -- second-generation warp-queue traversal step: TWO rays per lane.
//
// The guess_constant search (horizon_comp.cpp:387-498) needs, for almost every
// azimuth, the decisions of the two casts at table index prev+5 and prev-5.  They
// share the origin and the azimuth and differ by 0.5 deg in elevation, so their
// paths through the BVH are nearly the same: this step walks both with ONE
// traversal (a node is entered if either ray meets its box, a leaf candidate is
// tested against both rays).  Every decision is still tri_hit() of that ray on a
// primitive whose conservative boxes the ray meets -- bit-identical to two
// separate casts; only the number of node fetches, decodes and control
// instructions is shared.  A single ray is a packet whose rays coincide.
//
// Differences to hzb_wq.cuh (kept for the shadow kernel and for A/B runs):
//   * the 2^23 bias of the quantised planes is folded into the ray constant
//     (t = qb*A + B', B' = fma(-2^23, A, B)): one FFMA per plane, no FADD.  The
//     fold costs at most half a quantum of accuracy, which is why bvh_wide.cu
//     widens every quantised box by one quantum on each side;
//   * x/y planes are picked by the (shared) direction signs, z by min/max, so the
//     two rays may point to different sides of the horizontal;
//   * leaf candidates go to a per-lane pending list (plain shared-memory stores);
//     the warp-wide compaction (prefix scan of the list lengths, redux.or of the
//     start bits) runs once per flush instead of three ballots per step;
//   * the two triangles of a grid quad share their diagonal: the Pluecker edge
//     function of the reversed edge is the exact negative, so five edge functions
//     serve both triangles, and their cross products serve both rays.
#pragma once
#include "hzb_geom.cuh"
#include "hzb_wq.cuh"
#include "hzb_box.cuh"

namespace hzb {

constexpr int WQ2_PEND_N = 12;   // per-lane pending leaf candidates (a flush is forced above PEND_N - 4)

struct Wq2Shared {
    uint32_t stack[WQ_STACK_N + 3][WQ_BLOCK];     // three spare rows: a step pushes at most three entries
    uint32_t pend[WQ2_PEND_N][WQ_BLOCK];
    float ray[WQ_NWARPS][9][32];              // O.xyz, D1.xyz, D2.xyz per lane
    unsigned int hit1[WQ_NWARPS], hit2[WQ_NWARPS];
    unsigned int rank_owner[WQ_NWARPS][32];
    unsigned int tbest[WQ_NWARPS][32];        // closest-hit mode: bits of the nearest hit distance per lane (t >= 0: unsigned order = float order)
};

// Start the packet (D1, D2).  Returns false when the two rays cannot share the
// plane selectors (different x or y direction signs): ray 2 is then a copy of ray 1.
__device__ __forceinline__ bool wq2_start(const SceneView& sv, Wq2Shared& sh, int warp, int lane, Wq2Lane& L, F3 O, F3 D1, F3 D2) {
    const bool same = wq2_set_rays(L, sv.qorg, sv.qstep, O, D1, D2);      // hzb_box.cuh
    float* r = &sh.ray[warp][0][lane];
    r[0] = O.x; r[32] = O.y; r[64] = O.z;
    r[96] = D1.x; r[128] = D1.y; r[160] = D1.z;
    r[192] = D2.x; r[224] = D2.y; r[256] = D2.z;
    L.node = 0u; L.sp = 0; L.pc = 0; L.state = 1; L.hit1 = false; L.hit2 = false;
    return same;
}

constexpr uint32_t WQ_OVF = 0xFFFFFFFEu;   // L.node of a packet that retired because its stack was full
// A lane whose packet retired with WQ_OVF abandons its cell and marks the cell's first output element with
// this pattern; a fix-up kernel launched right behind (k_horizon_redo / k_terrain<.., REDO>) recomputes the
// marked cells with the per-lane walker of the binary BVH (96-entry local stack), which tests the same
// primitives with the same tri_hit -- identical results.  (No call inside the persistent loop: a device
// function call there makes the compiler guard every warp collective with a divergence check.)
constexpr uint32_t HZB_REDO_F32 = 0x7FC0DEADu;   // a quiet NaN with a payload no computation produces
constexpr uint8_t HZB_REDO_U8 = 0xFFu;           // shadow codes are 0..3

// Same decisions as prim_hit<false>(.., D1, ..) and (TWO) prim_hit<false>(.., D2, ..).
template <bool TWO>
__device__ __forceinline__ void prim_hit2(const SceneView& s, uint32_t prim, F3 O, F3 D1, F3 D2, float tfar, bool& h1, bool& h2) {
    h1 = false; h2 = false;
    if (prim < s.num_quads) {
        const uint32_t wq = (uint32_t)(s.W - 1);
        const uint32_t i = prim / wq, j = prim - i * wq;
        const float4* r0 = s.vert4 + (size_t)i * s.W + j;
        const float4* r1 = r0 + s.W;
        if (TWO) {
            quad_hit2<true>(ld_vert(r0), ld_vert(r0 + 1), ld_vert(r1), ld_vert(r1 + 1), O, D1, D2, tfar, h1, h2);   // hzb_tri.cuh
        } else {
            // single-ray form (shadow kernels), kept textually as validated on the GPU; to be folded into
            // quad_hit2<false> once a GPU run can confirm the regenerated code
            const F3 a = sub_rn(ld_vert(r0), O), b = sub_rn(ld_vert(r0 + 1), O), c = sub_rn(ld_vert(r1), O), d = sub_rn(ld_vert(r1 + 1), O);
            const F3 e0 = sub_rn(c, a), e1 = sub_rn(a, b), e2 = sub_rn(b, c);
            const F3 f0 = sub_rn(b, d), f1 = sub_rn(d, c);
            const F3 C0 = cross_f(e0, add_rn(c, a)), C1 = cross_f(e1, add_rn(a, b)), C2 = cross_f(e2, add_rn(b, c));
            const F3 G0 = cross_f(f0, add_rn(b, d)), G1 = cross_f(f1, add_rn(d, c));
            const float W1 = dot_f(C2, D1);
            const bool a11 = edges_accept(dot_f(C0, D1), dot_f(C1, D1), W1);
            const bool a21 = edges_accept(dot_f(G0, D1), dot_f(G1, D1), -W1);
            bool a12 = false, a22 = false;
            unsigned int acc = (a11 ? 1u : 0u) | (a12 ? 2u : 0u) | (a21 ? 4u : 0u) | (a22 ? 8u : 0u);
            while (acc) {
                const bool second = (acc & 3u) == 0u;
                const unsigned int pair = second ? (acc >> 2) : (acc & 3u);
                const F3 g0 = second ? f0 : e0, g1 = second ? f1 : e1;
                const F3 g2 = second ? f3(-e2.x, -e2.y, -e2.z) : e2;
                const F3 v0 = second ? d : a;
                const F3 Ng = tri_ng(g0, g1, g2);
                if ((pair & 1u) && !h1) h1 = depth_ok(v0, Ng, D1, tfar);
                if ((pair & 2u) && !h2) h2 = depth_ok(v0, Ng, D2, tfar);
                acc &= second ? 0u : 0xCu;
            }
        }
    } else {
        const float4* q = s.tin4 + 3 * (size_t)(prim - s.num_quads);
        const F3 p0 = ld_vert(q), p1 = ld_vert(q + 1), p2 = ld_vert(q + 2);
        float t;
        h1 = tri_hit(p0, p1, p2, O, D1, tfar, t);
        if (TWO) h2 = tri_hit(p0, p1, p2, O, D2, tfar, t);
    }
}

// One iteration of the warp's traversal loop.  pend_est: warp-uniform estimate of
// the number of pending candidates.  Returns the ballot of lanes that still own a packet.
// TWO: both rays of the packet are live (horizon search); otherwise ray 1 only (shadow).
// SORT: nearest hit child first (pays off for single any-hit rays that are often occluded).
//
// A full stack never invalidates a result: the lane stops its walk, drops its pending candidates and
// retires with L.node == WQ_OVF; the caller marks the cell for the fix-up kernel (see HZB_REDO_F32).
// stack_lim <= WQ_STACK_N (the tests lower it to force that path).
// CLOSEST (single-ray mode only; castRay_intersect1, horizon_comp.cpp:268-292): no early exit on a hit; every
// candidate is tested for its distance, the minimum per lane is kept in sh.tbest and in L.tfar, which culls the
// boxes beyond it.  The result is the minimum of tri_hit's t over all triangles the ray meets -- the same value
// in any traversal order.  The caller sets L.tfar and sh.tbest[warp][lane] to the ray's tfar at the start.
template <bool TWO, bool SORT, bool CLOSEST = false>
__device__ __forceinline__ unsigned int wq2_step(const SceneView& sv, Wq2Shared& sh, const int warp, const int lane, const int tid,
                                                 Wq2Lane& L, unsigned int& pend_est, const float tfar_in, const int wait_thr,
                                                 LaneCounters& cnt, const int stack_lim) {
    static_assert(!CLOSEST || !TWO, "closest-hit mode is single-ray");
    const float tfar = CLOSEST ? L.tfar : tfar_in;
    const unsigned int FULL = 0xffffffffu, lt_mask = (1u << lane) - 1u;
    // ---- 1. node step (all lanes; lanes without a traversing packet read the root and are masked)
    const bool trav = L.state == 1;
    const uint32_t nidx = trav ? L.node : 0u;
    const uint4* np = reinterpret_cast<const uint4*>(sv.nodes4 + nidx);
    const uint4 r0 = __ldg(np), r1 = __ldg(np + 1), r2 = __ldg(np + 2), r3 = __ldg(np + 3);
    float t0, t1, t2, t3;
    const bool h0 = wide2_child_test<TWO>(r0, L, tfar, t0), h1 = wide2_child_test<TWO>(r1, L, tfar, t1);
    const bool h2 = wide2_child_test<TWO>(r2, L, tfar, t2), h3 = wide2_child_test<TWO>(r3, L, tfar, t3);
    cnt.nodes += trav ? 1u : 0u;
    bool lf0, lf1, lf2, lf3;      // leaf children met by the packet -> pending list
    if (SORT) {
        const unsigned int hm = trav ? ((h0 ? 1u : 0u) | (h1 ? 2u : 0u) | (h2 ? 4u : 0u) | (h3 ? 8u : 0u)) : 0u;
        const unsigned int lfm = (r0.w >> 31) | ((r1.w >> 31) << 1) | ((r2.w >> 31) << 2) | ((r3.w >> 31) << 3);
        const unsigned int lm = hm & lfm;
        const unsigned int im = hm & ~lfm;
        lf0 = lm & 1u; lf1 = lm & 2u; lf2 = lm & 4u; lf3 = lm & 8u;
        // nearest hit internal child next, the rest onto the stack
        const float k0 = (im & 1u) ? t0 : INFINITY, k1 = (im & 2u) ? t1 : INFINITY;
        const float k2 = (im & 4u) ? t2 : INFINITY, k3 = (im & 8u) ? t3 : INFINITY;
        const float kmin = fminf(fminf(k0, k1), fminf(k2, k3));
        const unsigned int eq = ((k0 == kmin) ? 1u : 0u) | ((k1 == kmin) ? 2u : 0u) | ((k2 == kmin) ? 4u : 0u) | ((k3 == kmin) ? 8u : 0u);
        const unsigned int first = (im & eq) & (0u - (im & eq));     // lowest set bit (0 if no internal hit)
        const uint32_t firstc = (first & 1u) ? r0.w : ((first & 2u) ? r1.w : ((first & 4u) ? r2.w : r3.w));
        const unsigned int others = im & ~first;
        int sp = L.sp;
        const bool full = others != 0u && sp + 3 > stack_lim;
        if (!full) {
            if (others & 1u) { sh.stack[sp][tid] = r0.w; ++sp; }
            if (others & 2u) { sh.stack[sp][tid] = r1.w; ++sp; }
            if (others & 4u) { sh.stack[sp][tid] = r2.w; ++sp; }
            if (others & 8u) { sh.stack[sp][tid] = r3.w; ++sp; }
        }
        if (trav) {
            uint32_t next = firstc;
            if (!im) {
                next = WQ_NONE;
                if (sp > 0) { --sp; next = sh.stack[sp][tid]; }
            }
            L.sp = sp; L.node = next;
            if (next == WQ_NONE) L.state = 2;
            if (full) { L.node = WQ_OVF; L.state = 2; L.pc = 0; lf0 = lf1 = lf2 = lf3 = false; }
        }
    } else {
        // Children in slot order with three running values: the next node (first hit internal child), the
        // stack pointer (further internal hits) and the pending-list length (leaf hits).  No distance order:
        // the upper ray of a packet usually misses and has to visit every box it meets anyway (measured on
        // B200: the nearest-first selection cost 9 % and visited MORE nodes).  The stack has three spare
        // rows, so a full stack is detected once per step, not per push.
        int sp = L.sp, pc = L.pc;
        uint32_t next = WQ_NONE;
#define HZB_WQ2_CHILD(RK, HK)                                                      \
        {                                                                          \
            const bool a_ = trav && (HK);                                          \
            const bool lf_ = a_ && ((int)(RK).w < 0);      /* leaf flag = bit 31 */   \
            const bool in_ = a_ && ((int)(RK).w >= 0);                             \
            const bool push_ = in_ && (next != WQ_NONE);                           \
            const bool first_ = in_ && (next == WQ_NONE);                          \
            if (lf_) { sh.pend[pc][tid] = (RK).w; ++pc; }   /* the tester strips the flag */ \
            if (push_) { sh.stack[sp][tid] = (RK).w; ++sp; }                       \
            if (first_) next = (RK).w;                                             \
        }
        HZB_WQ2_CHILD(r0, h0) HZB_WQ2_CHILD(r1, h1) HZB_WQ2_CHILD(r2, h2) HZB_WQ2_CHILD(r3, h3)
#undef HZB_WQ2_CHILD
        lf0 = pc != L.pc; lf1 = lf2 = lf3 = false;
        L.pc = pc;
        if (trav) {
            const bool full = sp > stack_lim;            // the three spare rows took this step's pushes
            if (next == WQ_NONE && sp > 0) { --sp; next = sh.stack[sp][tid]; }
            L.sp = sp; L.node = next;
            if (next == WQ_NONE) L.state = 2;
            if (full) { L.node = WQ_OVF; L.state = 2; L.pc = 0; lf0 = false; }   // decided by the caller's fallback
        }
    }
    // ---- 2. leaf hits -> the lane's own pending list (room for 4 is guaranteed by the flush rule); the
    //         leaf flag (bit 31) stays on, the tester strips it
    {
        if (SORT) {
            int pc = L.pc;
            if (lf0) { sh.pend[pc][tid] = r0.w; ++pc; }
            if (lf1) { sh.pend[pc][tid] = r1.w; ++pc; }
            if (lf2) { sh.pend[pc][tid] = r2.w; ++pc; }
            if (lf3) { sh.pend[pc][tid] = r3.w; ++pc; }
            L.pc = pc;
        }
        pend_est += __popc(__ballot_sync(FULL, lf0 || lf1 || lf2 || lf3));   // lower bound: one per lane with new candidates
    }
    // ---- 3. leaf batches
    {
        const unsigned int b_wait = __ballot_sync(FULL, L.state == 2 && L.pc > 0);
        const unsigned int b_trav = __ballot_sync(FULL, L.state == 1);
        const bool any_full = __any_sync(FULL, L.pc > WQ2_PEND_N - 4);
        const bool forced = any_full || (pend_est > 0u && (__popc(b_wait) >= wait_thr || b_trav == 0u));
        if (pend_est >= 32u || forced) {
            // exclusive prefix of the list lengths: scan position p of lane o's list is its entry pc-1-(p-excl)
            unsigned int incl = (unsigned int)L.pc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            const unsigned int excl = incl - (unsigned int)L.pc;
            const unsigned int total = __shfl_sync(FULL, incl, 31);
            unsigned int base = 0;
            __syncwarp();
            while (total - base >= 32u || (forced && total > base)) {
                const unsigned int nb = min(total - base, 32u);
                const bool contrib = L.pc > 0 && incl > base && excl < base + 32u;
                const unsigned int s0 = (excl > base ? excl : base) - base;
                const unsigned int startmask = __reduce_or_sync(FULL, contrib ? (1u << s0) : 0u);
                const unsigned int cm = __ballot_sync(FULL, contrib);
                if (contrib) sh.rank_owner[warp][__popc(cm & lt_mask)] = (unsigned int)lane;
                __syncwarp();
                unsigned int owner = 0;
                if ((unsigned int)lane < nb) owner = sh.rank_owner[warp][__popc(startmask & (FULL >> (31 - lane))) - 1];
                const unsigned int ex_o = __shfl_sync(FULL, excl, owner);
                const unsigned int pc_o = __shfl_sync(FULL, (unsigned int)L.pc, owner);
                bool q1 = false, q2 = false;
                if ((unsigned int)lane < nb) {
                    const unsigned int k = base + (unsigned int)lane - ex_o;
                    const uint32_t prim = sh.pend[pc_o - 1u - k][(tid & ~31) + owner] & 0x7FFFFFFFu;
                    const float* r = &sh.ray[warp][0][owner];
                    const F3 O = f3(r[0], r[32], r[64]);
                    const F3 D1 = f3(r[96], r[128], r[160]), D2 = f3(r[192], r[224], r[256]);
                    if (CLOSEST) {
                        float tf = __uint_as_float(sh.tbest[warp][owner]);     // the owner's nearest hit so far
                        if (prim_hit<true>(sv, prim, O, D1, tf)) { atomicMin(&sh.tbest[warp][owner], __float_as_uint(tf)); q1 = true; }
                    } else {
                        prim_hit2<TWO>(sv, prim, O, D1, D2, tfar, q1, q2);
                    }
                    cnt.prims++;
                }
                if (q1) atomicOr(&sh.hit1[warp], 1u << owner);
                if (q2) atomicOr(&sh.hit2[warp], 1u << owner);
                base += nb;
                __syncwarp();
            }
            // what is left of this lane's list sits at its bottom
            const unsigned int end_o = excl + (unsigned int)L.pc;
            L.pc = (int)(end_o > base ? min(end_o - base, (unsigned int)L.pc) : 0u);
            pend_est = total - base;
            const unsigned int m1 = sh.hit1[warp], m2 = sh.hit2[warp];
            __syncwarp();
            if ((m1 | m2) != 0u && lane == 0) { sh.hit1[warp] = 0u; sh.hit2[warp] = 0u; }
            if ((m1 >> lane) & 1u) L.hit1 = true;
            if (((TWO ? m2 : m1) >> lane) & 1u) L.hit2 = true;
            if (CLOSEST) {
                const float tb = __uint_as_float(sh.tbest[warp][lane]);
                if (tb < L.tfar) L.tfar = tb;        // (a hit at exactly tfar is reported through hit1 like any other)
            } else
            if (L.state != 0) {
                if (L.hit1 && L.hit2) { L.state = 2; L.pc = 0; }          // both decided: drop the rest
                else if (!TWO) {}
                // one ray decided: it stops entering boxes (an infinite x offset makes its entry distance +inf),
                // the walk goes on for the other ray only
                else if (L.hit2) L.B2x = INFINITY;
                else if (L.hit1) L.B1x = INFINITY;
            }
        }
    }
    // ---- 4. retire
    if (L.state == 2 && L.pc == 0) L.state = 0;
    return __ballot_sync(FULL, L.state != 0);
}

}  // namespace hzb
