// hzb_geom.cuh -- ray/triangle and ray/box device code.
//
// tri_hit() implements the project's intersection specification (DESIGN.md,
// "Ray/triangle test"): the Pluecker-coordinate edge test that Embree uses for
// RTC_SCENE_FLAG_ROBUST scenes (the reference sets that flag at
// horizon_comp.cpp:106), two-sided, eps = ulp*|U+V+W|, depth from the stable
// geometric normal, accepted for 0 <= t <= tfar (tnear = 0 at
// horizon_comp.cpp:252).  Every operation is an explicitly rounded intrinsic
// (__fsub_rn, __fmul_rn, __fmaf_rn, __fdiv_rn) so that neither -fmad nor the
// optimiser can change a hit/miss decision.
#pragma once
#include "hzb_common.cuh"
#include "hzb_tri.cuh"
#include <float.h>

namespace hzb {

__device__ __forceinline__ F3 ld_vert(const float4* p) {
    const float4 v = __ldg(p);
    return f3(v.x, v.y, v.z);
}

// Test primitive `prim` (grid quad or TIN triangle).  ANY: return on first hit.
// CLOSEST: shrink tfar to the smallest t.
template <bool CLOSEST>
__device__ __forceinline__ bool prim_hit(const SceneView& s, uint32_t prim, F3 O, F3 D, float& tfar) {
    float t;
    if (prim < s.num_quads) {
        const uint32_t wq = (uint32_t)(s.W - 1);
        const uint32_t i = prim / wq, j = prim - i * wq;
        const float4* r0 = s.vert4 + (size_t)i * s.W + j;
        const float4* r1 = r0 + s.W;
        const F3 p00 = ld_vert(r0), p01 = ld_vert(r0 + 1), p10 = ld_vert(r1), p11 = ld_vert(r1 + 1);
        bool hit = false;
        if (tri_hit(p00, p01, p10, O, D, tfar, t)) {
            if (!CLOSEST) return true;
            tfar = t; hit = true;
        }
        if (tri_hit(p11, p10, p01, O, D, tfar, t)) {
            if (!CLOSEST) return true;
            tfar = t; hit = true;
        }
        return hit;
    } else {
        const float4* q = s.tin4 + 3 * (size_t)(prim - s.num_quads);
        if (tri_hit(ld_vert(q), ld_vert(q + 1), ld_vert(q + 2), O, D, tfar, t)) { tfar = t; return true; }
        return false;
    }
}

// ------------------------------------------------------------------- boxes
struct RayInv { float ix, iy, iz; };
__device__ __forceinline__ float safe_rcp(float d) {
    if (fabsf(d) < 1e-30f) d = copysignf(1e-30f, d);
    return 1.0f / d;
}
__device__ __forceinline__ RayInv make_inv(F3 D) { RayInv r; r.ix = safe_rcp(D.x); r.iy = safe_rcp(D.y); r.iz = safe_rcp(D.z); return r; }

// Conservative slab test over [0, tfar]: boxes are padded at build time and
// the exit distance gets a relative slack, so a ray accepted by tri_hit can
// never be culled by an ancestor box.
__device__ __forceinline__ bool slab(const float* lo, const float* hi, F3 O, RayInv inv, float tfar, float& tnear) {
    float ta = (lo[0] - O.x) * inv.ix, tb = (hi[0] - O.x) * inv.ix;
    float t0 = fminf(ta, tb), t1 = fmaxf(ta, tb);
    ta = (lo[1] - O.y) * inv.iy; tb = (hi[1] - O.y) * inv.iy;
    t0 = fmaxf(t0, fminf(ta, tb)); t1 = fminf(t1, fmaxf(ta, tb));
    ta = (lo[2] - O.z) * inv.iz; tb = (hi[2] - O.z) * inv.iz;
    t0 = fmaxf(t0, fminf(ta, tb)); t1 = fminf(t1, fmaxf(ta, tb));
    t0 = fmaxf(t0, 0.0f); t1 = fminf(t1, tfar);
    tnear = t0;
    return t0 <= t1 * 1.000001f;
}

struct LaneCounters { unsigned int rays, nodes, prims; };

// --------------------------------------------------- BVH2 per-thread trace
constexpr int HZB_STACK2 = 96;

template <bool CLOSEST>
__device__ bool trace_bvh2(const SceneView& s, F3 O, F3 D, float& tfar, LaneCounters& cnt, unsigned int* overflow) {
    const RayInv inv = make_inv(D);
    int stack[HZB_STACK2];
    int sp = 0;
    int node = 0;
    bool any = false;
    while (true) {
        const float4* np = reinterpret_cast<const float4*>(s.nodes2 + node);
        const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
        cnt.nodes++;
        const float lo0[3] = {n0.x, n0.y, n0.z}, hi0[3] = {n0.w, n1.x, n1.y};
        const float lo1[3] = {n1.z, n1.w, n2.x}, hi1[3] = {n2.y, n2.z, n2.w};
        const int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
        float tn0, tn1;
        bool h0 = slab(lo0, hi0, O, inv, tfar, tn0);
        bool h1 = slab(lo1, hi1, O, inv, tfar, tn1);
        if (h0 && c0 < 0) {
            cnt.prims++;
            if (prim_hit<CLOSEST>(s, (uint32_t)(~c0), O, D, tfar)) { if (!CLOSEST) return true; any = true; }
            h0 = false;
        }
        if (h1 && c1 < 0) {
            cnt.prims++;
            if (prim_hit<CLOSEST>(s, (uint32_t)(~c1), O, D, tfar)) { if (!CLOSEST) return true; any = true; }
            h1 = false;
        }
        if (h0 && h1) {
            int nearc = c0, farc = c1;
            if (tn1 < tn0) { nearc = c1; farc = c0; }
            if (sp < HZB_STACK2) stack[sp++] = farc; else atomicAdd(overflow, 1u);
            node = nearc;
        } else if (h0) node = c0;
        else if (h1) node = c1;
        else {
            if (sp == 0) break;
            node = stack[--sp];
        }
    }
    return any;
}

}  // namespace hzb
