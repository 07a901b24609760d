// hzb_tri.cuh -- ray/triangle arithmetic of the product (host/device source, see hzb_hd.cuh).
//
// tri_hit() implements the project's intersection specification ("Ray/triangle test"):
// the Pluecker-coordinate edge test that Embree uses for RTC_SCENE_FLAG_ROBUST scenes (the reference
// sets that flag at horizon_comp.cpp:106), two-sided, eps = ulp*|U+V+W|, depth from the stable
// geometric normal, accepted for 0 <= t <= tfar (tnear = 0 at horizon_comp.cpp:252).  Every operation
// is an explicitly rounded intrinsic so that neither -fmad nor the optimiser can change a decision.
#pragma once
#include "hzb_hd.cuh"

namespace hzb {

struct F3 { float x, y, z; };

HZB_HD F3 f3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
HZB_HD F3 sub_rn(F3 a, F3 b) { return f3(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }
HZB_HD F3 add_rn(F3 a, F3 b) { return f3(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z)); }
// cross = (fma(ay,bz,-(az*by)), fma(az,bx,-(ax*bz)), fma(ax,by,-(ay*bx)))
HZB_HD F3 cross_f(F3 a, F3 b) {
    return f3(__fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)), __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)),
              __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x)));
}
// dot = fma(ax,bx, fma(ay,by, az*bz))
HZB_HD float dot_f(F3 a, F3 b) {
    return __fmaf_rn(a.x, b.x, __fmaf_rn(a.y, b.y, __fmul_rn(a.z, b.z)));
}

// Pluecker edge function of edge (a -> b) style term: dot(cross(e, s), D)
// with e, s prepared by the caller.

// Depth part of the test, shared by both entry points below.
HZB_HD bool tri_depth(F3 v0, F3 e0, F3 e1, F3 e2, F3 D, float tfar, float& t_out) {
    const float ab_x = __fmul_rn(e0.z, e1.y), ab_y = __fmul_rn(e0.x, e1.z), ab_z = __fmul_rn(e0.y, e1.x);
    const float bc_x = __fmul_rn(e1.z, e2.y), bc_y = __fmul_rn(e1.x, e2.z), bc_z = __fmul_rn(e1.y, e2.x);
    const float cab_x = __fmaf_rn(e0.y, e1.z, -ab_x), cab_y = __fmaf_rn(e0.z, e1.x, -ab_y), cab_z = __fmaf_rn(e0.x, e1.y, -ab_z);
    const float cbc_x = __fmaf_rn(e1.y, e2.z, -bc_x), cbc_y = __fmaf_rn(e1.z, e2.x, -bc_y), cbc_z = __fmaf_rn(e1.x, e2.y, -bc_z);
    const F3 Ng = f3(fabsf(ab_x) < fabsf(bc_x) ? cab_x : cbc_x, fabsf(ab_y) < fabsf(bc_y) ? cab_y : cbc_y,
                     fabsf(ab_z) < fabsf(bc_z) ? cab_z : cbc_z);
    const float dn = dot_f(Ng, D);
    const float den = __fadd_rn(dn, dn);
    if (den == 0.0f) return false;
    const float tn = dot_f(v0, Ng);
    const float t = __fdiv_rn(__fadd_rn(tn, tn), den);
    if (!(t >= 0.0f && t <= tfar)) return false;
    t_out = t;
    return true;
}

HZB_HD bool tri_hit(F3 p0, F3 p1, F3 p2, F3 O, F3 D, float tfar, float& t_out) {
    const F3 v0 = sub_rn(p0, O), v1 = sub_rn(p1, O), v2 = sub_rn(p2, O);
    const F3 e0 = sub_rn(v2, v0), e1 = sub_rn(v0, v1), e2 = sub_rn(v1, v2);
    const float U = dot_f(cross_f(e0, add_rn(v2, v0)), D);
    const float V = dot_f(cross_f(e1, add_rn(v0, v1)), D);
    const float W = dot_f(cross_f(e2, add_rn(v1, v2)), D);
    const float UVW = __fadd_rn(__fadd_rn(U, V), W);
    const float eps = __fmul_rn(FLT_EPSILON, fabsf(UVW));
    const float mn = fminf(fminf(U, V), W), mx = fmaxf(fmaxf(U, V), W);
    if (!((mn >= -eps) || (mx <= eps))) return false;
    return tri_depth(v0, e0, e1, e2, D, tfar, t_out);
}

// ---- two rays against one primitive -------------------------------------------------
// Edge test of tri_hit() from the three edge functions of one ray.
HZB_HD bool edges_accept(float U, float V, float W) {
    const float UVW = __fadd_rn(__fadd_rn(U, V), W);
    const float eps = __fmul_rn(FLT_EPSILON, fabsf(UVW));
    const float mn = fminf(fminf(U, V), W), mx = fmaxf(fmaxf(U, V), W);
    return (mn >= -eps) || (mx <= eps);
}
// The stable geometric normal of tri_depth() (ray independent).
HZB_HD F3 tri_ng(F3 e0, F3 e1, F3 e2) {
    const float ab_x = __fmul_rn(e0.z, e1.y), ab_y = __fmul_rn(e0.x, e1.z), ab_z = __fmul_rn(e0.y, e1.x);
    const float bc_x = __fmul_rn(e1.z, e2.y), bc_y = __fmul_rn(e1.x, e2.z), bc_z = __fmul_rn(e1.y, e2.x);
    const float cab_x = __fmaf_rn(e0.y, e1.z, -ab_x), cab_y = __fmaf_rn(e0.z, e1.x, -ab_y), cab_z = __fmaf_rn(e0.x, e1.y, -ab_z);
    const float cbc_x = __fmaf_rn(e1.y, e2.z, -bc_x), cbc_y = __fmaf_rn(e1.z, e2.x, -bc_y), cbc_z = __fmaf_rn(e1.x, e2.y, -bc_z);
    return f3(fabsf(ab_x) < fabsf(bc_x) ? cab_x : cbc_x, fabsf(ab_y) < fabsf(bc_y) ? cab_y : cbc_y,
              fabsf(ab_z) < fabsf(bc_z) ? cab_z : cbc_z);
}
HZB_HD bool depth_ok(F3 v0, F3 Ng, F3 D, float tfar) {
    const float dn = dot_f(Ng, D);
    const float den = __fadd_rn(dn, dn);
    if (den == 0.0f) return false;
    const float tn = dot_f(v0, Ng);
    const float t = __fdiv_rn(__fadd_rn(tn, tn), den);
    return t >= 0.0f && t <= tfar;
}

// Both triangles of a grid quad against two rays (TWO) or ray 1 only: the decisions of
// tri_hit(p00, p01, p10, ..) || tri_hit(p11, p10, p01, ..) per ray.  The triangles share their diagonal
// and the Pluecker edge function of the reversed edge is the exact negative, so five edge functions
// serve both triangles and their cross products serve both rays; the depth test (stable normal + one
// division) runs only for (triangle, ray) pairs whose edge test passed, in one shared loop.
template <bool TWO>
HZB_HD void quad_hit2(F3 p00, F3 p01, F3 p10, F3 p11, F3 O, F3 D1, F3 D2, float tfar, bool& h1, bool& h2) {
    // triangle 1 = (p00, p01, p10) = (a, b, c); triangle 2 = (p11, p10, p01) = (d, c, b)
    const F3 a = sub_rn(p00, O), b = sub_rn(p01, O), c = sub_rn(p10, O), d = sub_rn(p11, O);
    const F3 e0 = sub_rn(c, a), e1 = sub_rn(a, b), e2 = sub_rn(b, c);          // triangle 1
    const F3 f0 = sub_rn(b, d), f1 = sub_rn(d, c);                             // triangle 2 (its e2 = c - b = -e2)
    const F3 C0 = cross_f(e0, add_rn(c, a)), C1 = cross_f(e1, add_rn(a, b)), C2 = cross_f(e2, add_rn(b, c));
    const F3 G0 = cross_f(f0, add_rn(b, d)), G1 = cross_f(f1, add_rn(d, c));
    const float W1 = dot_f(C2, D1);
    const bool a11 = edges_accept(dot_f(C0, D1), dot_f(C1, D1), W1);
    const bool a21 = edges_accept(dot_f(G0, D1), dot_f(G1, D1), -W1);   // reversed diagonal: exact negative
    bool a12 = false, a22 = false;
    if (TWO) {
        const float W2 = dot_f(C2, D2);
        a12 = edges_accept(dot_f(C0, D2), dot_f(C1, D2), W2);
        a22 = edges_accept(dot_f(G0, D2), dot_f(G1, D2), -W2);
    }
    // depth tests (rare): one (triangle, ray) combination per round, shared code
    unsigned int acc = (a11 ? 1u : 0u) | (a12 ? 2u : 0u) | (a21 ? 4u : 0u) | (a22 ? 8u : 0u);
    while (acc) {
        const bool second = (acc & 3u) == 0u;                     // triangle 2 once triangle 1 is done
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

}  // namespace hzb
