// hzb_box.cuh -- ray constants and the quantised box test of the packet traversal step
// (host/device source, see hzb_hd.cuh; the CPU suite builds it for the host and checks that the test
// never rejects a box the ray meets in exact arithmetic).
#pragma once
#include "hzb_hd.cuh"
#include "hzb_tri.cuh"
#include <stdint.h>

namespace hzb {

struct Wq2Lane {
    int state;            // 0 no ray, 1 traversing, 2 traversal over, waiting for its pending candidates
    bool hit1, hit2;
    uint32_t node; int sp, pc;
    float A1x, A1y, A1z, B1x, B1y, B1z;      // ray 1 (upper): t = qb * A + B, bias folded into B
    float A2x, A2y, A2z, B2x, B2y, B2z;      // ray 2 (lower)
    unsigned int selxy;                       // PRMT selectors of the near x (low half) / y (high half) planes, shared by both rays
    float tfar;                               // closest-hit mode only: distance of the nearest hit so far (box tests cull beyond it)
};

// Reciprocal for the box tests only (never for a hit decision): the clamp keeps
// 2^23 * A finite, the approximate reciprocal (2 ulp) is far inside the box padding.
HZB_HD float safe_rcp2(float d) {
    if (fabsf(d) < 1e-18f) d = copysignf(1e-18f, d);
    return __fdividef(1.0f, d);
}

HZB_HD void wq2_ray_consts(const float* qorg, const float* qstep, F3 O, float ix, float iy, float iz,
                                               float& Ax, float& Ay, float& Az, float& Bx, float& By, float& Bz) {
    const float M = 8388608.0f;
    Ax = qstep[0] * ix; Ay = qstep[1] * iy; Az = qstep[2] * iz;
    Bx = fmaf(-M, Ax, (qorg[0] - O.x) * ix);
    By = fmaf(-M, Ay, (qorg[1] - O.y) * iy);
    Bz = fmaf(-M, Az, (qorg[2] - O.z) * iz);
}

// Ray constants and plane selectors of a packet (the arithmetic part of wq2_start).  Returns false
// when the two rays cannot share the plane selectors (different x or y direction signs): ray 2 then
// becomes a copy of ray 1 (D2 is overwritten).
HZB_HD bool wq2_set_rays(Wq2Lane& L, const float* qorg, const float* qstep, F3 O, F3 D1, F3& D2) {
    const float i1x = safe_rcp2(D1.x), i1y = safe_rcp2(D1.y), i1z = safe_rcp2(D1.z);
    float i2x = safe_rcp2(D2.x), i2y = safe_rcp2(D2.y), i2z = safe_rcp2(D2.z);
    const bool same = ((i1x >= 0.f) == (i2x >= 0.f)) && ((i1y >= 0.f) == (i2y >= 0.f));
    if (!same) { D2 = D1; i2x = i1x; i2y = i1y; i2z = i1z; }
    wq2_ray_consts(qorg, qstep, O, i1x, i1y, i1z, L.A1x, L.A1y, L.A1z, L.B1x, L.B1y, L.B1z);
    wq2_ray_consts(qorg, qstep, O, i2x, i2y, i2z, L.A2x, L.A2y, L.A2z, L.B2x, L.B2y, L.B2z);
    L.selxy = (i1x >= 0.f ? 0x7410u : 0x7432u) | (i1y >= 0.f ? 0x74100000u : 0x74320000u);
    return same;
}

// Box test of one child for both rays (TWO) or for ray 1 only.  key = entry distance.
template <bool TWO>
HZB_HD bool wide2_child_test(const uint4 r, const Wq2Lane& L, float tfar, float& key) {
    const unsigned int selx = L.selxy, sely = L.selxy >> 16;       // PRMT reads the low 16 selector bits only
    const float qnx = __uint_as_float(__byte_perm(r.x, 0x4B000000u, selx));
    const float qfx = __uint_as_float(__byte_perm(r.x, 0x4B000000u, selx ^ 0x0022u));
    const float qny = __uint_as_float(__byte_perm(r.y, 0x4B000000u, sely));
    const float qfy = __uint_as_float(__byte_perm(r.y, 0x4B000000u, sely ^ 0x0022u));
    const float qzl = __uint_as_float(__byte_perm(r.z, 0x4B000000u, 0x7410u));
    const float qzh = __uint_as_float(__byte_perm(r.z, 0x4B000000u, 0x7432u));
    float za = fmaf(qzl, L.A1z, L.B1z), zb = fmaf(qzh, L.A1z, L.B1z);
    const float tmin1 = fmaxf(fmaxf(fmaf(qnx, L.A1x, L.B1x), fmaf(qny, L.A1y, L.B1y)), fmaxf(fminf(za, zb), 0.0f));
    const float tmax1 = fminf(fminf(fmaf(qfx, L.A1x, L.B1x), fmaf(qfy, L.A1y, L.B1y)), fminf(fmaxf(za, zb), tfar));
    // no relative slack on tmax: the extra quantum on every plane (bvh_wide.cu) leaves half a
    // quantum (7.6e-6 of the scene extent) beyond the fold error, the rounding of t is ~1e-7 of it
    bool hit = tmin1 <= tmax1;
    key = tmin1;
    if (TWO) {
        za = fmaf(qzl, L.A2z, L.B2z); zb = fmaf(qzh, L.A2z, L.B2z);
        const float tmin2 = fmaxf(fmaxf(fmaf(qnx, L.A2x, L.B2x), fmaf(qny, L.A2y, L.B2y)), fmaxf(fminf(za, zb), 0.0f));
        const float tmax2 = fminf(fminf(fmaf(qfx, L.A2x, L.B2x), fmaf(qfy, L.A2y, L.B2y)), fminf(fmaxf(za, zb), tfar));
        hit = hit || (tmin2 <= tmax2);
        key = fminf(tmin1, tmin2);
    }
    return hit && (r.w != WIDE_EMPTY);
}

}  // namespace hzb
