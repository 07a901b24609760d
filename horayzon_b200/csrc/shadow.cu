// shadow.cu -- shadow mask / shortwave correction kernels (B200, sm_100a).
//
// Replaces CppTerrain::shadow and ::sw_dir_cor (shadow_comp.cpp:386-491,
// 495-605): one lane per inner-domain cell, one any-hit ray with tfar = +inf
// towards the (optionally refracted) sun.  The sun-vector arithmetic keeps the
// reference's float operation order with explicitly rounded intrinsics; the
// refraction branch evaluates the libm calls in double and rounds once (the
// closest available stand-in for glibc's float functions).
#include "hzb_geom.cuh"
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace hzb {
namespace {

__device__ __forceinline__ float deg2rad_d(float a) { return __double2float_rn(((double)a / 180.0) * M_PI); }
__device__ __forceinline__ float rad2deg_d(float a) { return __double2float_rn(((double)a / M_PI) * 180.0); }

__device__ __forceinline__ void unit3(F3& v) {  // shadow_comp.cpp:96-106
    const float mag = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)), __fmul_rn(v.z, v.z)));
    v = f3(__fdiv_rn(v.x, mag), __fdiv_rn(v.y, mag), __fdiv_rn(v.z, mag));
}
__device__ __forceinline__ float dot3_rn(F3 a, F3 b) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}

// Saemundsson refraction (shadow_comp.cpp:135-159), degrees in / out
__device__ __forceinline__ float refraction_deg(float elev_true, float temp_c, float pressure) {
    elev_true = fmaxf(-1.0f, fminf(elev_true, 90.0f));
    const float arg = deg2rad_d(__double2float_rn((double)elev_true + 10.3 / ((double)elev_true + 5.11)));
    float r = __double2float_rn(1.02 / (double)__double2float_rn(tan((double)arg)));
    r = __double2float_rn((double)r + 0.0019279);
    r = __double2float_rn((double)r * (((double)pressure / 101.0) * (283.0 / (273.0 + (double)temp_c))));
    return __double2float_rn((double)r * (1.0 / 60.0));
}

template <bool SW>
__global__ void __launch_bounds__(128) k_terrain(SceneView sv, TerrainParams tp, float sunx, float suny, float sunz,
                                                 uint8_t* __restrict__ shadow, float* __restrict__ swc, Counters* counters) {
    const long long ncell = (long long)tp.dim_in_0 * tp.dim_in_1;
    const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    LaneCounters cnt; cnt.rays = cnt.nodes = cnt.prims = 0;
    unsigned int units = 0;
    if (c < ncell) {
        if (tp.mask[c] != 1) {
            if (SW) swc[c] = tp.sw_dir_cor_fill; else shadow[c] = 3;
        } else {
            units = 1;
            const int i = (int)(c / tp.dim_in_1), j = (int)(c - (long long)i * tp.dim_in_1);
            const F3 tilt = f3(tp.vec_tilt[3 * c], tp.vec_tilt[3 * c + 1], tp.vec_tilt[3 * c + 2]);
            const F3 nrm = f3(tp.vec_norm[3 * c], tp.vec_norm[3 * c + 1], tp.vec_norm[3 * c + 2]);
            const float4 v = sv.vert4[(size_t)(i + tp.offset_0) * sv.W + (j + tp.offset_1)];
            const float lift = 0.05f;  // shadow_comp.cpp:388, 497
            const F3 org = f3(__fadd_rn(v.x, __fmul_rn(nrm.x, lift)), __fadd_rn(v.y, __fmul_rn(nrm.y, lift)),
                              __fadd_rn(v.z, __fmul_rn(nrm.z, lift)));
            F3 sun = f3(__fsub_rn(sunx, org.x), __fsub_rn(suny, org.y), __fsub_rn(sunz, org.z));
            unit3(sun);
            float dns = dot3_rn(nrm, sun);
            if (tp.refrac_cor == 1) {  // shadow_comp.cpp:430-446
                const float elev_true = __double2float_rn(90.0 - (double)rad2deg_d(__double2float_rn(acos((double)dns))));
                const float temperature = __fsub_rn(tp.t_ref, __fmul_rn(tp.lapse, tp.elevation[c]));
                const float pressure = __fmul_rn(tp.p_ref, __double2float_rn(pow((double)__fdiv_rn(temperature, tp.t_ref), (double)tp.expo)));
                const float rc = refraction_deg(elev_true, __double2float_rn((double)temperature - 273.15), pressure);
                F3 k = f3(__fsub_rn(__fmul_rn(sun.y, nrm.z), __fmul_rn(sun.z, nrm.y)),
                          __fsub_rn(__fmul_rn(sun.z, nrm.x), __fmul_rn(sun.x, nrm.z)),
                          __fsub_rn(__fmul_rn(sun.x, nrm.y), __fmul_rn(sun.y, nrm.x)));
                unit3(k);
                const float th = deg2rad_d(rc);
                const float ct = __double2float_rn(cos((double)th)), st = __double2float_rn(sin((double)th));
                const float part = __double2float_rn((double)dot3_rn(k, sun) * (1.0 - (double)ct));
                const F3 r = f3(
                    __fadd_rn(__fadd_rn(__fmul_rn(sun.x, ct), __fmul_rn(__fsub_rn(__fmul_rn(k.y, sun.z), __fmul_rn(k.z, sun.y)), st)), __fmul_rn(k.x, part)),
                    __fadd_rn(__fadd_rn(__fmul_rn(sun.y, ct), __fmul_rn(__fsub_rn(__fmul_rn(k.z, sun.x), __fmul_rn(k.x, sun.z)), st)), __fmul_rn(k.y, part)),
                    __fadd_rn(__fadd_rn(__fmul_rn(sun.z, ct), __fmul_rn(__fsub_rn(__fmul_rn(k.x, sun.y), __fmul_rn(k.y, sun.x)), st)), __fmul_rn(k.z, part)));
                sun = r;
                dns = dot3_rn(nrm, sun);
            }
            const float dts = dot3_rn(tilt, sun);
            const float dot_min = SW ? tp.dot_prod_min : 0.0f;
            if (dts > dot_min) {
                cnt.rays++;
                float tfar = INFINITY;
                const bool occ = trace_bvh2<false>(sv, org, sun, tfar, cnt, reinterpret_cast<unsigned int*>(&counters->stack_overflow));
                if (SW) {
                    if (occ) swc[c] = 0.0f;
                    else {
                        if (dns < dot_min) dns = dot_min;
                        swc[c] = __fmul_rn(__fdiv_rn(dts, dns), tp.surf_enl_fac[c]);  // :581-585
                    }
                } else shadow[c] = occ ? 2 : 0;
            } else {
                if (SW) swc[c] = 0.0f; else shadow[c] = 1;
            }
        }
    }
    unsigned int r = cnt.rays, n = cnt.nodes, p = cnt.prims, u = units;
    for (int o = 16; o > 0; o >>= 1) {
        r += __shfl_xor_sync(0xffffffffu, r, o); n += __shfl_xor_sync(0xffffffffu, n, o);
        p += __shfl_xor_sync(0xffffffffu, p, o); u += __shfl_xor_sync(0xffffffffu, u, o);
    }
    if ((threadIdx.x & 31) == 0 && (r | n | p | u)) {
        atomicAdd(&counters->rays, (unsigned long long)r); atomicAdd(&counters->node_visits, (unsigned long long)n);
        atomicAdd(&counters->prim_tests, (unsigned long long)p); atomicAdd(&counters->units, (unsigned long long)u);
    }
}


// ---------------------------------------------------------------------------
// Production kernel: persistent warps, one lane per cell, warp-queue traversal
// of the compressed 4-wide BVH (same scheme as k_horizon_wq4 in horizon.cu):
// lanes refill with the next cell of the warp's block as soon as their ray
// retires; leaf candidates of all lanes are tested 32 at a time.
// ---------------------------------------------------------------------------
constexpr int TW_THREADS = 128;
constexpr int TW_WARPS = TW_THREADS / 32;
constexpr int TW_STACK = 40;
constexpr int TW_RING = 256;
constexpr int TW_BLOCK = 512;   // cells per warp work block
constexpr uint32_t TW_NONE = 0xFFFFFFFFu;

struct TwShared {
    uint32_t stack[TW_STACK][TW_THREADS];
    uint2 ring[TW_WARPS][TW_RING];
    float ray[TW_WARPS][6][32];
    unsigned int hitmask[TW_WARPS];
};

// Sun vector, self-shading test and ray set-up for one cell (shadow_comp.cpp:397-451 / 507-561).
// Returns true if an occlusion ray must be cast.
template <bool SW>
__device__ __forceinline__ bool terrain_cell_setup(const SceneView& sv, const TerrainParams& tp, long long c, float sunx,
                                                   float suny, float sunz, F3& org, F3& sun, float& dts, float& dns) {
    const int i = (int)(c / tp.dim_in_1), j = (int)(c - (long long)i * tp.dim_in_1);
    const F3 tilt = f3(tp.vec_tilt[3 * c], tp.vec_tilt[3 * c + 1], tp.vec_tilt[3 * c + 2]);
    const F3 nrm = f3(tp.vec_norm[3 * c], tp.vec_norm[3 * c + 1], tp.vec_norm[3 * c + 2]);
    const float4 v = sv.vert4[(size_t)(i + tp.offset_0) * sv.W + (j + tp.offset_1)];
    const float lift = 0.05f;  // shadow_comp.cpp:388, 497
    org = f3(__fadd_rn(v.x, __fmul_rn(nrm.x, lift)), __fadd_rn(v.y, __fmul_rn(nrm.y, lift)), __fadd_rn(v.z, __fmul_rn(nrm.z, lift)));
    sun = f3(__fsub_rn(sunx, org.x), __fsub_rn(suny, org.y), __fsub_rn(sunz, org.z));
    unit3(sun);
    dns = dot3_rn(nrm, sun);
    if (tp.refrac_cor == 1) {  // shadow_comp.cpp:430-446
        const float elev_true = __double2float_rn(90.0 - (double)rad2deg_d(__double2float_rn(acos((double)dns))));
        const float temperature = __fsub_rn(tp.t_ref, __fmul_rn(tp.lapse, tp.elevation[c]));
        const float pressure = __fmul_rn(tp.p_ref, __double2float_rn(pow((double)__fdiv_rn(temperature, tp.t_ref), (double)tp.expo)));
        const float rc = refraction_deg(elev_true, __double2float_rn((double)temperature - 273.15), pressure);
        F3 k = f3(__fsub_rn(__fmul_rn(sun.y, nrm.z), __fmul_rn(sun.z, nrm.y)), __fsub_rn(__fmul_rn(sun.z, nrm.x), __fmul_rn(sun.x, nrm.z)),
                  __fsub_rn(__fmul_rn(sun.x, nrm.y), __fmul_rn(sun.y, nrm.x)));
        unit3(k);
        const float th = deg2rad_d(rc);
        const float ct = __double2float_rn(cos((double)th)), st = __double2float_rn(sin((double)th));
        const float part = __double2float_rn((double)dot3_rn(k, sun) * (1.0 - (double)ct));
        const F3 r = f3(
            __fadd_rn(__fadd_rn(__fmul_rn(sun.x, ct), __fmul_rn(__fsub_rn(__fmul_rn(k.y, sun.z), __fmul_rn(k.z, sun.y)), st)), __fmul_rn(k.x, part)),
            __fadd_rn(__fadd_rn(__fmul_rn(sun.y, ct), __fmul_rn(__fsub_rn(__fmul_rn(k.z, sun.x), __fmul_rn(k.x, sun.z)), st)), __fmul_rn(k.y, part)),
            __fadd_rn(__fadd_rn(__fmul_rn(sun.z, ct), __fmul_rn(__fsub_rn(__fmul_rn(k.x, sun.y), __fmul_rn(k.y, sun.x)), st)), __fmul_rn(k.z, part)));
        sun = r;
        dns = dot3_rn(nrm, sun);
    }
    dts = dot3_rn(tilt, sun);
    const float dot_min = SW ? tp.dot_prod_min : 0.0f;
    return dts > dot_min;
}

template <bool SW>
__global__ void __launch_bounds__(TW_THREADS, 6) k_terrain_wq4(SceneView sv, TerrainParams tp, float sunx, float suny, float sunz,
                                                               uint8_t* __restrict__ shadow, float* __restrict__ swc,
                                                               Counters* counters, unsigned int* block_counter, int refill_thr,
                                                               int wait_thr) {
    __shared__ TwShared sh;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
    const unsigned int FULL = 0xffffffffu, lt_mask = (1u << lane) - 1u;
    const long long ncell = (long long)tp.dim_in_0 * tp.dim_in_1;
    const unsigned int num_blocks = (unsigned int)((ncell + TW_BLOCK - 1) / TW_BLOCK);
    const float dot_min = SW ? tp.dot_prod_min : 0.0f;
    const float tfar = INFINITY;
    unsigned int* overflow = reinterpret_cast<unsigned int*>(&counters->stack_overflow);
    LaneCounters cnt; cnt.rays = cnt.nodes = cnt.prims = 0;
    unsigned int units = 0;
    uint2* ring = sh.ring[warp];
    if (lane == 0) sh.hitmask[warp] = 0u;
    unsigned int pushed = 0, tested = 0;
    __syncwarp();

    long long next = 0, end = 0;          // warp-uniform: unassigned cells of the current block
    bool more_blocks = true;
    bool ray_active = false, ray_hit = false, have_queued = false;
    long long cell = -1; float dts = 0.f, dns = 0.f;
    float Ax = 0.f, Ay = 0.f, Az = 0.f, Bx = 0.f, By = 0.f, Bz = 0.f;
    unsigned int selnx = 0x7410u, selny = 0x7410u, selnz = 0x7410u;
    uint32_t node = TW_NONE; int sp = 0; unsigned int my_last = 0;

    while (true) {
        // (1) retire finished rays and hand out new cells
        if (cell >= 0 && !ray_active) {   // ray finished: write the result (shadow_comp.cpp:468-472 / 578-586)
            if (SW) {
                if (ray_hit) swc[cell] = 0.0f;
                else { if (dns < dot_min) dns = dot_min; swc[cell] = __fmul_rn(__fdiv_rn(dts, dns), tp.surf_enl_fac[cell]); }
            } else shadow[cell] = ray_hit ? 2 : 0;
            cell = -1;
        }
        while (true) {
            const bool want = !ray_active;
            const unsigned int wmask = __ballot_sync(FULL, want);
            if (wmask == 0u) break;
            if (next >= end) {
                if (!more_blocks) break;
                unsigned int b = 0;
                if (lane == 0) b = atomicAdd(block_counter, 1u);
                b = __shfl_sync(FULL, b, 0);
                if (b >= num_blocks) { more_blocks = false; break; }
                next = (long long)b * TW_BLOCK; end = min(next + (long long)TW_BLOCK, ncell);
            }
            const long long mine = next + __popc(wmask & lt_mask);
            next += __popc(wmask);
            if (want && mine < end) {
                if (tp.mask[mine] != 1) {
                    if (SW) swc[mine] = tp.sw_dir_cor_fill; else shadow[mine] = 3;
                } else {
                    units++;
                    F3 org, sun;
                    if (terrain_cell_setup<SW>(sv, tp, mine, sunx, suny, sunz, org, sun, dts, dns)) {
                        const RayInv inv = make_inv(sun);
                        Ax = sv.qstep[0] * inv.ix; Ay = sv.qstep[1] * inv.iy; Az = sv.qstep[2] * inv.iz;
                        Bx = (sv.qorg[0] - org.x) * inv.ix; By = (sv.qorg[1] - org.y) * inv.iy; Bz = (sv.qorg[2] - org.z) * inv.iz;
                        selnx = inv.ix >= 0.f ? 0x7410u : 0x7432u;
                        selny = inv.iy >= 0.f ? 0x7410u : 0x7432u;
                        selnz = inv.iz >= 0.f ? 0x7410u : 0x7432u;
                        sh.ray[warp][0][lane] = org.x; sh.ray[warp][1][lane] = org.y; sh.ray[warp][2][lane] = org.z;
                        sh.ray[warp][3][lane] = sun.x; sh.ray[warp][4][lane] = sun.y; sh.ray[warp][5][lane] = sun.z;
                        cell = mine; node = 0u; sp = 0; ray_active = true; ray_hit = false; have_queued = false; cnt.rays++;
                    } else {
                        if (SW) swc[mine] = 0.0f; else shadow[mine] = 1;   // self-shaded (:474-478 / :588-592)
                    }
                }
            }
        }
        const unsigned int act_mask = __ballot_sync(FULL, ray_active);
        if (act_mask == 0u) break;
        const int thr = (more_blocks || next < end) ? refill_thr : 1;
        __syncwarp();

        // (2) traversal (see k_horizon_wq4)
        while (true) {
            const bool do_node = ray_active && !ray_hit && node != TW_NONE;
            bool lf0 = false, lf1 = false, lf2 = false, lf3 = false;
            uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
            if (do_node) {
                const uint4* np = reinterpret_cast<const uint4*>(sv.nodes4 + node);
                const uint4 r0 = __ldg(np), r1 = __ldg(np + 1), r2 = __ldg(np + 2), r3 = __ldg(np + 3);
                cnt.nodes++;
                float t0, t1, t2, t3; bool h0, h1, h2, h3;
                wide_child_test(r0, selnx, selny, selnz, Ax, Ay, Az, Bx, By, Bz, tfar, t0, h0);
                wide_child_test(r1, selnx, selny, selnz, Ax, Ay, Az, Bx, By, Bz, tfar, t1, h1);
                wide_child_test(r2, selnx, selny, selnz, Ax, Ay, Az, Bx, By, Bz, tfar, t2, h2);
                wide_child_test(r3, selnx, selny, selnz, Ax, Ay, Az, Bx, By, Bz, tfar, t3, h3);
                w0 = r0.w; w1 = r1.w; w2 = r2.w; w3 = r3.w;
                lf0 = h0 && (w0 & WIDE_LEAF); lf1 = h1 && (w1 & WIDE_LEAF); lf2 = h2 && (w2 & WIDE_LEAF); lf3 = h3 && (w3 & WIDE_LEAF);
                const bool i0 = h0 && !(w0 & WIDE_LEAF), i1 = h1 && !(w1 & WIDE_LEAF), i2 = h2 && !(w2 & WIDE_LEAF), i3 = h3 && !(w3 & WIDE_LEAF);
                const float k0 = i0 ? t0 : INFINITY, k1 = i1 ? t1 : INFINITY, k2 = i2 ? t2 : INFINITY, k3 = i3 ? t3 : INFINITY;
                const float kmin = fminf(fminf(k0, k1), fminf(k2, k3));
                const bool any_int = i0 || i1 || i2 || i3;
                const int idx = (i0 && k0 == kmin) ? 0 : ((i1 && k1 == kmin) ? 1 : ((i2 && k2 == kmin) ? 2 : 3));
                const uint32_t nearest = idx == 0 ? w0 : (idx == 1 ? w1 : (idx == 2 ? w2 : w3));
                if (sp + 3 > TW_STACK) { if (any_int) atomicAdd(overflow, 1u); }
                else {
                    if (i0 && idx != 0) { sh.stack[sp][tid] = w0; ++sp; }
                    if (i1 && idx != 1) { sh.stack[sp][tid] = w1; ++sp; }
                    if (i2 && idx != 2) { sh.stack[sp][tid] = w2; ++sp; }
                    if (i3 && idx != 3) { sh.stack[sp][tid] = w3; ++sp; }
                }
                if (any_int) node = nearest;
                else if (sp > 0) { --sp; node = sh.stack[sp][tid]; }
                else node = TW_NONE;
            }
            {
                const unsigned int m0 = __ballot_sync(FULL, lf0), m1 = __ballot_sync(FULL, lf1);
                const unsigned int m2 = __ballot_sync(FULL, lf2), m3 = __ballot_sync(FULL, lf3);
                unsigned int base = pushed;
                if (lf0) { const unsigned int q = base + __popc(m0 & lt_mask); ring[q & (TW_RING - 1)] = make_uint2(w0 & 0x7FFFFFFFu, lane); my_last = q; have_queued = true; }
                base += __popc(m0);
                if (lf1) { const unsigned int q = base + __popc(m1 & lt_mask); ring[q & (TW_RING - 1)] = make_uint2(w1 & 0x7FFFFFFFu, lane); my_last = q; have_queued = true; }
                base += __popc(m1);
                if (lf2) { const unsigned int q = base + __popc(m2 & lt_mask); ring[q & (TW_RING - 1)] = make_uint2(w2 & 0x7FFFFFFFu, lane); my_last = q; have_queued = true; }
                base += __popc(m2);
                if (lf3) { const unsigned int q = base + __popc(m3 & lt_mask); ring[q & (TW_RING - 1)] = make_uint2(w3 & 0x7FFFFFFFu, lane); my_last = q; have_queued = true; }
                pushed = base + __popc(m3);
            }
            {
                const bool drained = !have_queued || (int)(tested - my_last) > 0;
                const bool waiting = ray_active && (ray_hit || node == TW_NONE) && !drained;
                const unsigned int wmask = __ballot_sync(FULL, waiting);
                const unsigned int trav = __ballot_sync(FULL, ray_active && !ray_hit && node != TW_NONE);
                unsigned int avail = pushed - tested;
                bool flush = avail > 0u && (__popc(wmask) >= wait_thr || trav == 0u);
                __syncwarp();
                while (avail >= 32u || flush) {
                    const unsigned int nb = min(avail, 32u);
                    bool hit = false; unsigned int owner = 0;
                    if ((unsigned int)lane < nb) {
                        const uint2 e = ring[(tested + lane) & (TW_RING - 1)];
                        owner = e.y;
                        const F3 O = f3(sh.ray[warp][0][owner], sh.ray[warp][1][owner], sh.ray[warp][2][owner]);
                        const F3 D = f3(sh.ray[warp][3][owner], sh.ray[warp][4][owner], sh.ray[warp][5][owner]);
                        float tf = tfar;
                        hit = prim_hit<false>(sv, e.x, O, D, tf);
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
                if (ray_active && (ray_hit || node == TW_NONE) && drained) ray_active = false;
            }
            if (__popc(__ballot_sync(FULL, ray_active)) < thr) break;
        }
    }
    // final results of lanes that retired in the last traversal round were written at loop top? No: write them here.
    if (cell >= 0 && !ray_active) {
        if (SW) {
            if (ray_hit) swc[cell] = 0.0f;
            else { if (dns < dot_min) dns = dot_min; swc[cell] = __fmul_rn(__fdiv_rn(dts, dns), tp.surf_enl_fac[cell]); }
        } else shadow[cell] = ray_hit ? 2 : 0;
    }
    unsigned int r = cnt.rays, n = cnt.nodes, pp = cnt.prims, u = units;
    for (int o = 16; o > 0; o >>= 1) {
        r += __shfl_xor_sync(FULL, r, o); n += __shfl_xor_sync(FULL, n, o);
        pp += __shfl_xor_sync(FULL, pp, o); u += __shfl_xor_sync(FULL, u, o);
    }
    if (lane == 0 && (r | n | pp | u)) {
        atomicAdd(&counters->rays, (unsigned long long)r); atomicAdd(&counters->node_visits, (unsigned long long)n);
        atomicAdd(&counters->prim_tests, (unsigned long long)pp); atomicAdd(&counters->units, (unsigned long long)u);
    }
}

}  // namespace

int launch_shadow(Scene& s, const TerrainParams& tp, const float* sun, uint8_t* d_out, cudaStream_t st) {
    const long long ncell = (long long)tp.dim_in_0 * tp.dim_in_1;
    if (ncell <= 0) return 0;
    static const bool simple = getenv("HZB_SHADOW_KERNEL") && !strcmp(getenv("HZB_SHADOW_KERNEL"), "simple");
    if (simple) k_terrain<false><<<(unsigned int)((ncell + 127) / 128), 128, 0, st>>>(s.view(), tp, sun[0], sun[1], sun[2], d_out, nullptr, s.d_counters);
    else {
        HZB_CUDA(cudaMemsetAsync(s.d_tile_counter, 0, sizeof(unsigned int), st));
        k_terrain_wq4<false><<<sm_count() * 6, TW_THREADS, 0, st>>>(s.view(), tp, sun[0], sun[1], sun[2], d_out, nullptr, s.d_counters, s.d_tile_counter, 24, 6);
    }
    HZB_CUDA(cudaGetLastError());
    return 0;
}
int launch_sw_dir_cor(Scene& s, const TerrainParams& tp, const float* sun, float* d_out, cudaStream_t st) {
    const long long ncell = (long long)tp.dim_in_0 * tp.dim_in_1;
    if (ncell <= 0) return 0;
    static const bool simple = getenv("HZB_SHADOW_KERNEL") && !strcmp(getenv("HZB_SHADOW_KERNEL"), "simple");
    if (simple) k_terrain<true><<<(unsigned int)((ncell + 127) / 128), 128, 0, st>>>(s.view(), tp, sun[0], sun[1], sun[2], nullptr, d_out, s.d_counters);
    else {
        HZB_CUDA(cudaMemsetAsync(s.d_tile_counter, 0, sizeof(unsigned int), st));
        k_terrain_wq4<true><<<sm_count() * 6, TW_THREADS, 0, st>>>(s.view(), tp, sun[0], sun[1], sun[2], nullptr, d_out, s.d_counters, s.d_tile_counter, 24, 6);
    }
    HZB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace hzb
