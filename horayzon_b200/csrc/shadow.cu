// shadow.cu -- shadow mask / shortwave correction kernels (B200, sm_100a).
//
// Replaces CppTerrain::shadow and ::sw_dir_cor (shadow_comp.cpp:386-491,
// 495-605): one lane per inner-domain cell, one any-hit ray with tfar = +inf
// towards the (optionally refracted) sun.  The sun-vector arithmetic keeps the
// reference's float operation order with explicitly rounded intrinsics; the
// refraction branch evaluates the libm calls in double and rounds once (the
// closest available stand-in for glibc's float functions).
#include "hzb_geom.cuh"
#include "hzb_wq2.cuh"
#include <math.h>
#include <algorithm>

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

// REDO: only the cells the production kernel marked (full traversal stack) are computed.
template <bool SW, bool REDO>
__global__ void __launch_bounds__(128) k_terrain(SceneView sv, TerrainParams tp, float sunx, float suny, float sunz,
                                                 uint8_t* shadow, float* swc, Counters* counters) {
    const long long ncell = (long long)tp.dim_in_0 * tp.dim_in_1;
    const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    LaneCounters cnt; cnt.rays = cnt.nodes = cnt.prims = 0;
    unsigned int units = 0;
    const bool todo = c < ncell && (!REDO || (SW ? __float_as_uint(swc[c]) == HZB_REDO_F32 : shadow[c] == HZB_REDO_U8));
    if (todo) {
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
// Production kernel: persistent warps, one lane per cell, the shared warp-queue
// traversal of the compressed 4-wide BVH (hzb_wq2.cuh): lanes refill with the
// next cell of the warp's block as soon as their ray retires; leaf candidates
// of all lanes are tested 32 at a time.  (k_terrain above is the reference-
// shaped per-lane kernel on the binary BVH: second implementation for the
// parity tests, hzb_debug_option("shadow_kernel", 1).)
// ---------------------------------------------------------------------------
constexpr int TW_BLOCK = 512;   // cells per warp work block

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

// The packet step of hzb_wq2.cuh in single-ray mode: folded decode bias, per-lane pending lists,
// shared-diagonal quad test.  SORT: nearest hit child first.
template <bool SW, bool SORT>
__global__ void __launch_bounds__(WQ_BLOCK, 6) k_terrain_wq2(SceneView sv, TerrainParams tp, float sunx, float suny, float sunz,
                                                            uint8_t* __restrict__ shadow, float* __restrict__ swc,
                                                            Counters* counters, unsigned int* block_counter, int refill_thr,
                                                            int wait_thr, int stack_lim) {
    __shared__ Wq2Shared sh;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
    const unsigned int FULL = 0xffffffffu, lt_mask = (1u << lane) - 1u;
    const long long ncell = (long long)tp.dim_in_0 * tp.dim_in_1;
    const unsigned int num_blocks = (unsigned int)((ncell + TW_BLOCK - 1) / TW_BLOCK);
    const float dot_min = SW ? tp.dot_prod_min : 0.0f;
    const float tfar = INFINITY;   // shadow_comp.cpp:462, 572
    LaneCounters cnt; cnt.rays = cnt.nodes = cnt.prims = 0;
    unsigned int units = 0;
    if (lane == 0) { sh.hit1[warp] = 0u; sh.hit2[warp] = 0u; }
    unsigned int pend_est = 0;
    __syncwarp();

    long long next = 0, end = 0;          // warp-uniform: unassigned cells of the current block
    bool more_blocks = true;
    long long cell = -1; float dts = 0.f, dns = 0.f;
    Wq2Lane L; L.state = 0; L.hit1 = L.hit2 = false; L.node = WQ_NONE; L.sp = 0; L.pc = 0;
    L.A1x = L.A1y = L.A1z = L.B1x = L.B1y = L.B1z = 0.f; L.A2x = L.A2y = L.A2z = L.B2x = L.B2y = L.B2z = 0.f;
    L.selxy = 0x74107410u;

    while (true) {
        // (1) retire finished rays (shadow_comp.cpp:468-472 / 578-586) and hand out new cells
        if (cell >= 0 && L.state == 0) {
            if (L.node == WQ_OVF) {   // full stack: the cell is left to the fix-up launch of k_terrain (same decision)
                L.node = WQ_NONE;
                if (SW) swc[cell] = __uint_as_float(HZB_REDO_F32); else shadow[cell] = HZB_REDO_U8;
                atomicAdd(&counters->fallback_packets, 1ull);
            } else if (SW) {
                if (L.hit1) swc[cell] = 0.0f;
                else { if (dns < dot_min) dns = dot_min; swc[cell] = __fmul_rn(__fdiv_rn(dts, dns), tp.surf_enl_fac[cell]); }
            } else shadow[cell] = L.hit1 ? 2 : 0;
            cell = -1;
        }
        while (true) {
            const bool want = L.state == 0;
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
                        wq2_start(sv, sh, warp, lane, L, org, sun, sun);
                        cell = mine; cnt.rays++;
                    } else {
                        if (SW) swc[mine] = 0.0f; else shadow[mine] = 1;   // self-shaded (:474-478 / :588-592)
                    }
                }
            }
        }
        if (__ballot_sync(FULL, L.state != 0) == 0u) break;
        const int thr = (more_blocks || next < end) ? refill_thr : 1;
        __syncwarp();
        // (2) shared warp-queue traversal (hzb_wq2.cuh)
        while (__popc(wq2_step<false, SORT>(sv, sh, warp, lane, tid, L, pend_est, tfar, wait_thr, cnt, stack_lim)) >= thr) {}
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

template <bool SW>
static int launch_terrain(Scene& s, const TerrainParams& tp, const float* sun, uint8_t* d_shadow, float* d_swc, cudaStream_t st) {
    const long long ncell = (long long)tp.dim_in_0 * tp.dim_in_1;
    if (ncell <= 0) return 0;
    const int kind = debug_options().shadow_kernel;
    if (kind == 1) {
        k_terrain<SW, false><<<(unsigned int)((ncell + 127) / 128), 128, 0, st>>>(s.view(), tp, sun[0], sun[1], sun[2], d_shadow, d_swc, s.d_counters);
    } else {
        unsigned int* block_counter = scene_tile_counter(s, st);
        if (!block_counter) return 1;
        const int stack_lim = std::max(1, std::min(debug_options().stack_limit, WQ_STACK_N));
        if (kind == 2) k_terrain_wq2<SW, true><<<sm_count() * 6, WQ_BLOCK, 0, st>>>(s.view(), tp, sun[0], sun[1], sun[2], d_shadow, d_swc, s.d_counters, block_counter, 24, 3, stack_lim);
        else k_terrain_wq2<SW, false><<<sm_count() * 6, WQ_BLOCK, 0, st>>>(s.view(), tp, sun[0], sun[1], sun[2], d_shadow, d_swc, s.d_counters, block_counter, 24, 3, stack_lim);
        // cells whose traversal stack was full (none in practice) are recomputed by the binary-BVH walker
        k_terrain<SW, true><<<(unsigned int)((ncell + 127) / 128), 128, 0, st>>>(s.view(), tp, sun[0], sun[1], sun[2], d_shadow, d_swc, s.d_counters);
    }
    HZB_CUDA(cudaGetLastError());
    return 0;
}

int launch_shadow(Scene& s, const TerrainParams& tp, const float* sun, uint8_t* d_out, cudaStream_t st) {
    return launch_terrain<false>(s, tp, sun, d_out, nullptr, st);
}
int launch_sw_dir_cor(Scene& s, const TerrainParams& tp, const float* sun, float* d_out, cudaStream_t st) {
    return launch_terrain<true>(s, tp, sun, nullptr, d_out, st);
}

}  // namespace hzb
