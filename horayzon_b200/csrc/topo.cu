// topo.cu -- azimuthal integrals over the horizon array (B200, sm_100a).
//
// Replaces _sky_view_factor_cy, _visible_sky_fraction_cy and
// _topographic_openness_cy (topo_param.pyx:412-460, 499-543, 577-603), which
// are single-threaded loops in the reference.  One warp per cell: lanes read 32
// consecutive azimuths (one 128-byte line), evaluate the per-term expression in
// double like the reference's libm calls, and the warp reduces in double (the
// reference accumulates in a float that is rounded every iteration; the
// difference is below its own 1.2e-6 self-noise).
#include "hzb_common.cuh"
#include <math.h>

namespace hzb {
namespace {

constexpr int SVF_THREADS = 256;

template <int KIND>  // 0 SVF, 1 VSF, 2 openness
__global__ void __launch_bounds__(SVF_THREADS) k_integral(const float* __restrict__ azim, const float* __restrict__ hori,
                                                          const float* __restrict__ tilt, long long cells, int K,
                                                          float* __restrict__ out) {
    extern __shared__ float sh[];  // azim_sin[K], azim_cos[K]
    float* as = sh; float* ac = sh + K;
    if (KIND != 2) {
        for (int k = threadIdx.x; k < K; k += blockDim.x) {  // topo_param.pyx:427-429 (double sin/cos, float store)
            as[k] = (float)sin((double)azim[k]);
            ac[k] = (float)cos((double)azim[k]);
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const float spac = (KIND != 2) ? __fsub_rn(azim[1], azim[0]) : 0.f;  // :433
    for (long long c = warp0; c < cells; c += nwarps) {
        const float* h = hori + c * K;
        float tx = 0.f, ty = 0.f, tz = 1.f;
        if (KIND != 2) { tx = tilt[3 * c]; ty = tilt[3 * c + 1]; tz = tilt[3 * c + 2]; }
        double agg = 0.0;
        for (int k = lane; k < K; k += 32) {
            const float hk = __ldg(h + k);
            if (KIND == 2) {
                agg += (M_PI / 2.0) - (double)hk;                                  // :600
            } else {
                const float a = __fsub_rn(__fdiv_rn(__fmul_rn(-as[k], tx), tz), __fdiv_rn(__fmul_rn(ac[k], ty), tz));
                const float hp = (float)atan((double)a);                           // :442-445
                const float he = (hk >= hp) ? hk : hp;                             // :446-449
                if (KIND == 0) {
                    double sn, cs;
                    sincos((double)he, &sn, &cs);
                    const float w = __fadd_rn(__fmul_rn(tx, as[k]), __fmul_rn(ty, ac[k]));
                    // sin(2h)/2 = sin h cos h
                    agg += (double)w * ((M_PI / 2.0) - (double)he - sn * cs) + (double)tz * cs * cs;  // :452-456
                } else {
                    agg += 1.0 - cos((M_PI / 2.0) - (double)he);                   // :539
                }
            }
        }
        for (int o = 16; o > 0; o >>= 1) agg += __shfl_xor_sync(0xffffffffu, agg, o);
        if (lane == 0) {
            if (KIND == 2) out[c] = __fdiv_rn((float)agg, (float)K);               // :601
            else out[c] = (float)(((double)spac / (2.0 * M_PI)) * (double)(float)agg);  // :458, :541
        }
    }
}

}  // namespace

int launch_svf(int kind, const float* d_azim, const float* d_hori, const float* d_tilt, long long cells, int K,
               float* d_out, cudaStream_t st) {
    if (cells <= 0) return 0;
    if (K < 1 || (kind != 2 && K < 2)) { set_error("azimuthal integrals need at least 2 azimuths"); return 1; }
    const size_t smem = (size_t)2 * K * sizeof(float);
    if (smem > 200 * 1024) { set_error("too many azimuths for the integral kernel"); return 1; }
    const long long warps_needed = cells;
    long long blocks = (warps_needed * 32 + SVF_THREADS - 1) / SVF_THREADS;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (kind == 0) {
        if (smem > 48 * 1024) HZB_CUDA(cudaFuncSetAttribute(k_integral<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_integral<0><<<(unsigned int)blocks, SVF_THREADS, smem, st>>>(d_azim, d_hori, d_tilt, cells, K, d_out);
    } else if (kind == 1) {
        if (smem > 48 * 1024) HZB_CUDA(cudaFuncSetAttribute(k_integral<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_integral<1><<<(unsigned int)blocks, SVF_THREADS, smem, st>>>(d_azim, d_hori, d_tilt, cells, K, d_out);
    } else {
        k_integral<2><<<(unsigned int)blocks, SVF_THREADS, 0, st>>>(d_azim, d_hori, d_tilt, cells, K, d_out);
    }
    HZB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace hzb
