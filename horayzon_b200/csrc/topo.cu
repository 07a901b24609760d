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


// ---- slope (topo_param.pyx:84-225, 284-372): one lane per cell
__device__ __forceinline__ void solve3_pp(float A[3][3], float b[3]) {   // LU with partial pivoting (sgesv)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int piv = c;
#pragma unroll
        for (int r = c + 1; r < 3; ++r) if (fabsf(A[r][c]) > fabsf(A[piv][c])) piv = r;
        if (piv != c) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { const float t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
            const float t = b[c]; b[c] = b[piv]; b[piv] = t;
        }
#pragma unroll
        for (int r = c + 1; r < 3; ++r) {
            const float f = __fdiv_rn(A[r][c], A[c][c]);
#pragma unroll
            for (int k = c; k < 3; ++k) A[r][k] = __fsub_rn(A[r][k], __fmul_rn(f, A[c][k]));
            b[r] = __fsub_rn(b[r], __fmul_rn(f, b[c]));
        }
    }
#pragma unroll
    for (int r = 2; r >= 0; --r) {
        float v = b[r];
#pragma unroll
        for (int k = r + 1; k < 3; ++k) v = __fsub_rn(v, __fmul_rn(A[r][k], b[k]));
        b[r] = __fdiv_rn(v, A[r][r]);
    }
}
__device__ __forceinline__ float dot3s(float a0, float a1, float a2, float b0, float b1, float b2) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1)), __fmul_rn(a2, b2));
}

template <int METH>  // 0 plane fit, 1 four-triangle average
__global__ void __launch_bounds__(256) k_slope(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z,
                                               const float* __restrict__ rot, int ny, int nx, int output_rot, float* __restrict__ out) {
    const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (c >= (long long)ny * nx) return;
    const int i = (int)(c / nx), j = (int)(c - (long long)i * nx);
    float vx = NAN, vy = NAN, vz = NAN;
    if (i >= 1 && i < ny - 1 && j >= 1 && j < nx - 1) {
        float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        if (rot) { for (int k = 0; k < 9; ++k) R[k] = rot[9 * c + k]; }
        const float x0 = x[c], y0 = y[c], z0 = z[c];
        if (METH == 0) {
            float sx = 0, sy = 0, sz = 0, sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0;
            for (int k = -1; k <= 1; ++k)
                for (int l = -1; l <= 1; ++l) {
                    const long long q = c + (long long)k * nx + l;
                    const float dx = __fsub_rn(x[q], x0), dy = __fsub_rn(y[q], y0), dz = __fsub_rn(z[q], z0);
                    const float lx = dot3s(R[0], R[1], R[2], dx, dy, dz), ly = dot3s(R[3], R[4], R[5], dx, dy, dz);
                    const float lz = dot3s(R[6], R[7], R[8], dx, dy, dz);
                    sx = __fadd_rn(sx, lx); sy = __fadd_rn(sy, ly); sz = __fadd_rn(sz, lz);
                    sxx = __fadd_rn(sxx, __fmul_rn(lx, lx)); sxy = __fadd_rn(sxy, __fmul_rn(lx, ly));
                    sxz = __fadd_rn(sxz, __fmul_rn(lx, lz)); syy = __fadd_rn(syy, __fmul_rn(ly, ly));
                    syz = __fadd_rn(syz, __fmul_rn(ly, lz));
                }
            float A[3][3] = {{sxx, sxy, sx}, {sxy, syy, sy}, {sx, sy, 9.0f}};
            float b[3] = {sxz, syz, sz};
            solve3_pp(A, b);
            vx = b[0]; vy = b[1]; vz = -1.0f;
        } else {
            const long long qa = c - 1, qb = c + nx, qc = c + 1, qd = c - nx;
            const float a0 = __fsub_rn(x[qa], x0), a1 = __fsub_rn(y[qa], y0), a2 = __fsub_rn(z[qa], z0);
            const float b0 = __fsub_rn(x[qb], x0), b1 = __fsub_rn(y[qb], y0), b2 = __fsub_rn(z[qb], z0);
            const float c0 = __fsub_rn(x[qc], x0), c1 = __fsub_rn(y[qc], y0), c2 = __fsub_rn(z[qc], z0);
            const float d0 = __fsub_rn(x[qd], x0), d1 = __fsub_rn(y[qd], y0), d2 = __fsub_rn(z[qd], z0);
#define HZB_CR(p, q, r, s_) __fsub_rn(__fmul_rn(p, q), __fmul_rn(r, s_))
            vx = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(HZB_CR(a1, b2, a2, b1), HZB_CR(b1, c2, b2, c1)), HZB_CR(c1, d2, c2, d1)), HZB_CR(d1, a2, d2, a1)), 0.25f);
            vy = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(HZB_CR(a2, b0, a0, b2), HZB_CR(b2, c0, b0, c2)), HZB_CR(c2, d0, c0, d2)), HZB_CR(d2, a0, d0, a2)), 0.25f);
            vz = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(HZB_CR(a0, b1, a1, b0), HZB_CR(b0, c1, b1, c0)), HZB_CR(c0, d1, c1, d0)), HZB_CR(d0, a1, d1, a0)), 0.25f);
#undef HZB_CR
        }
        const float mag = __fsqrt_rn(dot3s(vx, vy, vz, vx, vy, vz));
        vx = __fdiv_rn(vx, mag); vy = __fdiv_rn(vy, mag); vz = __fdiv_rn(vz, mag);
        if (vz < 0.0f) { vx = -vx; vy = -vy; vz = -vz; }
        if (METH == 0 && !output_rot) {           // back to the input frame (transpose)
            const float tx = dot3s(R[0], R[3], R[6], vx, vy, vz), ty = dot3s(R[1], R[4], R[7], vx, vy, vz);
            const float tz = dot3s(R[2], R[5], R[8], vx, vy, vz);
            vx = tx; vy = ty; vz = tz;
        } else if (METH == 1 && output_rot && rot) {
            const float tx = dot3s(R[0], R[1], R[2], vx, vy, vz), ty = dot3s(R[3], R[4], R[5], vx, vy, vz);
            const float tz = dot3s(R[6], R[7], R[8], vx, vy, vz);
            vx = tx; vy = ty; vz = tz;
        }
    }
    out[3 * c] = vx; out[3 * c + 1] = vy; out[3 * c + 2] = vz;
}

}  // namespace

int launch_slope(int method, const float* d_x, const float* d_y, const float* d_z, const float* d_rot, int ny, int nx,
                 int output_rot, float* d_out, cudaStream_t st) {
    const long long n = (long long)ny * nx;
    if (n <= 0) return 0;
    const unsigned int blocks = (unsigned int)((n + 255) / 256);
    if (method == 0) k_slope<0><<<blocks, 256, 0, st>>>(d_x, d_y, d_z, d_rot, ny, nx, output_rot, d_out);
    else k_slope<1><<<blocks, 256, 0, st>>>(d_x, d_y, d_z, d_rot, ny, nx, output_rot, d_out);
    HZB_CUDA(cudaGetLastError());
    return 0;
}

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
