// hzb_hd.cuh -- host/device portability of the pure-arithmetic product sources.
//
// hzb_search.cuh (search state machine) and hzb_tri.cuh (ray/triangle tests) are compiled for the
// device by the CUDA kernels and, unchanged, for the host by the CPU test infrastructure, which
// checks them against the specification.  On the host the explicitly rounded intrinsics are plain
// IEEE operations (the host build uses -ffp-contract=off, so only fmaf() fuses).
#pragma once
#include <math.h>
#include <float.h>
#include <string.h>

#ifdef __CUDACC__
#define HZB_HD __device__ __forceinline__
#else
#define HZB_HD inline
#include <algorithm>
namespace hzb {
using std::max; using std::min;
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline float __double2float_rn(double a) { return (float)a; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(unsigned int i) { float f; memcpy(&f, &i, 4); return f; }
static inline float __fdividef(float a, float b) { return a / b; }
struct uint4 { unsigned int x, y, z, w; };
// PRMT, default mode: result byte k = byte (selector nibble k & 7) of the 8-byte value {y, x} (x = bytes 0..3)
static inline unsigned int __byte_perm(unsigned int x, unsigned int y, unsigned int s) {
    const unsigned long long v = ((unsigned long long)y << 32) | x;
    unsigned int r = 0;
    for (int k = 0; k < 4; ++k) r |= (unsigned int)((v >> (8 * ((s >> (4 * k)) & 7))) & 0xFF) << (8 * k);
    return r;
}
}  // namespace hzb
#endif

namespace hzb {
// child references of the compressed 4-wide BVH (hzb_common.cuh: Bvh4Node)
constexpr unsigned int WIDE_EMPTY = 0xFFFFFFFFu;   // no child (box inverted)
constexpr unsigned int WIDE_LEAF = 0x80000000u;    // bit 31: leaf holding primitive (ref & 0x7FFFFFFF)
}  // namespace hzb
