// Large-argument log-gamma shared by every kernel (and compiled for the host by tests/ to measure its error against
// glibc's lgamma, the routine the reference calls: scicone/src/Tree.h:214-231, 1783-1797).
//
// Every lgamma on this path has the form lgamma(count + nu * copy_number * region_size) or lgamma(S + nu * z): the
// argument is almost always >= 16, where the Stirling series with five correction terms is below 1 ulp.  Its cost is
// one logarithm and one reciprocal.  The CUDA library's double-precision log + division take ~200 issue slots per call
// (measured: profiles/r1_v2_summary.md); here the logarithm is table driven (x = 2^k z, r = z * invc_i - 1 by one FMA,
// degree-6 polynomial, compensated sum: 15 fp64 operations) and the reciprocal, which only feeds the correction
// terms, is a single-precision estimate refined by one Newton step.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "log_table.h"

#if defined(__CUDACC__)
#define SCICONE_HD __host__ __device__ __forceinline__
#else
#define SCICONE_HD static inline
#endif

namespace scicone {

constexpr int kLogTableEntries = 128;
constexpr int kLogTableDoubles = 3 * kLogTableEntries;

// Polynomial coefficients and split constants.  On the device they live in constant memory: fp64 instructions take a
// constant-bank operand directly, whereas a 64-bit literal costs two move instructions each time it is used (they were a
// third of the instructions of one lgamma).  The host build (tests) reads the same values from a plain array.
#define SCICONE_LGAM_CONSTANTS                                                                                             \
    {-1.0 / 6.0, 1.0 / 5.0, -1.0 / 4.0, 1.0 / 3.0, -0.5, SCICONE_LN2_HI, SCICONE_LN2_LO, 1.0 / 1188.0, -1.0 / 1680.0,    \
     1.0 / 1260.0, -1.0 / 360.0, 1.0 / 12.0, 0.91893853320467274178, 2.0, -1.0, 0.5}
enum { K_L6, K_L5, K_L4, K_L3, K_L2, K_LN2HI, K_LN2LO, K_S9, K_S7, K_S5, K_S3, K_S1, K_HALFLN2PI, K_TWO, K_MINUS1, K_HALF };
#if defined(__CUDACC__)
__constant__ double c_lgam_k[16] = SCICONE_LGAM_CONSTANTS;
#endif
static const double h_lgam_k[16] = SCICONE_LGAM_CONSTANTS;
#if defined(__CUDA_ARCH__)
#define SCICONE_K(i) c_lgam_k[i]
#else
#define SCICONE_K(i) h_lgam_k[i]
#endif

// ln(x) for positive normal x.  Absolute error below 2^-58 + 0.52 ulp of the result for |ln x| >= 0.3.
SCICONE_HD double log_pos(double x, const double* __restrict__ table) {
    uint64_t ix;
#if defined(__CUDA_ARCH__)
    ix = (uint64_t)__double_as_longlong(x);
#else
    memcpy(&ix, &x, 8);
#endif
    const uint64_t tmp = ix - 0x3FE6000000000000ull;
    const int i = (int)((tmp >> 45) & 127u);
    const int k = (int)((int64_t)tmp >> 52);
    const uint64_t iz = ix - (tmp & 0xFFF0000000000000ull);
    double z;
#if defined(__CUDA_ARCH__)
    z = __longlong_as_double((long long)iz);
#else
    memcpy(&z, &iz, 8);
#endif
    const double invc = table[3 * i], logc_hi = table[3 * i + 1], logc_lo = table[3 * i + 2];
    const double r = fma(z, invc, SCICONE_K(K_MINUS1));
    const double kd = (double)k;
    const double w = fma(kd, SCICONE_K(K_LN2HI), logc_hi);  // exact: both are multiples of 2^-42 below 2^10
    const double hi = w + r;
    double lo = (w - hi) + r;
    lo += fma(kd, SCICONE_K(K_LN2LO), logc_lo);
    const double r2 = r * r;
    double p = fma(r, SCICONE_K(K_L6), SCICONE_K(K_L5));
    p = fma(r, p, SCICONE_K(K_L4));
    p = fma(r, p, SCICONE_K(K_L3));
    p = fma(r, p, SCICONE_K(K_L2));
    return hi + fma(r2, p, lo);
}

// lgamma(x) for x >= 16 (Stirling: (x - 1/2) ln x - x + ln sqrt(2 pi) + 1/(12x) - 1/(360x^3) + ...).
SCICONE_HD double lgam_large(double x, const double* __restrict__ table) {
    const double lx = log_pos(x, table);
    // 1/x to ~2^-44: it only enters the correction terms, which are below 1/190 of the result
    const float xf = (float)x;
#if defined(__CUDA_ARCH__)
    float rf;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"(xf));
#else
    const float rf = 1.0f / xf;
#endif
    double r = (double)rf;
    r = r * fma(-x, r, SCICONE_K(K_TWO));
    const double r2 = r * r;
    double p = fma(r2, SCICONE_K(K_S9), SCICONE_K(K_S7));
    p = fma(r2, p, SCICONE_K(K_S5));
    p = fma(r2, p, SCICONE_K(K_S3));
    p = fma(r2, p, SCICONE_K(K_S1));
    const double t = fma(x - SCICONE_K(K_HALF), lx, -x);
    return t + fma(p, r, SCICONE_K(K_HALFLN2PI));
}

}  // namespace scicone
