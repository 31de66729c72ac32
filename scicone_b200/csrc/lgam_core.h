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
    const double r = fma(z, invc, -1.0);
    const double kd = (double)k;
    const double w = fma(kd, SCICONE_LN2_HI, logc_hi);  // exact: both are multiples of 2^-42 below 2^10
    const double hi = w + r;
    double lo = (w - hi) + r;
    lo += fma(kd, SCICONE_LN2_LO, logc_lo);
    const double r2 = r * r;
    double p = fma(r, -1.0 / 6.0, 1.0 / 5.0);
    p = fma(r, p, -1.0 / 4.0);
    p = fma(r, p, 1.0 / 3.0);
    p = fma(r, p, -0.5);
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
    r = r * fma(-x, r, 2.0);
    const double r2 = r * r;
    double p = fma(r2, 1.0 / 1188.0, -1.0 / 1680.0);
    p = fma(r2, p, 1.0 / 1260.0);
    p = fma(r2, p, -1.0 / 360.0);
    p = fma(r2, p, 1.0 / 12.0);
    const double t = fma(x - 0.5, lx, -x);
    return t + fma(p, r, 0.91893853320467274178);
}

}  // namespace scicone
