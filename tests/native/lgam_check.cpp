// Host build of the device log-gamma (scicone_b200/csrc/lgam_core.h) measured against glibc: prints the worst error in
// ulps of log_pos vs log and of lgam_large vs lgamma over log-spaced and random arguments.  Driven by tests/test_lgam_cpu.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../scicone_b200/csrc/lgam_core.h"

static const double kTable[] = {SCICONE_LOG_TABLE_VALUES};

static double ulps(double got, double want) {
    if (got == want) return 0.0;
    int e;
    frexp(want, &e);
    return fabs(got - want) / ldexp(1.0, e - 53);
}

int main(int argc, char** argv) {
    const long n = argc > 1 ? atol(argv[1]) : 2000000;
    std::mt19937_64 gen(12345);
    double worst_log = 0, worst_lg = 0, at_log = 0, at_lg = 0;
    double sum_lg = 0;
    long bitexact = 0, total = 0;
    for (int decade = 0; decade < 8; ++decade) {
        const double lo = decade == 0 ? 16.0 : 16.0 * pow(10.0, decade), hi = 16.0 * pow(10.0, decade + 1);
        std::uniform_real_distribution<double> u(log(lo), log(hi));
        for (long s = 0; s < n / 8; ++s) {
            double x = exp(u(gen));
            if (s % 3 == 0) x = floor(x) + 0.5 * (s % 2);  // integer and half-integer counts are the common case
            if (x < 16.0) x = 16.0;
            const double e1 = ulps(scicone::log_pos(x, kTable), log(x));
            if (e1 > worst_log) { worst_log = e1; at_log = x; }
            const double got = scicone::lgam_large(x, kTable), want = lgamma(x);
            const double e2 = ulps(got, want);
            if (e2 > worst_lg) { worst_lg = e2; at_lg = x; }
            sum_lg += e2;
            bitexact += got == want;
            ++total;
        }
    }
    printf("{\"n\": %ld, \"log_max_ulp\": %.3f, \"log_at\": %.17g, \"lgamma_max_ulp\": %.3f, \"lgamma_at\": %.17g, "
           "\"lgamma_mean_ulp\": %.4f, \"lgamma_bit_exact_frac\": %.4f}\n",
           total, worst_log, at_log, worst_lg, at_lg, sum_lg / total, (double)bitexact / total);
    return 0;
}
