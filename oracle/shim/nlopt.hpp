// Oracle build shim: NLopt is only used by the reference's breakpoint-detection
// code (scicone/src/MathOp.cpp:818-903), which is outside the hot path and is
// never executed by the oracle binaries.  This stub only lets MathOp.cpp compile.
// Test infrastructure only.
#ifndef ORACLE_SHIM_NLOPT_HPP
#define ORACLE_SHIM_NLOPT_HPP
#include <stdexcept>
#include <vector>
namespace nlopt {
enum algorithm { LN_BOBYQA = 34 };
enum result { SUCCESS = 1 };
typedef double (*vfunc)(const std::vector<double>&, std::vector<double>&, void*);
class opt {
public:
    opt(algorithm, unsigned) {}
    void set_lower_bounds(const std::vector<double>&) {}
    void set_upper_bounds(const std::vector<double>&) {}
    void set_min_objective(vfunc, void*) {}
    void set_max_objective(vfunc, void*) {}
    void set_xtol_rel(double) {}
    // Leaves x at its starting point: enough for the reference's own unit-test
    // runner to get through test_breakpoint_detection (which asserts nothing).
    result optimize(std::vector<double>&, double& f) { f = 0.0; return SUCCESS; }
};
}  // namespace nlopt
#endif
