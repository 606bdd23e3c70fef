// oracle/g2o_lm_stub -- TEST INFRASTRUCTURE (see core/optimization_algorithm_with_hessian.h): the clock and the two
// macros of g2o/stuff/{timeutil.h,macros.h} that optimization_algorithm_levenberg.cpp uses.
#ifndef VILBA_G2O_LM_STUB_TIMEUTIL_H
#define VILBA_G2O_LM_STUB_TIMEUTIL_H
#include <cmath>
#include <iomanip>
namespace g2o {
inline double get_monotonic_time() { return 0.0; }
}  // namespace g2o
#define g2o_isfinite(x) std::isfinite(x)
#define FIXED(s) std::fixed << s << std::resetiosflags(std::ios_base::fixed)
#endif
