// oracle/g2o_lm_stub -- TEST INFRASTRUCTURE: declared in optimization_algorithm_with_hessian.h of this directory.
#include "optimization_algorithm_with_hessian.h"
