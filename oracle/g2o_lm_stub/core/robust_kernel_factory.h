// oracle/g2o_lm_stub -- TEST INFRASTRUCTURE (see optimization_algorithm_with_hessian.h): no kernel registry.
#ifndef VILBA_G2O_LM_STUB_ROBUST_KERNEL_FACTORY_H
#define VILBA_G2O_LM_STUB_ROBUST_KERNEL_FACTORY_H
#define G2O_REGISTER_ROBUST_KERNEL(name, classname)
#endif
