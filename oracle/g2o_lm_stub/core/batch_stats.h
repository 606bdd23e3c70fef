// oracle/g2o_lm_stub -- TEST INFRASTRUCTURE (see optimization_algorithm_with_hessian.h): statistics switched off.
#ifndef VILBA_G2O_LM_STUB_BATCH_STATS_H
#define VILBA_G2O_LM_STUB_BATCH_STATS_H
namespace g2o {
struct G2OBatchStatistics {
    double timeResiduals = 0, timeQuadraticForm = 0, timeLinearSolution = 0, timeUpdate = 0;
    int levenbergIterations = 0;
    static G2OBatchStatistics* globalStats() { return nullptr; }
};
}  // namespace g2o
#endif
