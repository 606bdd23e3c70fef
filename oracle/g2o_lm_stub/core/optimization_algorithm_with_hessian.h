// oracle/g2o_lm_stub -- TEST INFRASTRUCTURE.  Stand-ins for the headers that g2o's OWN Levenberg-Marquardt source
// (Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.{h,cpp}) and robust-kernel source (robust_kernel.cpp,
// robust_kernel_impl.cpp) include besides their own class declarations: the abstract faces of the solver and of the
// optimiser -- exactly the member functions those two .cpp files call -- a property map, batch statistics that are
// switched off, a clock.  They let the reference's unmodified optimization_algorithm_levenberg.cpp be compiled and
// EXECUTED with the oracle's linear algebra behind it (oracle/ref_harness_lm.cpp, oracle/Makefile target `ref`):
// the pin of the oracle's restatement of solve(), computeLambdaInit() and computeScale().  Written for this
// repository; these files shadow the real headers through `-I... -I-` and are not derived from g2o's sources.
#ifndef VILBA_G2O_LM_STUB_WITH_HESSIAN_H
#define VILBA_G2O_LM_STUB_WITH_HESSIAN_H
#include <cassert>
#include <cmath>
#include <cstddef>
#include <iostream>
#include <limits>
#include <string>
#include <vector>

namespace g2o {

template <typename T>
class Property {
public:
    Property(const std::string&, const T& v) : _v(v) {}
    const T& value() const { return _v; }
    void setValue(const T& v) { _v = v; }

private:
    T _v;
};
class PropertyMap {
public:
    template <typename P, typename T>
    P* makeProperty(const std::string& name, const T& v) {
        return new P(name, v);  // (leaked on purpose: lives as long as the algorithm object of a test)
    }
};

namespace OptimizableGraph {
class Vertex {
public:
    virtual ~Vertex() {}
    virtual int dimension() const = 0;
    virtual double hessian(int i, int j) const = 0;
};
typedef std::vector<Vertex*> VertexContainer;
}  // namespace OptimizableGraph

class SparseOptimizer {
public:
    virtual ~SparseOptimizer() {}
    virtual void computeActiveErrors() = 0;
    virtual double activeRobustChi2() const = 0;
    virtual void push() = 0;
    virtual void pop() = 0;
    virtual void discardTop() = 0;
    virtual void update(const double* x) = 0;
    virtual bool terminate() = 0;
    virtual const OptimizableGraph::VertexContainer& indexMapping() const = 0;
};

class Solver {
public:
    virtual ~Solver() {}
    virtual bool buildStructure(bool zeroBlocks = false) = 0;
    virtual bool buildSystem() = 0;
    virtual bool setLambda(double lambda, bool backup = false) = 0;
    virtual void restoreDiagonal() = 0;
    virtual bool solve() = 0;
    virtual double* x() = 0;
    virtual double* b() = 0;
    virtual size_t vectorSize() const = 0;
    virtual bool schur() = 0;
    SparseOptimizer* optimizer() const { return _optimizer; }
    void setOptimizer(SparseOptimizer* o) { _optimizer = o; }

protected:
    SparseOptimizer* _optimizer = nullptr;
};

class OptimizationAlgorithm {
public:
    enum SolverResult { Terminate = 2, OK = 1, Fail = -1 };
    virtual ~OptimizationAlgorithm() {}
    virtual SolverResult solve(int iteration, bool online = false) = 0;
    virtual void printVerbose(std::ostream&) const {}
    void setOptimizer(SparseOptimizer* o) { _optimizer = o; }

protected:
    SparseOptimizer* _optimizer = nullptr;
    PropertyMap _properties;
};

class OptimizationAlgorithmWithHessian : public OptimizationAlgorithm {
public:
    explicit OptimizationAlgorithmWithHessian(Solver* solver) : _solver(solver) {}

protected:
    Solver* _solver;
};
}  // namespace g2o
#endif
