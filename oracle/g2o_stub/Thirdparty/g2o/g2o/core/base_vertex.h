// oracle/g2o_stub -- TEST INFRASTRUCTURE.  A minimal stand-in for the g2o base classes that the reference's
// src/IMU/g2otypes.{h,cpp} derive from (Thirdparty/g2o/g2o/core/base_vertex.h, base_unary_edge.h, base_binary_edge.h,
// base_multi_edge.h): exactly the members those edge / vertex classes touch -- estimate, measurement, error, the Jacobian
// slots, the vertex list -- and nothing of g2o's graph, solver or robust-kernel machinery.  It exists so that the
// reference's OWN computeError() / linearizeOplus() / oplusImpl() can be compiled unmodified and executed as the pin of
// oracle/edges.h (oracle/Makefile target `ref`).  Written for this repository; not derived from g2o's sources.
#ifndef VILBA_G2O_STUB_BASE_VERTEX_H
#define VILBA_G2O_STUB_BASE_VERTEX_H
#include <Eigen/Core>
#include <cstddef>
#include <iostream>
#include <vector>

namespace g2o {
using namespace Eigen;

namespace HyperGraph {
class Vertex {
public:
    virtual ~Vertex() {}
};
}  // namespace HyperGraph

namespace OptimizableGraph {
class Vertex : public HyperGraph::Vertex {
public:
    virtual void setToOriginImpl() = 0;
    virtual void oplusImpl(const double* update) = 0;
    virtual bool read(std::istream& is) = 0;
    virtual bool write(std::ostream& os) const = 0;
    void oplus(const double* v) { oplusImpl(v); }
    void setId(int id) { _id = id; }
    int id() const { return _id; }
    void setFixed(bool f) { _fixed = f; }
    bool fixed() const { return _fixed; }
    void setMarginalized(bool m) { _marg = m; }
    bool marginalized() const { return _marg; }

protected:
    int _id = 0;
    bool _fixed = false, _marg = false;
};
}  // namespace OptimizableGraph

template <int D, typename T>
class BaseVertex : public OptimizableGraph::Vertex {
public:
    typedef T EstimateType;
    static const int Dimension = D;
    const EstimateType& estimate() const { return _estimate; }
    void setEstimate(const EstimateType& e) { _estimate = e; }

protected:
    EstimateType _estimate;
};

// what every stub edge shares: the vertex list and the bookkeeping g2otypes reads
class StubEdgeBase {
public:
    virtual ~StubEdgeBase() {}
    virtual void computeError() = 0;
    virtual void linearizeOplus() = 0;
    void setVertex(size_t i, HyperGraph::Vertex* v) { _vertices[i] = v; }
    const std::vector<HyperGraph::Vertex*>& vertices() const { return _vertices; }
    HyperGraph::Vertex* vertex(size_t i) const { return _vertices[i]; }
    void setLevel(int l) { _level = l; }
    int level() const { return _level; }

protected:
    std::vector<HyperGraph::Vertex*> _vertices;
    int _level = 0;
};

template <int D, typename E>
class StubEdge : public StubEdgeBase {
public:
    static const int Dimension = D;
    typedef E Measurement;
    typedef Matrix<double, D, 1> ErrorVector;
    typedef Matrix<double, D, D> InformationType;
    virtual bool read(std::istream& is) = 0;
    virtual bool write(std::ostream& os) const = 0;
    const Measurement& measurement() const { return _measurement; }
    void setMeasurement(const Measurement& m) { _measurement = m; }
    const ErrorVector& error() const { return _error; }
    const InformationType& information() const { return _information; }
    void setInformation(const InformationType& i) { _information = i; }

protected:
    Measurement _measurement;
    ErrorVector _error;
    InformationType _information;
};
}  // namespace g2o
#endif
