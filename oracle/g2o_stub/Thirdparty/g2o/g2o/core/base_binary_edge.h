// oracle/g2o_stub -- TEST INFRASTRUCTURE (see base_vertex.h in this directory).
#ifndef VILBA_G2O_STUB_BASE_BINARY_EDGE_H
#define VILBA_G2O_STUB_BASE_BINARY_EDGE_H
#include "base_vertex.h"
namespace g2o {
template <int D, typename E, typename VertexXi, typename VertexXj>
class BaseBinaryEdge : public StubEdge<D, E> {
public:
    typedef Matrix<double, D, VertexXi::Dimension> JacobianXiOplusType;
    typedef Matrix<double, D, VertexXj::Dimension> JacobianXjOplusType;
    BaseBinaryEdge() { this->_vertices.resize(2, nullptr); }
    const JacobianXiOplusType& jacobianOplusXi() const { return _jacobianOplusXi; }
    const JacobianXjOplusType& jacobianOplusXj() const { return _jacobianOplusXj; }

protected:
    JacobianXiOplusType _jacobianOplusXi;
    JacobianXjOplusType _jacobianOplusXj;
};
}  // namespace g2o
#endif
