// oracle/g2o_stub -- TEST INFRASTRUCTURE (see base_vertex.h in this directory).
#ifndef VILBA_G2O_STUB_BASE_UNARY_EDGE_H
#define VILBA_G2O_STUB_BASE_UNARY_EDGE_H
#include "base_vertex.h"
namespace g2o {
template <int D, typename E, typename VertexXi>
class BaseUnaryEdge : public StubEdge<D, E> {
public:
    typedef Matrix<double, D, VertexXi::Dimension> JacobianXiOplusType;
    BaseUnaryEdge() { this->_vertices.resize(1, nullptr); }
    const JacobianXiOplusType& jacobianOplusXi() const { return _jacobianOplusXi; }

protected:
    JacobianXiOplusType _jacobianOplusXi;
};
}  // namespace g2o
#endif
