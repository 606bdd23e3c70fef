// oracle/g2o_stub -- TEST INFRASTRUCTURE (see base_vertex.h in this directory).
#ifndef VILBA_G2O_STUB_BASE_MULTI_EDGE_H
#define VILBA_G2O_STUB_BASE_MULTI_EDGE_H
#include "base_vertex.h"
namespace g2o {
// a Jacobian slot of a multi-edge: g2o keeps dynamically sized maps; the edges only ever assign fixed-size matrices to them
class StubDynMatrix {
public:
    int rows() const { return _r; }
    int cols() const { return _c; }
    double operator()(int i, int j) const { return _d[(size_t)i * _c + j]; }
    template <int R, int C>
    StubDynMatrix& operator=(const Matrix<double, R, C>& m) {
        _r = R, _c = C;
        _d.resize((size_t)R * C);
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < C; ++j) _d[(size_t)i * C + j] = m(i, j);
        return *this;
    }

private:
    int _r = 0, _c = 0;
    std::vector<double> _d;
};

template <int D, typename E>
class BaseMultiEdge : public StubEdge<D, E> {
public:
    typedef StubDynMatrix JacobianType;
    void resize(size_t n) {
        this->_vertices.resize(n, nullptr);
        _jacobianOplus.resize(n);
    }
    const std::vector<JacobianType>& jacobianOplus() const { return _jacobianOplus; }

protected:
    std::vector<JacobianType> _jacobianOplus;
};
}  // namespace g2o
#endif
