// oracle/g2o_stub -- TEST INFRASTRUCTURE (see core/base_vertex.h).  Of g2o's types_six_dof_expmap.h / types_sba.h the
// reference's g2otypes only uses the landmark vertex: a 3-vector estimate with additive update.
#ifndef VILBA_G2O_STUB_TYPES_SIX_DOF_EXPMAP_H
#define VILBA_G2O_STUB_TYPES_SIX_DOF_EXPMAP_H
#include "../core/base_vertex.h"
namespace g2o {
class VertexSBAPointXYZ : public BaseVertex<3, Vector3d> {
public:
    virtual bool read(std::istream&) { return true; }
    virtual bool write(std::ostream&) const { return true; }
    virtual void setToOriginImpl() { _estimate = Vector3d(0.0, 0.0, 0.0); }
    virtual void oplusImpl(const double* u) { _estimate = _estimate + Vector3d(u[0], u[1], u[2]); }
};
}  // namespace g2o
#endif
