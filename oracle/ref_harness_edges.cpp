// oracle/ref_harness_edges.cpp -- TEST INFRASTRUCTURE.  C entry points around the reference's OWN factor classes,
// compiled from the unmodified /root/reference/src/IMU/g2otypes.cpp (with so3 / IMUPreintegrator / NavState / imudata)
// against the Eigen stand-in (oracle/eigen_stub) and the g2o base-class stand-in (oracle/g2o_stub); oracle/Makefile
// target `ref`, output oracle/_ref/libref_imu.so.  Nothing here restates reference arithmetic: the functions build the
// reference's vertices and edges, call setVertex / setMeasurement / SetParams / computeError / linearizeOplus /
// oplusImpl exactly as src/Optimizer.cpp:2405-2639 and g2o's optimiser do, and copy _error and the Jacobians out
// (row-major).  They are the pin of oracle/edges.h (EdgeNavStatePVR, EdgeNavStateBias, EdgeNavStatePVRPointXYZ,
// VertexNavStatePVR, VertexNavStateBias: src/IMU/g2otypes.h:480-706, g2otypes.cpp:500-788).
#include <cstdint>

#include "IMU/g2otypes.h"

using namespace ORB_SLAM2;

namespace {
Sophus::SO3 so3_of(const double q[4]) { return Sophus::SO3(Eigen::Quaterniond(q[0], q[1], q[2], q[3])); }
NavState ns_of(const double s[22]) {  // 22 doubles as in include/vilba.h: P, V, q(w,x,y,z), bg, ba, dbg, dba
    NavState n;
    n.Set_Pos(Eigen::Vector3d(s[0], s[1], s[2]));
    n.Set_Vel(Eigen::Vector3d(s[3], s[4], s[5]));
    n.Set_Rot(so3_of(s + 6));
    n.Set_BiasGyr(Eigen::Vector3d(s[10], s[11], s[12]));
    n.Set_BiasAcc(Eigen::Vector3d(s[13], s[14], s[15]));
    n.Set_DeltaBiasGyr(Eigen::Vector3d(s[16], s[17], s[18]));
    n.Set_DeltaBiasAcc(Eigen::Vector3d(s[19], s[20], s[21]));
    return n;
}
void put_ns(const NavState& n, double s[22]) {
    const Eigen::Vector3d P = n.Get_P(), V = n.Get_V(), bg = n.Get_BiasGyr(), ba = n.Get_BiasAcc(), dbg = n.Get_dBias_Gyr(),
                          dba = n.Get_dBias_Acc();
    const Eigen::Quaterniond q = n.Get_R().unit_quaternion();
    for (int i = 0; i < 3; ++i) s[i] = P(i), s[3 + i] = V(i), s[10 + i] = bg(i), s[13 + i] = ba(i), s[16 + i] = dbg(i), s[19 + i] = dba(i);
    s[6] = q.w(), s[7] = q.x(), s[8] = q.y(), s[9] = q.z();
}
template <class M>
void put_rows(const M& m, int R, int C, double* o) {
    for (int i = 0; i < R; ++i)
        for (int j = 0; j < C; ++j) o[C * i + j] = m(i, j);
}
// the measurement of the IMU edges: the reference's pre-integrator fed with samples, as KeyFrame::ComputePreInt does
IMUPreintegrator preint_of(int32_t n, const double* gyro, const double* acc, const double* dt, const double* bg, const double* ba) {
    IMUPreintegrator p;
    p.reset();
    const Eigen::Vector3d vbg(bg[0], bg[1], bg[2]), vba(ba[0], ba[1], ba[2]);
    for (int32_t s = 0; s < n; ++s)
        p.update(Eigen::Vector3d(gyro[3 * s], gyro[3 * s + 1], gyro[3 * s + 2]) - vbg,
                 Eigen::Vector3d(acc[3 * s], acc[3 * s + 1], acc[3 * s + 2]) - vba, dt[s]);
    return p;
}
}  // namespace

extern "C" {

// EdgeNavStatePVR(PVR_i, PVR_j, Bias_i) with the pre-integration of `n` samples as measurement.
// err 9, Ji 9x9, Jj 9x9, Jb 9x6 (row-major)
void ref_edge_pvr(int32_t n, const double* gyro, const double* acc, const double* dt, const double* bg, const double* ba,
                  const double ns_i[22], const double ns_j[22], const double ns_bias_i[22], const double g[3], double* err,
                  double* Ji, double* Jj, double* Jb) {
    g2o::VertexNavStatePVR vi, vj;
    g2o::VertexNavStateBias vb;
    vi.setEstimate(ns_of(ns_i)), vj.setEstimate(ns_of(ns_j)), vb.setEstimate(ns_of(ns_bias_i));
    g2o::EdgeNavStatePVR e;
    e.setVertex(0, &vi), e.setVertex(1, &vj), e.setVertex(2, &vb);
    e.setMeasurement(preint_of(n, gyro, acc, dt, bg, ba));
    e.SetParams(Eigen::Vector3d(g[0], g[1], g[2]));
    e.computeError();
    e.linearizeOplus();
    put_rows(e.error(), 9, 1, err);
    put_rows(e.jacobianOplus()[0], 9, 9, Ji);
    put_rows(e.jacobianOplus()[1], 9, 9, Jj);
    put_rows(e.jacobianOplus()[2], 9, 6, Jb);
}

// EdgeNavStateBias(Bias_i, Bias_j): err 6, Ji 6x6, Jj 6x6
void ref_edge_bias(const double ns_i[22], const double ns_j[22], double* err, double* Ji, double* Jj) {
    g2o::VertexNavStateBias vi, vj;
    vi.setEstimate(ns_of(ns_i)), vj.setEstimate(ns_of(ns_j));
    g2o::EdgeNavStateBias e;
    e.setVertex(0, &vi), e.setVertex(1, &vj);
    e.computeError();
    e.linearizeOplus();
    put_rows(e.error(), 6, 1, err);
    put_rows(e.jacobianOplusXi(), 6, 6, Ji);
    put_rows(e.jacobianOplusXj(), 6, 6, Jj);
}

// EdgeNavStatePVRPointXYZ(point, PVR): calib = fx, fy, cx, cy, Rbc (row-major 9), Pbc (3); err 2, Jpoint 2x3, Jpvr 2x9
void ref_edge_mono(const double ns[22], const double pw[3], const double calib[16], const double uv[2], double* err,
                   double* Jpoint, double* Jpvr, int32_t* depth_positive) {
    g2o::VertexSBAPointXYZ vp;
    vp.setEstimate(Eigen::Vector3d(pw[0], pw[1], pw[2]));
    g2o::VertexNavStatePVR vn;
    vn.setEstimate(ns_of(ns));
    g2o::EdgeNavStatePVRPointXYZ e;
    e.setVertex(0, &vp), e.setVertex(1, &vn);
    e.setMeasurement(Eigen::Vector2d(uv[0], uv[1]));
    Eigen::Matrix3d Rbc;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Rbc(i, j) = calib[4 + 3 * i + j];
    e.SetParams(calib[0], calib[1], calib[2], calib[3], Rbc, Eigen::Vector3d(calib[13], calib[14], calib[15]));
    e.computeError();
    e.linearizeOplus();
    put_rows(e.error(), 2, 1, err);
    put_rows(e.jacobianOplusXi(), 2, 3, Jpoint);
    put_rows(e.jacobianOplusXj(), 2, 9, Jpvr);
    *depth_positive = e.isDepthPositive() ? 1 : 0;
}

// VertexNavStatePVR::oplusImpl / VertexNavStateBias::oplusImpl (what g2o's update() calls with the solver's increment)
void ref_vertex_pvr_oplus(double s[22], const double d[9]) {
    g2o::VertexNavStatePVR v;
    v.setEstimate(ns_of(s));
    v.oplus(d);
    put_ns(v.estimate(), s);
}
void ref_vertex_bias_oplus(double s[22], const double d[6]) {
    g2o::VertexNavStateBias v;
    v.setEstimate(ns_of(s));
    v.oplus(d);
    put_ns(v.estimate(), s);
}
}
