// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE.  C entry points around the reference's OWN classes, compiled from the
// unmodified sources under /root/reference/src/IMU (so3.cpp, IMUPreintegrator.cpp, NavState.cpp, imudata.cpp) against
// the Eigen stand-in in oracle/eigen_stub/ (oracle/Makefile target `ref`; output oracle/_ref/libref_imu.so).  Nothing
// in this file restates reference arithmetic: it only constructs the reference objects, calls their public methods in
// the order the reference does (KeyFrame::ComputePreInt, src/KeyFrame.cpp:195-252) and copies the getters out.
#include <cstdint>

#include "IMU/IMUPreintegrator.h"
#include "IMU/NavState.h"
#include "IMU/imudata.h"
#include "IMU/so3.h"

using namespace ORB_SLAM2;

namespace {
void put3(const Eigen::Vector3d& v, double* o) { o[0] = v(0), o[1] = v(1), o[2] = v(2); }
void put33(const Eigen::Matrix3d& m, double* o) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) o[3 * i + j] = m(i, j);
}
Sophus::SO3 so3_of(const double q[4]) { return Sophus::SO3(Eigen::Quaterniond(q[0], q[1], q[2], q[3])); }
void put_q(const Sophus::SO3& r, double q[4]) {
    const Eigen::Quaterniond& u = r.unit_quaternion();
    q[0] = u.w(), q[1] = u.x(), q[2] = u.y(), q[3] = u.z();
}
}  // namespace

extern "C" {

// IMUPreintegrator: reset() then update(gyro - bg, acc - ba, dt) per sample (the caller subtracts the biases exactly
// like KeyFrame.cpp:218-247 does); out = 142 doubles in the layout of include/vilba.h (VILBA_PI_*)
void ref_preintegrate(int32_t n_samples, const double* gyro, const double* acc, const double* dt, const double* bg,
                      const double* ba, double* out) {
    IMUPreintegrator p;
    p.reset();
    const Eigen::Vector3d vbg(bg[0], bg[1], bg[2]), vba(ba[0], ba[1], ba[2]);
    for (int32_t s = 0; s < n_samples; ++s) {
        const Eigen::Vector3d w(gyro[3 * s], gyro[3 * s + 1], gyro[3 * s + 2]);
        const Eigen::Vector3d a(acc[3 * s], acc[3 * s + 1], acc[3 * s + 2]);
        p.update(w - vbg, a - vba, dt[s]);
    }
    put3(p.getDeltaP(), out + 0);
    put3(p.getDeltaV(), out + 3);
    put33(p.getDeltaR(), out + 6);
    put33(p.getJPBiasg(), out + 15);
    put33(p.getJPBiasa(), out + 24);
    put33(p.getJVBiasg(), out + 33);
    put33(p.getJVBiasa(), out + 42);
    put33(p.getJRBiasg(), out + 51);
    const Matrix9d c = p.getCovPVPhi();
    for (int i = 0; i < 9; ++i)
        for (int j = 0; j < 9; ++j) out[60 + 9 * i + j] = c(i, j);
    out[141] = p.getDeltaTime();
}

void ref_so3_exp(const double w[3], double q_wxyz[4]) { put_q(Sophus::SO3::exp(Eigen::Vector3d(w[0], w[1], w[2])), q_wxyz); }
void ref_so3_log(const double q_wxyz[4], double w[3]) { put3(so3_of(q_wxyz).log(), w); }
void ref_so3_mul(const double a[4], const double b[4], double out[4]) { put_q(so3_of(a) * so3_of(b), out); }
void ref_so3_inverse(const double a[4], double out[4]) { put_q(so3_of(a).inverse(), out); }
void ref_so3_matrix(const double a[4], double R[9]) { put33(so3_of(a).matrix(), R); }
void ref_so3_from_matrix(const double R[9], double q[4]) {
    Eigen::Matrix3d m;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) m(i, j) = R[3 * i + j];
    put_q(Sophus::SO3(m), q);
}
void ref_so3_rotate(const double a[4], const double v[3], double out[3]) { put3(so3_of(a) * Eigen::Vector3d(v[0], v[1], v[2]), out); }
void ref_jacobian_r(const double w[3], double J[9]) { put33(IMUPreintegrator::JacobianR(Eigen::Vector3d(w[0], w[1], w[2])), J); }
void ref_jacobian_r_inv(const double w[3], double J[9]) { put33(IMUPreintegrator::JacobianRInv(Eigen::Vector3d(w[0], w[1], w[2])), J); }
void ref_so3_jacobian_r(const double w[3], double J[9]) { put33(Sophus::SO3::JacobianR(Eigen::Vector3d(w[0], w[1], w[2])), J); }
void ref_so3_jacobian_r_inv(const double w[3], double J[9]) { put33(Sophus::SO3::JacobianRInv(Eigen::Vector3d(w[0], w[1], w[2])), J); }

// NavState flattened to 22 doubles as in include/vilba.h: P, V, q(w,x,y,z), bg, ba, dbg, dba
static NavState ns_of(const double s[22]) {
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
static void put_ns(const NavState& n, double s[22]) {
    put3(n.Get_P(), s), put3(n.Get_V(), s + 3);
    put_q(n.Get_R(), s + 6);
    put3(n.Get_BiasGyr(), s + 10), put3(n.Get_BiasAcc(), s + 13), put3(n.Get_dBias_Gyr(), s + 16), put3(n.Get_dBias_Acc(), s + 19);
}
void ref_navstate_inc_pvr(double s[22], const double d[9]) {
    NavState n = ns_of(s);
    Vector9d v;
    for (int i = 0; i < 9; ++i) v(i) = d[i];
    n.IncSmallPVR(v);
    put_ns(n, s);
}
void ref_navstate_inc_bias(double s[22], const double d[6]) {
    NavState n = ns_of(s);
    Vector6d v;
    for (int i = 0; i < 6; ++i) v(i) = d[i];
    n.IncSmallBias(v);
    put_ns(n, s);
}
// IMUData statics: gyr meas cov (0,0), acc meas cov (0,0), gyr bias rw2, acc bias rw2
void ref_imu_constants(double out[4]) {
    out[0] = IMUData::getGyrMeasCov()(0, 0);
    out[1] = IMUData::getAccMeasCov()(0, 0);
    out[2] = IMUData::getGyrBiasRW2();
    out[3] = IMUData::getAccBiasRW2();
}
const char* ref_build_info(void) {
    return "unmodified /root/reference/src/IMU/{so3,IMUPreintegrator,NavState,imudata,g2otypes}.cpp + oracle/eigen_stub + oracle/g2o_stub";
}
}
