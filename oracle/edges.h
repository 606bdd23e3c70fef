// oracle/edges.h -- TEST INFRASTRUCTURE (CPU oracle), not part of the product path.
//
// Restatement of the NavState type (src/IMU/NavState.h:14-140, src/IMU/NavState.cpp:81-109) and of
// the three g2o edge classes on the hot path (src/IMU/g2otypes.h:553-706,
// src/IMU/g2otypes.cpp:529-699,703-734,738-788) plus RobustKernelHuber
// (Thirdparty/g2o/g2o/core/robust_kernel_impl.cpp:65-91).  The edges and the vertex updates are pinned against the
// reference's own classes, executed (tests/test_oracle_edges_vs_ref.py); the Huber kernel is restated.
#pragma once
#include "preint.h"

namespace oracle {

struct NavState {
    Vec3 P, V;
    SO3 R;
    Vec3 bg, ba, dbg, dba;

    void load(const double* s) {  // layout: include/vilba.h VILBA_NS_DOUBLES
        P = Vec3::from(s);
        V = Vec3::from(s + 3);
        R.q.w = s[6];
        R.q.x = s[7];
        R.q.y = s[8];
        R.q.z = s[9];
        bg = Vec3::from(s + 10);
        ba = Vec3::from(s + 13);
        dbg = Vec3::from(s + 16);
        dba = Vec3::from(s + 19);
    }
    void store(double* s) const {
        P.store(s);
        V.store(s + 3);
        s[6] = R.q.w;
        s[7] = R.q.x;
        s[8] = R.q.y;
        s[9] = R.q.z;
        bg.store(s + 10);
        ba.store(s + 13);
        dbg.store(s + 16);
        dba.store(s + 19);
    }
    SO3 Get_R() const { return R.copy(); }                    // NavState.h:27-30 (returns by value)
    Mat3 Get_RotMatrix() const { return R.matrix(); }         // NavState.h:32-35
    void IncSmallPVR(const double* u) {                       // NavState.cpp:81-96
        P += vec3(u[0], u[1], u[2]);
        V += vec3(u[3], u[4], u[5]);
        SO3 dR = so3_exp(vec3(u[6], u[7], u[8]));
        R = Get_R().mul(dR);
    }
    void IncSmallBias(const double* u) {                      // NavState.cpp:100-109
        dbg += vec3(u[0], u[1], u[2]);
        dba += vec3(u[3], u[4], u[5]);
    }
};

// RobustKernelHuber::robustify (robust_kernel_impl.cpp:78-91).  dsqr = delta * delta is set by setDelta (:65-69) into a
// FLOAT member in this g2o (robust_kernel_impl.h:84: `float dsqr;`), so the inlier test and rho(e) see delta^2 rounded to
// single precision -- found by executing the reference's kernel (tests/test_oracle_lm_vs_ref.py::test_huber_kernel).
inline void huber(double e, double delta, double rho[3]) {
    const double dsqr = (double)(float)(delta * delta);
    if (e <= dsqr) {
        rho[0] = e;
        rho[1] = 1.;
        rho[2] = 0.;
    } else {
        double sqrte = std::sqrt(e);
        rho[0] = 2 * sqrte * delta - dsqr;
        rho[1] = delta / sqrte;
        rho[2] = -0.5 * rho[1] / e;
    }
}

struct Calib {
    double fx, fy, cx, cy;
    Mat3 Rbc;
    Vec3 Pbc;
};

// ------------------------------------------------------------------------------------------------
// EdgeNavStatePVRPointXYZ (g2otypes.h:616-706, g2otypes.cpp:738-788)
// ------------------------------------------------------------------------------------------------
inline Vec3 mono_compute_pc(const NavState& ns, const Vec3& Pw, const Calib& c) {  // g2otypes.h:650-665
    Mat3 Rwb = ns.Get_RotMatrix();
    Mat3 Rcb = transpose(c.Rbc);
    return (Rcb * transpose(Rwb)) * (Pw - ns.P) - Rcb * c.Pbc;
}

inline void mono_error(const NavState& ns, const Vec3& Pw, const Calib& c, double u, double v, double e[2]) {
    Vec3 Pc = mono_compute_pc(ns, Pw, c);  // g2otypes.h:636-642,667-684
    double px = Pc[0] / Pc[2];
    double py = Pc[1] / Pc[2];
    e[0] = u - (px * c.fx + c.cx);
    e[1] = v - (py * c.fy + c.cy);
}

inline bool mono_depth_positive(const NavState& ns, const Vec3& Pw, const Calib& c) {  // g2otypes.h:644-648
    return mono_compute_pc(ns, Pw, c)[2] > 0.0;
}

// Jacobians: Jpoint 2x3 (_jacobianOplusXi), Jpvr 2x9 ordered [P,V,Phi] (_jacobianOplusXj)
inline void mono_linearize(const NavState& ns, const Vec3& Pw, const Calib& c, Mat<2, 3>& Jpoint,
                           Mat<2, 9>& Jpvr) {  // g2otypes.cpp:738-788
    Mat3 Rwb = ns.Get_RotMatrix();
    Mat3 Rcb = transpose(c.Rbc);
    Vec3 Pc = (Rcb * transpose(Rwb)) * (Pw - ns.P) - Rcb * c.Pbc;
    double x = Pc[0], y = Pc[1], z = Pc[2];
    Mat<2, 3> Maux = Mat<2, 3>::zero();
    Maux(0, 0) = c.fx;
    Maux(0, 1) = 0;
    Maux(0, 2) = -x / z * c.fx;
    Maux(1, 0) = 0;
    Maux(1, 1) = c.fy;
    Maux(1, 2) = -y / z * c.fy;
    Mat<2, 3> Jpi;
    for (int i = 0; i < 6; ++i) Jpi.a[i] = Maux.a[i] / z;  // Eigen: Maux / z divides each coefficient
    Jpoint = ((-Jpi) * Rcb) * transpose(Rwb);
    Mat<2, 3> JdPwb = (-Jpi) * ((-Rcb) * transpose(Rwb));
    Vec3 Paux = (Rcb * transpose(Rwb)) * (Pw - ns.P);
    Mat<2, 3> JdRwb = (-Jpi) * (hat(Paux) * Rcb);
    Jpvr = Mat<2, 9>::zero();
    set_block(Jpvr, 0, 0, JdPwb);
    set_block(Jpvr, 0, 6, JdRwb);
}

// ------------------------------------------------------------------------------------------------
// EdgeNavStatePVR (g2otypes.h:553-586, g2otypes.cpp:529-699); vertices (PVR_i, PVR_j, Bias_i)
// ------------------------------------------------------------------------------------------------
inline Mat<9, 1> pvr_error(const NavState& NSPVRi, const NavState& NSPVRj, const NavState& NSBiasi,
                           const Preintegrator& M, const Vec3& g) {  // g2otypes.cpp:529-585
    Vec3 Pi = NSPVRi.P, Vi = NSPVRi.V;
    SO3 Ri = NSPVRi.Get_R();
    Vec3 dBgi = NSBiasi.dbg, dBai = NSBiasi.dba;
    Vec3 Pj = NSPVRj.P, Vj = NSPVRj.V;
    SO3 Rj = NSPVRj.Get_R();
    double dTij = M.dt;
    double dT2 = dTij * dTij;
    SO3 dRij = SO3::from_matrix(M.dR);
    SO3 RiT = Ri.inverse();
    Vec3 rPij = RiT.rotate(Pj - Pi - Vi * dTij - (0.5 * g) * dT2) - (M.dP + M.JPg * dBgi + M.JPa * dBai);
    Vec3 rVij = RiT.rotate(Vj - Vi - g * dTij) - (M.dV + M.JVg * dBgi + M.JVa * dBai);
    SO3 dR_dbg = so3_exp(M.JRg * dBgi);
    SO3 rRij = dRij.mul(dR_dbg).inverse().mul(RiT).mul(Rj);
    Vec3 rPhiij = so3_log(rRij);
    Mat<9, 1> err;
    for (int k = 0; k < 3; ++k) {
        err[k] = rPij[k];
        err[3 + k] = rVij[k];
        err[6 + k] = rPhiij[k];
    }
    return err;
}

// Jacobians wrt (PVR_i 9, PVR_j 9, Bias_i 6); `err` is the cached _error (g2otypes.cpp:619)
inline void pvr_linearize(const NavState& NSPVRi, const NavState& NSPVRj, const NavState& NSBiasi,
                          const Preintegrator& M, const Vec3& g, const Mat<9, 1>& err, Mat9& JPVRi,
                          Mat9& JPVRj, Mat<9, 6>& JBiasi) {  // g2otypes.cpp:587-699
    Vec3 Pi = NSPVRi.P, Vi = NSPVRi.V;
    Mat3 Ri = NSPVRi.Get_RotMatrix();
    Vec3 dBgi = NSBiasi.dbg;
    Vec3 Pj = NSPVRj.P, Vj = NSPVRj.V;
    Mat3 Rj = NSPVRj.Get_RotMatrix();
    double dTij = M.dt;
    double dT2 = dTij * dTij;
    Mat3 RiT = transpose(Ri);
    Mat3 RjT = transpose(Rj);
    Vec3 rPhiij = vec3(err[6], err[7], err[8]);
    Mat3 JrInv_rPhi = jacobian_r_inv(rPhiij);
    Mat3 J_rPhi_dbg = M.JRg;

    JPVRi = Mat9::zero();
    set_block(JPVRi, 0, 0, -RiT);
    set_block(JPVRi, 0, 3, (-RiT) * dTij);
    set_block(JPVRi, 0, 6, hat(RiT * (Pj - Pi - Vi * dTij - (0.5 * g) * dT2)));
    set_block(JPVRi, 3, 3, -RiT);
    set_block(JPVRi, 3, 6, hat(RiT * (Vj - Vi - g * dTij)));
    set_block(JPVRi, 6, 6, ((-JrInv_rPhi) * RjT) * Ri);

    JPVRj = Mat9::zero();
    set_block(JPVRj, 0, 0, RiT);
    set_block(JPVRj, 3, 3, RiT);
    set_block(JPVRj, 6, 6, JrInv_rPhi);

    JBiasi = Mat<9, 6>::zero();
    Mat3 ExprPhiijTrans = so3_exp(rPhiij).inverse().matrix();
    Mat3 JrBiasGCorr = jacobian_r(J_rPhi_dbg * dBgi);
    set_block(JBiasi, 0, 0, -M.JPg);
    set_block(JBiasi, 0, 3, -M.JPa);
    set_block(JBiasi, 3, 0, -M.JVg);
    set_block(JBiasi, 3, 3, -M.JVa);
    set_block(JBiasi, 6, 0, (((-JrInv_rPhi) * ExprPhiijTrans) * JrBiasGCorr) * J_rPhi_dbg);
}

// ------------------------------------------------------------------------------------------------
// EdgeNavStateBias (g2otypes.h:589-613, g2otypes.cpp:703-734); J_i = -I6, J_j = +I6
// ------------------------------------------------------------------------------------------------
inline Mat<6, 1> bias_error(const NavState& NSi, const NavState& NSj) {
    Vec3 rBiasG = (NSj.bg + NSj.dbg) - (NSi.bg + NSi.dbg);
    Vec3 rBiasA = (NSj.ba + NSj.dba) - (NSi.ba + NSi.dba);
    Mat<6, 1> e;
    for (int k = 0; k < 3; ++k) {
        e[k] = rBiasG[k];
        e[3 + k] = rBiasA[k];
    }
    return e;
}

}  // namespace oracle
