// oracle/preint.h -- TEST INFRASTRUCTURE (CPU oracle), not part of the product path.
//
// Restatement of ORB_SLAM2::IMUPreintegrator (src/IMU/IMUPreintegrator.h:17-200,
// src/IMU/IMUPreintegrator.cpp:16-112): Forster on-manifold pre-integration with the 9x9
// covariance (order P,V,Phi) and the five bias Jacobians, written with dense 9x9 products the
// way the reference's Eigen expressions evaluate.
#pragma once
#include "so3.h"
#include "../include/vilba.h"

namespace oracle {

struct NoiseModel {  // src/IMU/imudata.cpp:25-37 (static members of IMUData)
    double gyr_meas_cov;  // diagonal value of _gyrMeasCov
    double acc_meas_cov;  // diagonal value of _accMeasCov
};

struct Preintegrator {
    Vec3 dP, dV;
    Mat3 dR;
    Mat3 JPg, JPa, JVg, JVa, JRg;
    Mat9 cov;
    double dt;

    Preintegrator() { reset(); }

    void reset() {  // IMUPreintegrator.cpp:39-56
        dP = Vec3::zero();
        dV = Vec3::zero();
        dR = Mat3::identity();
        JPg = JPa = JVg = JVa = JRg = Mat3::zero();
        cov = Mat9::zero();
        dt = 0.0;
    }

    static Mat3 normalize_rotation(const Mat3& R) {  // IMUPreintegrator.h:156-174
        Quat q = quat_from_matrix(R);
        if (q.w < 0) {
            q.w = -q.w;
            q.x = -q.x;
            q.y = -q.y;
            q.z = -q.z;
        }
        return quat_to_matrix(quat_normalized(q));
    }

    // IMUPreintegrator.cpp:63-112.  omega/acc are already bias-subtracted by the caller.
    void update(const Vec3& omega, const Vec3& acc, double dt_, const NoiseModel& nm) {
        const double dt2 = dt_ * dt_;
        Mat3 dRk = so3_exp(omega * dt_).matrix();  // Expmap, IMUPreintegrator.h:94-97
        Mat3 Jr = jacobian_r(omega * dt_);

        Mat3 I3 = Mat3::identity();
        Mat9 A = Mat9::identity();
        Mat3 acc_hat = hat(acc);
        set_block(A, 6, 6, transpose(dRk));
        set_block(A, 3, 6, ((-dR) * acc_hat) * dt_);
        set_block(A, 0, 6, ((-0.5 * dR) * acc_hat) * dt2);
        set_block(A, 0, 3, I3 * dt_);
        Mat<9, 3> Bg = Mat<9, 3>::zero();
        set_block(Bg, 6, 0, Jr * dt_);
        Mat<9, 3> Ca = Mat<9, 3>::zero();
        set_block(Ca, 3, 0, dR * dt_);
        set_block(Ca, 0, 0, (0.5 * dR) * dt2);
        Mat3 Sg = Mat3::identity() * nm.gyr_meas_cov;
        Mat3 Sa = Mat3::identity() * nm.acc_meas_cov;
        cov = (A * cov) * transpose(A) + (Bg * Sg) * transpose(Bg) + (Ca * Sa) * transpose(Ca);

        // bias Jacobians: each line uses the pre-update values of the ones below it (:98-102)
        JPa += JVa * dt_ - (0.5 * dR) * dt2;
        JPg += JVg * dt_ - (((0.5 * dR) * acc_hat) * JRg) * dt2;
        JVa += (-dR) * dt_;
        JVg += (((-dR) * acc_hat) * JRg) * dt_;
        JRg = transpose(dRk) * JRg - Jr * dt_;

        // deltas (:106-110)
        dP += dV * dt_ + ((0.5 * dR) * acc) * dt2;
        dV += (dR * acc) * dt_;
        dR = normalize_rotation(dR * dRk);
        dt += dt_;
    }

    void store(double* o) const {  // layout: include/vilba.h VILBA_PI_*
        dP.store(o + VILBA_PI_DP);
        dV.store(o + VILBA_PI_DV);
        dR.store(o + VILBA_PI_DR);
        JPg.store(o + VILBA_PI_JPG);
        JPa.store(o + VILBA_PI_JPA);
        JVg.store(o + VILBA_PI_JVG);
        JVa.store(o + VILBA_PI_JVA);
        JRg.store(o + VILBA_PI_JRG);
        cov.store(o + VILBA_PI_COV);
        o[VILBA_PI_DT] = dt;
    }
    void load(const double* o) {
        dP = Vec3::from(o + VILBA_PI_DP);
        dV = Vec3::from(o + VILBA_PI_DV);
        dR = Mat3::from(o + VILBA_PI_DR);
        JPg = Mat3::from(o + VILBA_PI_JPG);
        JPa = Mat3::from(o + VILBA_PI_JPA);
        JVg = Mat3::from(o + VILBA_PI_JVG);
        JVa = Mat3::from(o + VILBA_PI_JVA);
        JRg = Mat3::from(o + VILBA_PI_JRG);
        cov = Mat9::from(o + VILBA_PI_COV);
        dt = o[VILBA_PI_DT];
    }
};

}  // namespace oracle
