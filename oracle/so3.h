// oracle/so3.h -- TEST INFRASTRUCTURE (CPU oracle), not part of the product path.
//
// Restatement of Sophus::SO3 as vendored by the reference (src/IMU/so3.h:38-117,
// src/IMU/so3.cpp:33-300) and of the Eigen::Quaterniond operations it calls.  Eigen is NOT
// vendored in /root/reference (CMakeLists.txt:27 find_package(Eigen3), version unpinned, >= 3.1.0);
// the quaternion algorithms below restate Eigen's published ones (Geometry/Quaternion.h):
// Shoemake matrix->quaternion, toRotationMatrix, Hamilton product, normalize, _transformVector.
#pragma once
#include "smallmat.h"

namespace oracle {

static const double SMALL_EPS = 1e-10;  // src/IMU/so3.h:36

struct Quat {
    double w, x, y, z;
};

inline Quat quat_normalized(const Quat& q) {  // Eigen: coeffs() /= norm()
    double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    Quat o = {q.w / n, q.x / n, q.y / n, q.z / n};
    return o;
}

inline Quat quat_mul(const Quat& a, const Quat& b) {  // Eigen quat_product, Hamilton convention
    Quat o;
    o.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    o.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    o.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    o.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return o;
}

inline Quat quat_from_matrix(const Mat3& m) {  // Eigen quaternionbase_assign_impl<Other,3,3>
    Quat q;
    double t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > 0.0) {
        t = std::sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (m(2, 1) - m(1, 2)) * t;
        q.y = (m(0, 2) - m(2, 0)) * t;
        q.z = (m(1, 0) - m(0, 1)) * t;
    } else {
        int i = 0;
        if (m(1, 1) > m(0, 0)) i = 1;
        if (m(2, 2) > m(i, i)) i = 2;
        int j = (i + 1) % 3;
        int k = (j + 1) % 3;
        t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
        double v[3];
        v[i] = 0.5 * t;
        t = 0.5 / t;
        q.w = (m(k, j) - m(j, k)) * t;
        v[j] = (m(j, i) + m(i, j)) * t;
        v[k] = (m(k, i) + m(i, k)) * t;
        q.x = v[0];
        q.y = v[1];
        q.z = v[2];
    }
    return q;
}

inline Mat3 quat_to_matrix(const Quat& q) {  // Eigen QuaternionBase::toRotationMatrix
    const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    Mat3 r;
    r(0, 0) = 1.0 - (tyy + tzz);
    r(0, 1) = txy - twz;
    r(0, 2) = txz + twy;
    r(1, 0) = txy + twz;
    r(1, 1) = 1.0 - (txx + tzz);
    r(1, 2) = tyz - twx;
    r(2, 0) = txz - twy;
    r(2, 1) = tyz + twx;
    r(2, 2) = 1.0 - (txx + tyy);
    return r;
}

inline Vec3 quat_rotate(const Quat& q, const Vec3& v) {  // Eigen QuaternionBase::_transformVector
    Vec3 qv = vec3(q.x, q.y, q.z);
    Vec3 uv = cross(qv, v);
    uv = uv + uv;
    return v + uv * q.w + cross(qv, uv);
}

inline Mat3 hat(const Vec3& v) {  // so3.cpp:263-271
    Mat3 m = Mat3::zero();
    m(0, 1) = -v[2];
    m(0, 2) = v[1];
    m(1, 0) = v[2];
    m(1, 2) = -v[0];
    m(2, 0) = -v[1];
    m(2, 1) = v[0];
    return m;
}

// Sophus::SO3 -- a unit quaternion that is re-normalised by every constructor.
struct SO3 {
    Quat q;
    SO3() {
        q.w = 1.0;
        q.x = q.y = q.z = 0.0;
    }
    static SO3 from_quat(const Quat& qq) {  // so3.cpp:105-109
        SO3 s;
        s.q = quat_normalized(qq);
        return s;
    }
    static SO3 from_matrix(const Mat3& R) {  // so3.cpp:99-102
        return from_quat(quat_from_matrix(R));
    }
    // copy constructor re-normalises (so3.cpp:93-96); used where the reference copies by value
    SO3 copy() const { return from_quat(q); }
    SO3 mul(const SO3& o) const {  // so3.cpp:127-133: copy, *=, normalize
        SO3 r = copy();
        r.q = quat_normalized(quat_mul(r.q, o.q));
        return r;
    }
    Vec3 rotate(const Vec3& v) const { return quat_rotate(q, v); }  // so3.cpp:142-145
    SO3 inverse() const {                                            // so3.cpp:149-152
        Quat c = {q.w, -q.x, -q.y, -q.z};
        return from_quat(c);
    }
    Mat3 matrix() const { return quat_to_matrix(q); }  // so3.cpp:155-158
};

inline SO3 so3_exp(const Vec3& omega) {  // so3.cpp:237-261
    double theta = norm(omega);
    double half_theta = 0.5 * theta;
    double imag_factor;
    double real_factor = std::cos(half_theta);
    if (theta < SMALL_EPS) {
        double theta_sq = theta * theta;
        double theta_po4 = theta_sq * theta_sq;
        imag_factor = 0.5 - 0.0208333 * theta_sq + 0.000260417 * theta_po4;
    } else {
        double sin_half_theta = std::sin(half_theta);
        imag_factor = sin_half_theta / theta;
    }
    Quat q = {real_factor, imag_factor * omega[0], imag_factor * omega[1], imag_factor * omega[2]};
    return SO3::from_quat(q);
}

inline Vec3 so3_log(const SO3& s) {  // so3.cpp:190-228 (the +-pi/n branch at :211-221 is dead code)
    double n = std::sqrt(s.q.x * s.q.x + s.q.y * s.q.y + s.q.z * s.q.z);
    double w = s.q.w;
    double squared_w = w * w;
    double two_atan_nbyw_by_n;
    if (n < SMALL_EPS) {
        two_atan_nbyw_by_n = 2. / w - 2. * (n * n) / (w * squared_w);
    } else {
        two_atan_nbyw_by_n = 2 * std::atan(n / w) / n;
    }
    return vec3(two_atan_nbyw_by_n * s.q.x, two_atan_nbyw_by_n * s.q.y, two_atan_nbyw_by_n * s.q.z);
}

inline Vec3 normalized(const Vec3& w) {
    double n = norm(w);
    return vec3(w[0] / n, w[1] / n, w[2] / n);
}

inline Mat3 jacobian_r(const Vec3& w) {  // so3.cpp:33-50 == IMUPreintegrator.h:102-120
    Mat3 Jr = Mat3::identity();
    double theta = norm(w);
    if (theta < 0.00001) return Jr;
    Vec3 k = normalized(w);
    Mat3 K = hat(k);
    Jr = Mat3::identity() - ((1 - std::cos(theta)) / theta) * K + ((1 - std::sin(theta) / theta) * K) * K;
    return Jr;
}

inline Mat3 jacobian_r_inv(const Vec3& w) {  // so3.cpp:53-72
    Mat3 Jrinv = Mat3::identity();
    double theta = norm(w);
    if (theta < 0.00001) return Jrinv;
    Vec3 k = normalized(w);
    Mat3 K = hat(k);
    Jrinv = Mat3::identity() + 0.5 * hat(w) +
            ((1.0 - (1.0 + std::cos(theta)) * theta / (2.0 * std::sin(theta))) * K) * K;
    return Jrinv;
}

}  // namespace oracle
