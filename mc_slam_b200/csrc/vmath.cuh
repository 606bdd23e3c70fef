// vmath.cuh -- FP64 small-matrix / SO(3) device math for the VI local-BA kernels (sm_100a).
//
// Device-side counterpart of the reference's Sophus::SO3 + Eigen quaternion calls
// (src/IMU/so3.cpp:87-271, src/IMU/IMUPreintegrator.h:102-174).  Written for registers: 3-vectors
// and 3x3 matrices are plain structs, products are fully unrolled, nothing is dynamically indexed.
#pragma once
#include <cuda_runtime.h>

namespace vilba {

struct V3 {
    double x, y, z;
};
struct M3 {  // row-major
    double a00, a01, a02, a10, a11, a12, a20, a21, a22;
};
struct Q4 {  // unit quaternion (w, x, y, z)
    double w, x, y, z;
};

#define VD __device__ __forceinline__

VD V3 v3(double x, double y, double z) { return V3{x, y, z}; }
VD V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
VD V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
VD V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
VD V3 operator*(V3 a, double s) { return V3{a.x * s, a.y * s, a.z * s}; }
VD V3 operator*(double s, V3 a) { return V3{a.x * s, a.y * s, a.z * s}; }
VD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
VD V3 cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
VD double norm(V3 a) { return sqrt(dot(a, a)); }
VD V3 ld3(const double* p) { return V3{p[0], p[1], p[2]}; }
VD void st3(double* p, V3 v) {
    p[0] = v.x;
    p[1] = v.y;
    p[2] = v.z;
}

VD M3 m3_identity() { return M3{1, 0, 0, 0, 1, 0, 0, 0, 1}; }
VD M3 m3_zero() { return M3{0, 0, 0, 0, 0, 0, 0, 0, 0}; }
VD M3 ldm3(const double* p) { return M3{p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8]}; }
VD void stm3(double* p, const M3& m) {
    p[0] = m.a00, p[1] = m.a01, p[2] = m.a02;
    p[3] = m.a10, p[4] = m.a11, p[5] = m.a12;
    p[6] = m.a20, p[7] = m.a21, p[8] = m.a22;
}
VD M3 transpose(const M3& m) { return M3{m.a00, m.a10, m.a20, m.a01, m.a11, m.a21, m.a02, m.a12, m.a22}; }
VD M3 operator*(const M3& a, const M3& b) {
    M3 o;
    o.a00 = a.a00 * b.a00 + a.a01 * b.a10 + a.a02 * b.a20;
    o.a01 = a.a00 * b.a01 + a.a01 * b.a11 + a.a02 * b.a21;
    o.a02 = a.a00 * b.a02 + a.a01 * b.a12 + a.a02 * b.a22;
    o.a10 = a.a10 * b.a00 + a.a11 * b.a10 + a.a12 * b.a20;
    o.a11 = a.a10 * b.a01 + a.a11 * b.a11 + a.a12 * b.a21;
    o.a12 = a.a10 * b.a02 + a.a11 * b.a12 + a.a12 * b.a22;
    o.a20 = a.a20 * b.a00 + a.a21 * b.a10 + a.a22 * b.a20;
    o.a21 = a.a20 * b.a01 + a.a21 * b.a11 + a.a22 * b.a21;
    o.a22 = a.a20 * b.a02 + a.a21 * b.a12 + a.a22 * b.a22;
    return o;
}
// a^T * b
VD M3 mul_tn(const M3& a, const M3& b) { return transpose(a) * b; }
VD V3 operator*(const M3& a, V3 v) {
    return V3{a.a00 * v.x + a.a01 * v.y + a.a02 * v.z, a.a10 * v.x + a.a11 * v.y + a.a12 * v.z,
              a.a20 * v.x + a.a21 * v.y + a.a22 * v.z};
}
// a^T * v
VD V3 mul_t(const M3& a, V3 v) {
    return V3{a.a00 * v.x + a.a10 * v.y + a.a20 * v.z, a.a01 * v.x + a.a11 * v.y + a.a21 * v.z,
              a.a02 * v.x + a.a12 * v.y + a.a22 * v.z};
}
VD M3 operator*(const M3& a, double s) {
    return M3{a.a00 * s, a.a01 * s, a.a02 * s, a.a10 * s, a.a11 * s, a.a12 * s, a.a20 * s, a.a21 * s, a.a22 * s};
}
VD M3 operator+(const M3& a, const M3& b) {
    return M3{a.a00 + b.a00, a.a01 + b.a01, a.a02 + b.a02, a.a10 + b.a10, a.a11 + b.a11,
              a.a12 + b.a12, a.a20 + b.a20, a.a21 + b.a21, a.a22 + b.a22};
}
VD M3 operator-(const M3& a, const M3& b) {
    return M3{a.a00 - b.a00, a.a01 - b.a01, a.a02 - b.a02, a.a10 - b.a10, a.a11 - b.a11,
              a.a12 - b.a12, a.a20 - b.a20, a.a21 - b.a21, a.a22 - b.a22};
}
VD M3 operator-(const M3& a) { return a * -1.0; }
VD M3 hat(V3 v) { return M3{0, -v.z, v.y, v.z, 0, -v.x, -v.y, v.x, 0}; }  // so3.cpp:263-271

// ---- quaternions: Eigen's algorithms as the reference calls them ------------------------------
VD Q4 q_normalized(Q4 q) {
    double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    return Q4{q.w / n, q.x / n, q.y / n, q.z / n};
}
VD Q4 q_mul(Q4 a, Q4 b) {
    return Q4{a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
              a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
VD Q4 q_conj(Q4 q) { return Q4{q.w, -q.x, -q.y, -q.z}; }
VD M3 q_to_matrix(Q4 q) {
    const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    return M3{1.0 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1.0 - (txx + tzz),
              tyz - twx, txz - twy, tyz + twx, 1.0 - (txx + tyy)};
}
VD Q4 q_from_matrix(const M3& m) {  // Shoemake, branch order of Eigen's Quaternion(Matrix3)
    double t = m.a00 + m.a11 + m.a22;
    Q4 q;
    if (t > 0.0) {
        t = sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (m.a21 - m.a12) * t;
        q.y = (m.a02 - m.a20) * t;
        q.z = (m.a10 - m.a01) * t;
    } else if (m.a00 >= m.a11 && m.a00 >= m.a22) {  // i = 0
        t = sqrt(m.a00 - m.a11 - m.a22 + 1.0);
        q.x = 0.5 * t;
        t = 0.5 / t;
        q.w = (m.a21 - m.a12) * t;
        q.y = (m.a10 + m.a01) * t;
        q.z = (m.a20 + m.a02) * t;
    } else if (m.a11 > m.a00 && m.a11 >= m.a22) {  // i = 1
        t = sqrt(m.a11 - m.a22 - m.a00 + 1.0);
        q.y = 0.5 * t;
        t = 0.5 / t;
        q.w = (m.a02 - m.a20) * t;
        q.z = (m.a21 + m.a12) * t;
        q.x = (m.a01 + m.a10) * t;
    } else {  // i = 2
        t = sqrt(m.a22 - m.a00 - m.a11 + 1.0);
        q.z = 0.5 * t;
        t = 0.5 / t;
        q.w = (m.a10 - m.a01) * t;
        q.x = (m.a02 + m.a20) * t;
        q.y = (m.a12 + m.a21) * t;
    }
    return q;
}

// ---- Sophus::SO3 ------------------------------------------------------------------------------
VD Q4 so3_exp(V3 omega) {  // so3.cpp:237-261 (+ normalisation of the SO3(Quaterniond) ctor, :105-109)
    double theta = norm(omega);
    double half_theta = 0.5 * theta;
    double imag, real;
    if (theta < 1e-10) {
        double t2 = theta * theta;
        imag = 0.5 - 0.0208333 * t2 + 0.000260417 * (t2 * t2);
        real = cos(half_theta);
    } else {
        double s;
        sincos(half_theta, &s, &real);
        imag = s / theta;
    }
    return q_normalized(Q4{real, imag * omega.x, imag * omega.y, imag * omega.z});
}
VD V3 so3_log(Q4 q) {  // so3.cpp:190-228; always 2*atan(n/w)/n (the +-pi branch is dead code)
    double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
    double w = q.w;
    double f;
    if (n < 1e-10)
        f = 2. / w - 2. * (n * n) / (w * (w * w));
    else
        f = 2 * atan(n / w) / n;
    return V3{f * q.x, f * q.y, f * q.z};
}
// SO3 * SO3 as the reference evaluates it: copy (normalise), product, normalise (so3.cpp:93-96,127-133)
VD Q4 so3_mul(Q4 a, Q4 b) { return q_normalized(q_mul(q_normalized(a), b)); }
VD Q4 so3_inverse(Q4 a) { return q_normalized(q_conj(a)); }  // so3.cpp:149-152
VD V3 q_rotate(Q4 q, V3 v) {                                // Eigen _transformVector
    V3 qv = v3(q.x, q.y, q.z);
    V3 uv = cross(qv, v);
    uv = uv + uv;
    return v + uv * q.w + cross(qv, uv);
}

VD M3 jacobian_r(V3 w) {  // so3.cpp:33-50
    double theta = norm(w);
    if (theta < 0.00001) return m3_identity();
    V3 k = v3(w.x / theta, w.y / theta, w.z / theta);
    M3 K = hat(k);
    double s, c;
    sincos(theta, &s, &c);
    return m3_identity() - K * ((1 - c) / theta) + (K * (1 - s / theta)) * K;
}
VD M3 jacobian_r_inv(V3 w) {  // so3.cpp:53-72
    double theta = norm(w);
    if (theta < 0.00001) return m3_identity();
    V3 k = v3(w.x / theta, w.y / theta, w.z / theta);
    M3 K = hat(k);
    double s, c;
    sincos(theta, &s, &c);
    return m3_identity() + hat(w) * 0.5 + (K * (1.0 - (1.0 + c) * theta / (2.0 * s))) * K;
}

// symmetric 3x3 (H_ll + lambda I) inverse through the adjugate; `ok` false if det == 0
struct S3 {
    double xx, xy, xz, yy, yz, zz;
};
VD S3 s3_inverse(const S3& a, bool& ok) {
    double c00 = a.yy * a.zz - a.yz * a.yz;
    double c01 = a.xz * a.yz - a.xy * a.zz;
    double c02 = a.xy * a.yz - a.xz * a.yy;
    double det = a.xx * c00 + a.xy * c01 + a.xz * c02;
    ok = (det != 0.0);
    double id = 1.0 / det;
    S3 o;
    o.xx = c00 * id;
    o.xy = c01 * id;
    o.xz = c02 * id;
    o.yy = (a.xx * a.zz - a.xz * a.xz) * id;
    o.yz = (a.xy * a.xz - a.xx * a.yz) * id;
    o.zz = (a.xx * a.yy - a.xy * a.xy) * id;
    return o;
}
VD V3 s3_mul(const S3& a, V3 v) {
    return V3{a.xx * v.x + a.xy * v.y + a.xz * v.z, a.xy * v.x + a.yy * v.y + a.yz * v.z,
              a.xz * v.x + a.yz * v.y + a.zz * v.z};
}

// warp helpers
VD double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
VD double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace vilba
