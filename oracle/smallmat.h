// oracle/smallmat.h -- TEST INFRASTRUCTURE (CPU oracle), not part of the product path.
//
// Minimal fixed-size dense matrices standing in for the Eigen types the reference uses
// (Eigen is an un-vendored dependency of /root/reference and is not present in this image).
// Products are evaluated densely with the inner index running 0..K-1, which is what a plain
// Eigen fixed-size product does up to vectorisation order; zeros/ones in structured operands
// therefore contribute exactly like in the reference.
#pragma once
#include <cmath>
#include <cstring>

namespace oracle {

template <int R, int C>
struct Mat {
    double a[R * C];
    double& operator()(int r, int c) { return a[r * C + c]; }
    const double& operator()(int r, int c) const { return a[r * C + c]; }
    double& operator[](int i) { return a[i]; }
    const double& operator[](int i) const { return a[i]; }
    static Mat zero() {
        Mat m;
        for (int i = 0; i < R * C; ++i) m.a[i] = 0.0;
        return m;
    }
    static Mat identity() {
        Mat m = zero();
        for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = 1.0;
        return m;
    }
    static Mat from(const double* p) {
        Mat m;
        std::memcpy(m.a, p, sizeof(double) * R * C);
        return m;
    }
    void store(double* p) const { std::memcpy(p, a, sizeof(double) * R * C); }
};

typedef Mat<3, 1> Vec3;
typedef Mat<3, 3> Mat3;
typedef Mat<9, 9> Mat9;

template <int R, int K, int C>
inline Mat<R, C> operator*(const Mat<R, K>& A, const Mat<K, C>& B) {
    Mat<R, C> o;
    for (int r = 0; r < R; ++r)
        for (int c = 0; c < C; ++c) {
            double s = 0.0;
            for (int k = 0; k < K; ++k) s += A(r, k) * B(k, c);
            o(r, c) = s;
        }
    return o;
}
template <int R, int C>
inline Mat<R, C> operator+(const Mat<R, C>& A, const Mat<R, C>& B) {
    Mat<R, C> o;
    for (int i = 0; i < R * C; ++i) o.a[i] = A.a[i] + B.a[i];
    return o;
}
template <int R, int C>
inline Mat<R, C> operator-(const Mat<R, C>& A, const Mat<R, C>& B) {
    Mat<R, C> o;
    for (int i = 0; i < R * C; ++i) o.a[i] = A.a[i] - B.a[i];
    return o;
}
template <int R, int C>
inline Mat<R, C> operator-(const Mat<R, C>& A) {
    Mat<R, C> o;
    for (int i = 0; i < R * C; ++i) o.a[i] = -A.a[i];
    return o;
}
template <int R, int C>
inline Mat<R, C> operator*(const Mat<R, C>& A, double s) {
    Mat<R, C> o;
    for (int i = 0; i < R * C; ++i) o.a[i] = A.a[i] * s;
    return o;
}
template <int R, int C>
inline Mat<R, C> operator*(double s, const Mat<R, C>& A) {
    return A * s;
}
template <int R, int C>
inline Mat<R, C>& operator+=(Mat<R, C>& A, const Mat<R, C>& B) {
    for (int i = 0; i < R * C; ++i) A.a[i] += B.a[i];
    return A;
}
template <int R, int C>
inline Mat<R, C>& operator-=(Mat<R, C>& A, const Mat<R, C>& B) {
    for (int i = 0; i < R * C; ++i) A.a[i] -= B.a[i];
    return A;
}
template <int R, int C>
inline Mat<C, R> transpose(const Mat<R, C>& A) {
    Mat<C, R> o;
    for (int r = 0; r < R; ++r)
        for (int c = 0; c < C; ++c) o(c, r) = A(r, c);
    return o;
}
template <int R, int C, int BR, int BC>
inline void set_block(Mat<R, C>& M, int r0, int c0, const Mat<BR, BC>& B) {
    for (int r = 0; r < BR; ++r)
        for (int c = 0; c < BC; ++c) M(r0 + r, c0 + c) = B(r, c);
}
template <int BR, int BC, int R, int C>
inline Mat<BR, BC> get_block(const Mat<R, C>& M, int r0, int c0) {
    Mat<BR, BC> o;
    for (int r = 0; r < BR; ++r)
        for (int c = 0; c < BC; ++c) o(r, c) = M(r0 + r, c0 + c);
    return o;
}
template <int N>
inline double dot(const Mat<N, 1>& a, const Mat<N, 1>& b) {
    double s = 0.0;
    for (int i = 0; i < N; ++i) s += a[i] * b[i];
    return s;
}
inline double norm(const Vec3& v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
inline Vec3 vec3(double x, double y, double z) {
    Vec3 v;
    v[0] = x;
    v[1] = y;
    v[2] = z;
    return v;
}
inline Vec3 cross(const Vec3& a, const Vec3& b) {
    return vec3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}

// General inverse by LU with partial pivoting -- the algorithm behind Eigen's
// Matrix<double,N,N>::inverse() for N > 4 (PartialPivLU) and for dynamic sizes; used by the
// reference at src/Optimizer.cpp:2510 (9x9) and g2o/core/block_solver.hpp:389 (3x3 MatrixXd).
// Returns false on an exactly-zero pivot (Eigen would return inf/nan entries).
template <int N>
inline bool inverse_lu(const Mat<N, N>& A, Mat<N, N>& inv) {
    double lu[N][N];
    int perm[N];
    for (int i = 0; i < N; ++i) {
        perm[i] = i;
        for (int j = 0; j < N; ++j) lu[i][j] = A(i, j);
    }
    bool ok = true;
    for (int k = 0; k < N; ++k) {
        int piv = k;
        double best = std::fabs(lu[k][k]);
        for (int i = k + 1; i < N; ++i)
            if (std::fabs(lu[i][k]) > best) {
                best = std::fabs(lu[i][k]);
                piv = i;
            }
        if (best == 0.0) ok = false;
        if (piv != k) {
            for (int j = 0; j < N; ++j) {
                double t = lu[k][j];
                lu[k][j] = lu[piv][j];
                lu[piv][j] = t;
            }
            int t = perm[k];
            perm[k] = perm[piv];
            perm[piv] = t;
        }
        for (int i = k + 1; i < N; ++i) {
            lu[i][k] /= lu[k][k];
            for (int j = k + 1; j < N; ++j) lu[i][j] -= lu[i][k] * lu[k][j];
        }
    }
    for (int c = 0; c < N; ++c) {
        double y[N];
        for (int i = 0; i < N; ++i) {
            double s = (perm[i] == c) ? 1.0 : 0.0;
            for (int j = 0; j < i; ++j) s -= lu[i][j] * y[j];
            y[i] = s;
        }
        for (int i = N - 1; i >= 0; --i) {
            double s = y[i];
            for (int j = i + 1; j < N; ++j) s -= lu[i][j] * inv(j, c);
            inv(i, c) = s / lu[i][i];
        }
    }
    return ok;
}

}  // namespace oracle
