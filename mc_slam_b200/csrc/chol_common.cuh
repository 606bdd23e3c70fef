// chol_common.cuh -- device helpers shared by the reduced-system factorisation kernels (chol_la.cu, chol_big.cu).
#pragma once
#include <cfloat>

#include "kernels.h"

namespace vilba {

// ~1 ulp reciprocal: MUFU seed + two Newton steps (shorter dependent chain than an IEEE division)
__device__ __forceinline__ double fast_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    return fma(r, e, r);
}

// SimplicialLDLT semantics (g2o/solvers/linear_solver_eigen.h:104-111): the factorisation goes on through NEGATIVE
// pivots and only reports failure for a pivot that is exactly zero (or not finite).
__device__ __forceinline__ bool pivot_bad(double d) { return !(fabs(d) > 0.0 && fabs(d) <= DBL_MAX); }

// LDL^T of an NB x NB block (NB = 16 or 32) held one row per lane (lane r: a[c] = A(r,c) for c <= r; rows/cols
// beyond the real size are identity-padded by the caller).  On return lane r holds the unit-lower l(r,c) in
// a[c], c < r, and d_r in a[r].  `sm` is 16 + 4 NB doubles of warp-private shared memory.  Returns false if a
// pivot was zero / not finite.
template <int NB>
__device__ __forceinline__ bool warp_ldlt_mb4(double (&a)[NB], int lane, double* sm) {
    double* Bc = sm;       // 4 x 4  pivot block
    double* Lb = sm + 16;  // NB x 4 scaled rows l(r, j..j+3)
    bool ok = true;
#pragma unroll
    for (int s = 0; s < NB / 4; ++s) {
        const int j = 4 * s;
        if (lane >= j && lane < j + 4) {
            double* p = Bc + 4 * (lane - j);
            p[0] = a[j], p[1] = a[j + 1], p[2] = a[j + 2], p[3] = a[j + 3];
        }
        __syncwarp();
        const double b00 = Bc[0], b10 = Bc[4], b20 = Bc[8], b30 = Bc[12];
        double b11 = Bc[5], b21 = Bc[9], b31 = Bc[13], b22 = Bc[10], b32 = Bc[14], b33 = Bc[15];
        // 4x4 LDL^T, redundantly in every lane
        const double d0 = b00, r0 = fast_rcp(d0);
        const double l10 = b10 * r0, l20 = b20 * r0, l30 = b30 * r0;
        b11 -= l10 * b10, b21 -= l20 * b10, b31 -= l30 * b10;
        b22 -= l20 * b20, b32 -= l30 * b20, b33 -= l30 * b30;
        const double d1 = b11, r1 = fast_rcp(d1);
        const double l21 = b21 * r1, l31 = b31 * r1;
        b22 -= l21 * b21, b32 -= l31 * b21, b33 -= l31 * b31;
        const double d2 = b22, r2 = fast_rcp(d2);
        const double l32 = b32 * r2;
        b33 -= l32 * b32;
        const double d3 = b33, r3 = fast_rcp(d3);
        if (pivot_bad(d0) || pivot_bad(d1) || pivot_bad(d2) || pivot_bad(d3)) ok = false;
        // rows below the pivot block: u = A(r, j..j+3) Lb^-T (unscaled), l = u D^-1
        double u0 = a[j], u1 = a[j + 1], u2 = a[j + 2], u3 = a[j + 3];
        u1 -= u0 * l10;
        u2 -= u0 * l20 + u1 * l21;
        u3 -= u0 * l30 + u1 * l31 + u2 * l32;
        const double q0 = u0 * r0, q1 = u1 * r1, q2 = u2 * r2, q3 = u3 * r3;
        if (lane >= j + 4) {
            a[j] = q0, a[j + 1] = q1, a[j + 2] = q2, a[j + 3] = q3;
            if (lane < NB) {
                double* p = Lb + 4 * lane;
                p[0] = q0, p[1] = q1, p[2] = q2, p[3] = q3;
            }
        } else if (lane >= j) {  // rows of the pivot block itself
            const int i = lane - j;
            a[j] = (i == 0) ? d0 : (i == 1) ? l10 : (i == 2) ? l20 : l30;
            a[j + 1] = (i == 1) ? d1 : (i == 2) ? l21 : (i == 3) ? l31 : 0.0;
            a[j + 2] = (i == 2) ? d2 : (i == 3) ? l32 : 0.0;
            a[j + 3] = (i == 3) ? d3 : 0.0;
        }
        if (s < NB / 4 - 1) {
            __syncwarp();
            // trailing part of the block: A(r,c) -= sum_k u(r,k) l(c,k), c = j+4 .. r
#pragma unroll
            for (int c = j + 4; c < NB; ++c) {
                const double* p = Lb + 4 * c;
                if (lane >= c) a[c] -= u0 * p[0] + u1 * p[1] + u2 * p[2] + u3 * p[3];
            }
        }
        __syncwarp();
    }
    return ok;
}

// Lower-triangular 4x4 tiles enumerated column by column: t = off(tj) + (ti - tj), off(tj) = tj (2 trp - tj + 1) / 2,
// so that consecutive threads own consecutive tile ROWS of one tile column (coalesced global access).
__device__ __forceinline__ void decode_tile(int t, int trp, int& ti, int& tj) {
    const float bq = 2.0f * (float)trp + 1.0f;
    int j = (int)((bq - sqrtf(fmaxf(bq * bq - 8.0f * (float)t, 0.0f))) * 0.5f);
    j = max(0, min(j, trp - 1));
    while (j > 0 && j * (2 * trp - j + 1) / 2 > t) --j;
    while (j + 1 < trp && (j + 1) * (2 * trp - j) / 2 <= t) ++j;
    tj = j;
    ti = j + (t - j * (2 * trp - j + 1) / 2);
}

}  // namespace vilba
