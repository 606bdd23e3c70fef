// preint.cu -- K1: batched IMU pre-integration for sm_100a.
//
// Replaces the IMUPreintegrator::reset()+update() loop that KeyFrame::ComputePreInt drives
// (reference: src/IMU/IMUPreintegrator.cpp:63-112, src/KeyFrame.cpp:210-249) for many key-frame
// pairs at once.  FP64 throughout (the covariance spans ~14 orders of magnitude once inverted).
//
// Mapping: a group of G lanes owns one key-frame pair (32/G pairs per warp).
//   * prologue, sample-parallel: lane g of the group turns sample base+g into dR = Exp(w dt),
//     Jr(w dt) and the bias-corrected acceleration -- all the transcendental work (sincos, sqrt,
//     divides) leaves the serial chain -- and parks the 22 doubles in the group's shared-memory slot;
//   * recurrence, entry-parallel: the 81 covariance entries and the 4x9 V/P bias-Jacobian entries
//     are distributed over the G lanes; the 3x3 pieces every entry needs (delta_R, J_R_bg, R*a^)
//     are replicated in registers.  The structured A = [[I, hI, -R a^ h^2/2],[0, I, -R a^ h],
//     [0, 0, dR^T]] is applied block-wise, so no 9x9 product is formed.
// Lanes exchange data through warp-private shared memory with __syncwarp only (no block barriers).
#include "vmath.cuh"
#include "kernels.h"

namespace vilba {

template <int G>
struct PreintSmem {
    // per pair: S (81) | T (81) | chunk of G samples x 22
    static constexpr int kPerPair = 81 + 81 + 22 * G;
};

#define FOR9(M) M(0, a00) M(1, a01) M(2, a02) M(3, a10) M(4, a11) M(5, a12) M(6, a20) M(7, a21) M(8, a22)

template <int G>
__global__ void __launch_bounds__(kPreintThreads)
preint_batch_kernel(int n_pairs, const int* __restrict__ sample_begin, const double* __restrict__ gyro,
                    const double* __restrict__ acc, const double* __restrict__ dt,
                    const double* __restrict__ bg, const double* __restrict__ ba, double* __restrict__ out,
                    double gyr_cov, double acc_cov) {
    constexpr int PPW = 32 / G;
    constexpr int NSLOT = (9 + G - 1) / G;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane % G, gi = lane / G;
    const int pair = (blockIdx.x * (kPreintThreads / 32) + warp) * PPW + gi;
    const bool valid = pair < n_pairs;
    double* sS = smem + (size_t)(warp * PPW + gi) * PreintSmem<G>::kPerPair;
    double* sT = sS + 81;
    double* sC = sT + 81;

    const int s0 = valid ? sample_begin[pair] : 0;
    const int cnt = valid ? sample_begin[pair + 1] - s0 : 0;
    int maxcnt = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxcnt = max(maxcnt, __shfl_xor_sync(0xffffffffu, maxcnt, o));

    const V3 bgv = valid ? ld3(bg + 3 * (size_t)pair) : v3(0, 0, 0);
    const V3 bav = valid ? ld3(ba + 3 * (size_t)pair) : v3(0, 0, 0);

    // replicated state (IMUPreintegrator.cpp:39-56 reset values)
    M3 R = m3_identity(), JRg = m3_zero();
    V3 dP = v3(0, 0, 0), dV = v3(0, 0, 0);
    double T = 0.0;
    // distributed state: entries e = gl + G*slot of J_P_bg, J_P_ba, J_V_bg, J_V_ba
    double jpg[NSLOT], jpa[NSLOT], jvg[NSLOT], jva[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) jpg[s] = jpa[s] = jvg[s] = jva[s] = 0.0;
    for (int e = gl; e < 81; e += G) sS[e] = 0.0;
    __syncwarp();

    for (int base = 0; base < maxcnt; base += G) {
        // ---- prologue: one sample per lane --------------------------------------------------------
        if (base + gl < cnt) {
            const size_t s = (size_t)(s0 + base + gl);
            const double h = dt[s];
            const V3 w = ld3(gyro + 3 * s) - bgv;  // omega = gyro - bg   (KeyFrame.cpp:218,240)
            const V3 a = ld3(acc + 3 * s) - bav;
            const V3 wh = w * h;
            const M3 dRk = q_to_matrix(so3_exp(wh));  // Expmap (IMUPreintegrator.h:94-97)
            const M3 Jr = jacobian_r(wh);
            double* c = sC + 22 * gl;
            stm3(c, dRk);
            stm3(c + 9, Jr);
            st3(c + 18, a);
            c[21] = h;
        }
        __syncwarp();
        // ---- recurrence over the chunk -------------------------------------------------------------
        for (int k = 0; k < G; ++k) {
            const bool act = (base + k) < cnt;  // uniform inside a group
            const double* c = sC + 22 * k;
            M3 dRk = m3_identity(), Jr = m3_identity();
            V3 a = v3(0, 0, 0);
            double h = 0.0;
            if (act) {
                dRk = ldm3(c);
                Jr = ldm3(c + 9);
                a = ld3(c + 18);
                h = c[21];
            }
            const double h2 = h * h;
            const M3 M1 = R * hat(a);              // R a^
            const M3 Avp = M1 * (-h);              // A(3:6,6:9)
            const M3 App = M1 * (-0.5 * h2);       // A(0:3,6:9)
            // phase 1: T = A * Sigma
            if (act) {
                for (int e = gl; e < 81; e += G) {
                    const int r = e / 9, cc = e - 9 * r;
                    const double s6 = sS[54 + cc], s7 = sS[63 + cc], s8 = sS[72 + cc];
                    double t;
                    if (r < 3) {
                        const double p0 = (r == 0) ? App.a00 : (r == 1) ? App.a10 : App.a20;
                        const double p1 = (r == 0) ? App.a01 : (r == 1) ? App.a11 : App.a21;
                        const double p2 = (r == 0) ? App.a02 : (r == 1) ? App.a12 : App.a22;
                        t = sS[e] + h * sS[e + 27] + (p0 * s6 + p1 * s7 + p2 * s8);
                    } else if (r < 6) {
                        const double p0 = (r == 3) ? Avp.a00 : (r == 4) ? Avp.a10 : Avp.a20;
                        const double p1 = (r == 3) ? Avp.a01 : (r == 4) ? Avp.a11 : Avp.a21;
                        const double p2 = (r == 3) ? Avp.a02 : (r == 4) ? Avp.a12 : Avp.a22;
                        t = sS[e] + (p0 * s6 + p1 * s7 + p2 * s8);
                    } else {  // A(6:9,6:9) = dR^T
                        const double p0 = (r == 6) ? dRk.a00 : (r == 7) ? dRk.a01 : dRk.a02;
                        const double p1 = (r == 6) ? dRk.a10 : (r == 7) ? dRk.a11 : dRk.a12;
                        const double p2 = (r == 6) ? dRk.a20 : (r == 7) ? dRk.a21 : dRk.a22;
                        t = p0 * s6 + p1 * s7 + p2 * s8;
                    }
                    sT[e] = t;
                }
            }
            __syncwarp();
            // phase 2: Sigma = T * A^T + Bg Sg Bg^T + Ca Sa Ca^T
            if (act) {
                const M3 RRt = R * transpose(R);
                const M3 JJt = Jr * transpose(Jr);
                for (int e = gl; e < 81; e += G) {
                    const int r = e / 9, cc = e - 9 * r;
                    const double t6 = sT[9 * r + 6], t7 = sT[9 * r + 7], t8 = sT[9 * r + 8];
                    double v;
                    if (cc < 3) {
                        const double p0 = (cc == 0) ? App.a00 : (cc == 1) ? App.a10 : App.a20;
                        const double p1 = (cc == 0) ? App.a01 : (cc == 1) ? App.a11 : App.a21;
                        const double p2 = (cc == 0) ? App.a02 : (cc == 1) ? App.a12 : App.a22;
                        v = sT[e] + h * sT[e + 3] + (t6 * p0 + t7 * p1 + t8 * p2);
                    } else if (cc < 6) {
                        const double p0 = (cc == 3) ? Avp.a00 : (cc == 4) ? Avp.a10 : Avp.a20;
                        const double p1 = (cc == 3) ? Avp.a01 : (cc == 4) ? Avp.a11 : Avp.a21;
                        const double p2 = (cc == 3) ? Avp.a02 : (cc == 4) ? Avp.a12 : Avp.a22;
                        v = sT[e] + (t6 * p0 + t7 * p1 + t8 * p2);
                    } else {
                        const double p0 = (cc == 6) ? dRk.a00 : (cc == 7) ? dRk.a01 : dRk.a02;
                        const double p1 = (cc == 6) ? dRk.a10 : (cc == 7) ? dRk.a11 : dRk.a12;
                        const double p2 = (cc == 6) ? dRk.a20 : (cc == 7) ? dRk.a21 : dRk.a22;
                        v = t6 * p0 + t7 * p1 + t8 * p2;
                    }
                    // measurement noise: Ca = [R h^2/2; R h; 0], Bg = [0; 0; Jr h]
                    const int rb = r / 3, cb = cc / 3, ri = r - 3 * rb, ci = cc - 3 * cb;
                    if (rb == 2 && cb == 2) {
                        const double j = (ri == 0) ? ((ci == 0) ? JJt.a00 : (ci == 1) ? JJt.a01 : JJt.a02)
                                         : (ri == 1) ? ((ci == 0) ? JJt.a10 : (ci == 1) ? JJt.a11 : JJt.a12)
                                                     : ((ci == 0) ? JJt.a20 : (ci == 1) ? JJt.a21 : JJt.a22);
                        v += gyr_cov * h2 * j;
                    } else if (rb < 2 && cb < 2) {
                        const double q = (ri == 0) ? ((ci == 0) ? RRt.a00 : (ci == 1) ? RRt.a01 : RRt.a02)
                                         : (ri == 1) ? ((ci == 0) ? RRt.a10 : (ci == 1) ? RRt.a11 : RRt.a12)
                                                     : ((ci == 0) ? RRt.a20 : (ci == 1) ? RRt.a21 : RRt.a22);
                        const double fr = (rb == 0) ? 0.5 * h2 : h;
                        const double fc = (cb == 0) ? 0.5 * h2 : h;
                        v += acc_cov * (fr * fc) * q;
                    }
                    sS[e] = v;
                }
                // bias Jacobians; every line uses the pre-update values of the ones below it (:98-102)
                const M3 M2 = M1 * JRg;  // R a^ J_R_bg
#define JUPD(E, F)                                                   \
    if ((E) % G == gl) {                                             \
        const int s_ = (E) / G;                                      \
        jpa[s_] += jva[s_] * h - 0.5 * R.F * h2;                     \
        jpg[s_] += jvg[s_] * h - 0.5 * M2.F * h2;                    \
        jva[s_] -= R.F * h;                                          \
        jvg[s_] -= M2.F * h;                                         \
    }
                FOR9(JUPD)
#undef JUPD
                JRg = mul_tn(dRk, JRg) - Jr * h;
                // deltas (:106-110)
                const V3 Ra = R * a;
                dP = dP + dV * h + Ra * (0.5 * h2);
                dV = dV + Ra * h;
                Q4 q = q_from_matrix(R * dRk);  // normalizeRotationM (IMUPreintegrator.h:156-174)
                if (q.w < 0) q = Q4{-q.w, -q.x, -q.y, -q.z};
                R = q_to_matrix(q_normalized(q));
                T += h;
            }
            __syncwarp();
        }
    }

    if (valid) {
        double* o = out + (size_t)pair * 142;
        if (gl == 0) {
            st3(o + 0, dP);
            st3(o + 3, dV);
            stm3(o + 6, R);
            stm3(o + 51, JRg);
            o[141] = T;
        }
#define JST(E, F)                    \
    if ((E) % G == gl) {             \
        o[15 + (E)] = jpg[(E) / G];  \
        o[24 + (E)] = jpa[(E) / G];  \
        o[33 + (E)] = jvg[(E) / G];  \
        o[42 + (E)] = jva[(E) / G];  \
    }
        FOR9(JST)
#undef JST
        for (int e = gl; e < 81; e += G) o[60 + e] = sS[e];
    }
}


// ================================================================================================
// K1 as a scan.  The update() recurrence (IMUPreintegrator.cpp:63-112) is a chain of affine maps, and every quantity it
// produces has a closed form in terms of PREFIX PRODUCTS of the per-sample rotations and PREFIX / SUFFIX SUMS of
// rotated vectors, so one warp per key-frame pair processes 32 samples at a time with no per-sample dependence:
//   R_j     = dR_0 ... dR_{j-1}                      (quaternion product scan; R_j = delta_R BEFORE sample j)
//   v_j     = R_j a_j h_j,  weight_j = T_end - T_{j+1} + h_j / 2          (T = prefix sum of h)
//   delta_V = sum v_j,      delta_P = sum v_j weight_j                     (second-order integrator unrolled)
//   J_R_bg,j = -R_j^T C_j,  C_j = sum_{i<j} R_{i+1} Jr_i h_i               (from J <- dR^T J - Jr h)
//   J_V_ba = -sum R_j h_j,  J_P_ba = -sum R_j h_j weight_j;  J_V_bg, J_P_bg the same with G_j = R_j a_j^ J_R_bg,j
//   Sigma   = sum_j Phi_j Q_j Phi_j^T,  Phi_j = A_{end-1} ... A_{j+1} = [[I, t I, X],[0, I, Y],[0, 0, Z]] with
//             t = T_end - T_{j+1}, Z = R_end^T R_{j+1}, Y = -hat(sum_{m>j} v_m) R_{j+1}, X = -hat(sum_{m>j} v_m weight_m) R_{j+1}
//             (R a^ R^T = hat(R a) turns the products of the structured A into suffix sums of the same v, v weight)
// Pairs with more than 32 samples run tile after tile; the carry into the next tile is the same algebra with the
// tile's totals (Sigma <- Phi_tile Sigma Phi_tile^T + tile sum, entry-parallel over the lanes).  Sums are formed in
// tree order instead of sample order: results agree with the sequential recurrence to round-off (tests: <= 1e-12
// absolute on the deltas, 1e-10 relative on Jacobians and covariance, against the compiled reference).
// ================================================================================================
namespace {

constexpr int kScanWarps = 4;                         // warps per CTA (static shared memory stays under 48 KB at 4 pairs per warp)
constexpr int kScanSmemPerPair = 81 + 81 + 81 + 36;   // Sigma | Phi_tile | tmp | bias-Jacobian accumulators

// shuffles / scans / sums inside a group of G lanes (G = 8, 16 or 32; groups are aligned segments of the warp)
template <int G> VD double gshfl_up(double v, int off) { return __shfl_up_sync(0xffffffffu, v, off, G); }
template <int G> VD double gshfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src, G); }
template <int G> VD Q4 gshfl_up(Q4 q, int off) { return Q4{gshfl_up<G>(q.w, off), gshfl_up<G>(q.x, off), gshfl_up<G>(q.y, off), gshfl_up<G>(q.z, off)}; }
template <int G> VD Q4 gshfl(Q4 q, int src) { return Q4{gshfl<G>(q.w, src), gshfl<G>(q.x, src), gshfl<G>(q.y, src), gshfl<G>(q.z, src)}; }
template <int G> VD V3 gshfl(V3 v, int src) { return V3{gshfl<G>(v.x, src), gshfl<G>(v.y, src), gshfl<G>(v.z, src)}; }
template <int G> VD M3 gshfl(const M3& m, int src) {
    return M3{gshfl<G>(m.a00, src), gshfl<G>(m.a01, src), gshfl<G>(m.a02, src), gshfl<G>(m.a10, src), gshfl<G>(m.a11, src),
              gshfl<G>(m.a12, src), gshfl<G>(m.a20, src), gshfl<G>(m.a21, src), gshfl<G>(m.a22, src)};
}
template <int G> VD double gscan(double v, int gl) {  // inclusive prefix sum over the group
#pragma unroll
    for (int off = 1; off < G; off <<= 1) {
        const double t = gshfl_up<G>(v, off);
        if (gl >= off) v += t;
    }
    return v;
}
template <int G> VD V3 gscan(V3 v, int gl) { return V3{gscan<G>(v.x, gl), gscan<G>(v.y, gl), gscan<G>(v.z, gl)}; }
template <int G> VD M3 gscan(const M3& m, int gl) {
    return M3{gscan<G>(m.a00, gl), gscan<G>(m.a01, gl), gscan<G>(m.a02, gl), gscan<G>(m.a10, gl), gscan<G>(m.a11, gl),
              gscan<G>(m.a12, gl), gscan<G>(m.a20, gl), gscan<G>(m.a21, gl), gscan<G>(m.a22, gl)};
}
template <int G> VD double gsum(double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int G> VD M3 gsum(const M3& m) {
    return M3{gsum<G>(m.a00), gsum<G>(m.a01), gsum<G>(m.a02), gsum<G>(m.a10), gsum<G>(m.a11), gsum<G>(m.a12),
              gsum<G>(m.a20), gsum<G>(m.a21), gsum<G>(m.a22)};
}
// a * b^T
VD M3 mul_nt(const M3& a, const M3& b) { return a * transpose(b); }
// the first lane of the group adds the 3x3 block m (already summed over the group) at block (bi, bj) of the 9x9
// accumulator, and its transpose at (bj, bi) when the block is off the diagonal
VD void add_block(double* S, int bi, int bj, const M3& m, bool writer) {
    if (!writer) return;
    const double v[9] = {m.a00, m.a01, m.a02, m.a10, m.a11, m.a12, m.a20, m.a21, m.a22};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            S[9 * (3 * bi + r) + 3 * bj + c] += v[3 * r + c];
            if (bi != bj) S[9 * (3 * bj + c) + 3 * bi + r] += v[3 * r + c];
        }
}

}  // namespace

// G lanes per key-frame pair (32 / G pairs per warp), tiles of G samples: G = 8 suits the ~40 samples of a 0.2 s
// key-frame interval at 200 Hz (5 tiles, no idle lanes), G = 32 long intervals.
template <int G>
__global__ void __launch_bounds__(32 * kScanWarps)
preint_scan_kernel(int n_pairs, const int* __restrict__ sample_begin, const double* __restrict__ gyro,
                   const double* __restrict__ acc, const double* __restrict__ dt, const double* __restrict__ bg,
                   const double* __restrict__ ba, double* __restrict__ out, double gyr_cov, double acc_cov) {
    constexpr int PPW = 32 / G;
    __shared__ double smem[kScanWarps * PPW * kScanSmemPerPair];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane % G, gi = lane / G;
    const int pair = (blockIdx.x * kScanWarps + warp) * PPW + gi;
    const bool valid = pair < n_pairs;
    double* Sig = smem + (warp * PPW + gi) * kScanSmemPerPair;  // 9x9 covariance, order [P, V, Phi]
    double* Phi = Sig + 81;
    double* Tmp = Phi + 81;
    double* Jac = Tmp + 81;                        // J_P_bg | J_P_ba | J_V_bg | J_V_ba (9 each)
    const int s0 = valid ? sample_begin[pair] : 0;
    const int cnt = valid ? sample_begin[pair + 1] - s0 : 0;
    int maxcnt = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxcnt = max(maxcnt, __shfl_xor_sync(0xffffffffu, maxcnt, o));
    const V3 bgv = valid ? ld3(bg + 3 * (size_t)pair) : v3(0, 0, 0);
    const V3 bav = valid ? ld3(ba + 3 * (size_t)pair) : v3(0, 0, 0);
    for (int e = gl; e < 81; e += G) Sig[e] = 0.0;
    for (int e = gl; e < 36; e += G) Jac[e] = 0.0;
    __syncwarp();
    // carry (replicated in every lane of the group)
    Q4 qc = Q4{1.0, 0.0, 0.0, 0.0};
    V3 dP = v3(0, 0, 0), dV = v3(0, 0, 0);
    M3 Cc = m3_zero();
    double Tc = 0.0;

    for (int base = 0; base < maxcnt; base += G) {   // (uniform over the warp: the shuffles below involve every lane)
        const bool grp = base < cnt;                  // this group still has samples
        const bool act = base + gl < cnt;
        double h = 0.0;
        V3 a = v3(0, 0, 0);
        Q4 dq = Q4{1.0, 0.0, 0.0, 0.0};
        M3 Jr = m3_identity();
        if (act) {
            const size_t s = (size_t)(s0 + base + gl);
            h = dt[s];
            const V3 w = ld3(gyro + 3 * s) - bgv;  // omega = gyro - bg   (KeyFrame.cpp:218,240)
            a = ld3(acc + 3 * s) - bav;
            dq = so3_exp(w * h);                   // Expmap (IMUPreintegrator.h:94-97)
            Jr = jacobian_r(w * h);
        }
        // ---- rotations: inclusive product scan, R_j = before sample j, R_{j+1} = after it ----
        Q4 qi = dq;
#pragma unroll
        for (int off = 1; off < G; off <<= 1) {
            const Q4 t = gshfl_up<G>(qi, off);
            if (gl >= off) qi = q_mul(t, qi);
        }
        const Q4 q1 = q_normalized(q_mul(qc, qi));  // normalizeRotationM after every product (IMUPreintegrator.h:156-174)
        Q4 q0 = gshfl_up<G>(q1, 1);
        if (gl == 0) q0 = qc;
        const Q4 qend = gshfl<G>(q1, G - 1);
        const M3 R0 = q_to_matrix(q0), R1 = q_to_matrix(q1), Rend = q_to_matrix(qend), Rstart = q_to_matrix(qc);
        // ---- prefix sums of h, v = R a h, M = R_{j+1} Jr h ----
        const V3 v = (R0 * a) * h;
        const M3 Mx = (R1 * Jr) * h;
        const double Ti = gscan<G>(h, gl);
        const V3 Vi = gscan<G>(v, gl);
        const M3 Ci = gscan<G>(Mx, gl);
        const double Tt = gshfl<G>(Ti, G - 1);
        const V3 Vt = gshfl<G>(Vi, G - 1);
        const M3 Ct = gshfl<G>(Ci, G - 1);
        const double wgt = Tt - Ti + 0.5 * h;
        const V3 wv = v * wgt;
        const V3 Wi = gscan<G>(wv, gl);
        const V3 Wt = gshfl<G>(Wi, G - 1);
        // ---- bias Jacobians (IMUPreintegrator.cpp:98-102) ----
        const M3 JRg = -(mul_tn(R0, Cc + (Ci - Mx)));          // J_R_bg before sample j
        const M3 Gm = (R0 * hat(a)) * JRg;                     // delta_R a^ J_R_bg
        {
            const M3 Rh = R0 * h, Gh = Gm * h;
            const M3 sRh = gsum<G>(Rh), sRhw = gsum<G>(Rh * wgt), sGh = gsum<G>(Gh), sGhw = gsum<G>(Gh * wgt);
            if (gl == 0 && grp) {
                const double rh[9] = {sRh.a00, sRh.a01, sRh.a02, sRh.a10, sRh.a11, sRh.a12, sRh.a20, sRh.a21, sRh.a22};
                const double rw[9] = {sRhw.a00, sRhw.a01, sRhw.a02, sRhw.a10, sRhw.a11, sRhw.a12, sRhw.a20, sRhw.a21, sRhw.a22};
                const double gh[9] = {sGh.a00, sGh.a01, sGh.a02, sGh.a10, sGh.a11, sGh.a12, sGh.a20, sGh.a21, sGh.a22};
                const double gw[9] = {sGhw.a00, sGhw.a01, sGhw.a02, sGhw.a10, sGhw.a11, sGhw.a12, sGhw.a20, sGhw.a21, sGhw.a22};
#pragma unroll
                for (int e = 0; e < 9; ++e) {
                    Jac[e] += Jac[18 + e] * Tt - gw[e];       // J_P_bg += J_V_bg T_tile - sum G h weight
                    Jac[9 + e] += Jac[27 + e] * Tt - rw[e];   // J_P_ba += J_V_ba T_tile - sum R h weight
                    Jac[18 + e] -= gh[e];                     // J_V_bg
                    Jac[27 + e] -= rh[e];                     // J_V_ba
                }
            }
        }
        // ---- covariance: carry through the tile, then the tile's own sum ----
        if (base > 0) {
            const M3 Xt = -(hat(Wt) * Rstart), Yt = -(hat(Vt) * Rstart), Zt = mul_tn(Rend, Rstart);
            if (gl == 0 && grp) {
                for (int e = 0; e < 81; ++e) Phi[e] = 0.0;
                for (int i = 0; i < 6; ++i) Phi[10 * i] = 1.0;
                for (int i = 0; i < 3; ++i) Phi[9 * i + 3 + i] = Tt;
                const double x[9] = {Xt.a00, Xt.a01, Xt.a02, Xt.a10, Xt.a11, Xt.a12, Xt.a20, Xt.a21, Xt.a22};
                const double y[9] = {Yt.a00, Yt.a01, Yt.a02, Yt.a10, Yt.a11, Yt.a12, Yt.a20, Yt.a21, Yt.a22};
                const double z[9] = {Zt.a00, Zt.a01, Zt.a02, Zt.a10, Zt.a11, Zt.a12, Zt.a20, Zt.a21, Zt.a22};
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) {
                        Phi[9 * r + 6 + c] = x[3 * r + c];
                        Phi[9 * (3 + r) + 6 + c] = y[3 * r + c];
                        Phi[9 * (6 + r) + 6 + c] = z[3 * r + c];
                    }
            }
            __syncwarp();
            if (grp) {
                // Tmp = Phi Sigma with Phi = [[I, t I, X],[0, I, Y],[0, 0, Z]]: identity / t I blocks applied directly
                for (int e = gl; e < 81; e += G) {
                    const int r = e / 9, c = e - 9 * r;
                    double t = (r < 6) ? Sig[e] : 0.0;
                    if (r < 3) t = fma(Tt, Sig[e + 27], t);
#pragma unroll
                    for (int k = 6; k < 9; ++k) t = fma(Phi[9 * r + k], Sig[9 * k + c], t);
                    Tmp[e] = t;
                }
            }
            __syncwarp();
            if (grp) {
                for (int e = gl; e < 81; e += G) {  // Sigma = Tmp Phi^T
                    const int r = e / 9, c = e - 9 * r;
                    double t = (c < 6) ? Tmp[e] : 0.0;
                    if (c < 3) t = fma(Tt, Tmp[e + 3], t);
#pragma unroll
                    for (int k = 6; k < 9; ++k) t = fma(Tmp[9 * r + k], Phi[9 * c + k], t);
                    Sig[e] = t;
                }
            }
            __syncwarp();
        }
        {
            const double tj = Tt - Ti;                             // time from the end of sample j to the end of the tile
            const M3 X = -(hat(Wt - Wi) * R1), Y = -(hat(Vt - Vi) * R1), Z = mul_tn(Rend, R1);
            const M3 E = mul_nt(R0, R0);                           // R R^T as the reference forms it (Ca Sigma_a Ca^T)
            const M3 XJ = X * Jr, YJ = Y * Jr, ZJ = Z * Jr;        // Bg = [0; 0; Jr h]
            const double g = gyr_cov * h * h;
            const double cp = 0.5 * h * h + tj * h;                // Phi Ca = [R (h^2/2 + t h); R h; 0]
            const bool wr = gl == 0 && grp;
            add_block(Sig, 0, 0, gsum<G>(E * (acc_cov * cp * cp) + mul_nt(XJ, XJ) * g), wr);
            add_block(Sig, 0, 1, gsum<G>(E * (acc_cov * cp * h) + mul_nt(XJ, YJ) * g), wr);
            add_block(Sig, 1, 1, gsum<G>(E * (acc_cov * h * h) + mul_nt(YJ, YJ) * g), wr);
            add_block(Sig, 0, 2, gsum<G>(mul_nt(XJ, ZJ) * g), wr);
            add_block(Sig, 1, 2, gsum<G>(mul_nt(YJ, ZJ) * g), wr);
            add_block(Sig, 2, 2, gsum<G>(mul_nt(ZJ, ZJ) * g), wr);
        }
        // ---- deltas (IMUPreintegrator.cpp:106-110) and the carry into the next tile ----
        if (grp) {
            dP = dP + dV * Tt + Wt;
            dV = dV + Vt;
            Cc = Cc + Ct;
            qc = qend;
            Tc += Tt;
        }
        __syncwarp();
    }
    if (!valid) return;
    double* o = out + (size_t)pair * 142;
    const M3 R = q_to_matrix(qc);
    if (gl == 0) {
        st3(o + 0, dP);
        st3(o + 3, dV);
        stm3(o + 6, R);
        stm3(o + 51, -(mul_tn(R, Cc)));
        o[141] = Tc;
    }
    for (int e = gl; e < 36; e += G) o[15 + e] = Jac[e];
    for (int e = gl; e < 81; e += G) o[60 + e] = Sig[e];
}

size_t preint_smem_bytes(int group) {
    const int ppw = 32 / group;
    const int per_pair = 81 + 81 + 22 * group;
    return sizeof(double) * (size_t)(kPreintThreads / 32) * ppw * per_pair;
}

cudaError_t launch_preint_batch(cudaStream_t stream, int n_pairs, const int* sample_begin, const double* gyro,
                                const double* acc, const double* dt, const double* bg, const double* ba,
                                double* out, double gyr_cov, double acc_cov, int group) {
    if (n_pairs <= 0) return cudaSuccess;
    if (group <= 0) {  // scan kernels: -8 / -16 / -32 = lanes per key-frame pair (0: 8)
        const int g = group == 0 ? 8 : -group;
        const int ppc = kScanWarps * (32 / g);
        const int grid = (n_pairs + ppc - 1) / ppc;
        if (g == 32) preint_scan_kernel<32><<<grid, 32 * kScanWarps, 0, stream>>>(n_pairs, sample_begin, gyro, acc, dt, bg, ba, out, gyr_cov, acc_cov);
        else if (g == 16) preint_scan_kernel<16><<<grid, 32 * kScanWarps, 0, stream>>>(n_pairs, sample_begin, gyro, acc, dt, bg, ba, out, gyr_cov, acc_cov);
        else preint_scan_kernel<8><<<grid, 32 * kScanWarps, 0, stream>>>(n_pairs, sample_begin, gyro, acc, dt, bg, ba, out, gyr_cov, acc_cov);
        return cudaGetLastError();
    }
    const int warps_per_cta = kPreintThreads / 32;
    const size_t smem = preint_smem_bytes(group);
#define LAUNCH(Gv)                                                                                          \
    {                                                                                                       \
        const int pairs_per_cta = warps_per_cta * (32 / Gv);                                                \
        const int grid = (n_pairs + pairs_per_cta - 1) / pairs_per_cta;                                     \
        cudaError_t e = opt_in_max_smem(preint_batch_kernel<Gv>);       \
        if (e != cudaSuccess) return e;                                                                     \
        preint_batch_kernel<Gv><<<grid, kPreintThreads, smem, stream>>>(n_pairs, sample_begin, gyro, acc,   \
                                                                         dt, bg, ba, out, gyr_cov, acc_cov); \
    }
    if (group == 32) LAUNCH(32) else if (group == 16) LAUNCH(16) else LAUNCH(8)
#undef LAUNCH
    return cudaGetLastError();
}

}  // namespace vilba
