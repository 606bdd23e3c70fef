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

size_t preint_smem_bytes(int group) {
    const int ppw = 32 / group;
    const int per_pair = 81 + 81 + 22 * group;
    return sizeof(double) * (size_t)(kPreintThreads / 32) * ppw * per_pair;
}

cudaError_t launch_preint_batch(cudaStream_t stream, int n_pairs, const int* sample_begin, const double* gyro,
                                const double* acc, const double* dt, const double* bg, const double* ba,
                                double* out, double gyr_cov, double acc_cov, int group) {
    if (n_pairs <= 0) return cudaSuccess;
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
