// lba_v2.cu -- atomic-free, run-to-run deterministic versions of the accumulation kernels.
//
//   linearize_v2 : same arithmetic as linearize_mono/imu (EdgeNavStatePVRPointXYZ / EdgeNavStatePVR /
//                  EdgeNavStateBias linearizeOplus + constructQuadraticForm, src/IMU/g2otypes.cpp:587-699,
//                  724-734,738-788; g2o/core/base_binary_edge.hpp:55-120; base_multi_edge.hpp:171-222), but the
//                  pose-block contributions of the mono edges are summed in WARP-PRIVATE shared-memory
//                  accumulators (the edges of one point hit distinct key-frames, so a warp never collides
//                  with itself), reduced per CTA in a fixed order and written as one partial per CTA;
//                  every IMU edge pair writes its own 30x30 slot.  IMU CTAs run in the same launch.
//   assemble_hpp : H_pp / b_p = fixed-order sum of the CTA partials and the IMU slots
//                  (BlockSolver::buildSystem's flush, g2o/core/block_solver.hpp:547-557); its last CTA to arrive starts
//                  the LM iteration (lambda init, bookkeeping: lm_iter_begin_cta).
//   schur_rec / schur_tile / schur_finish : the same landmark loop for windows of <= 32 key-frames as a TILE SCAN over
//                  the map points (TMA bulk copies, key-frame masks, no pair lists); options: one lane per block pair,
//                  one FP64 MMA per hit.  See the comments above each kernel.
//   schur_gather : S(a,b) = H_pp(a,b) + lambda I - sum_l W_a,l D_l^-1 W_b,l^T as a GATHER over precomputed
//                  (edge_a, edge_b) lists per key-frame block pair, one CTA per block pair, register
//                  accumulation, fixed reduction tree -- replaces 6.5 M FP64 atomics per LM trial
//                  (landmark loop of BlockSolver::solve, block_solver.hpp:381-439).
#include "lba_common.cuh"

namespace vilba {

constexpr int kAccStride = 27;

// sum over the 8 lanes of a sub-warp group (xor tree: every lane gets the same bits)
__device__ __forceinline__ double group8_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}  // 21 upper entries of the 6x6 [P,Phi] block + 6 rhs entries

size_t linearize_v2_smem_bytes(int K, int n_free, int warps) {
    return kf_smem_bytes(K) + sizeof(double) * (size_t)warps * n_free * kAccStride + 16;
}

__global__ void __launch_bounds__(32) linearize_imu_v2_kernel(const DevWindow* __restrict__ wp) {
    const DevWindow w = wp[blockIdx.y];  // one window per grid row
    if (w.lm->phase != PH_LINEARIZE) return;
    // one warp per IMU edge pair (EdgeNavStatePVR + EdgeNavStateBias)
    __shared__ double J[216];   // 9 x 24: PVR_i (9) | Bias_i (6) | PVR_j (9)
    __shared__ double Om[81];
    __shared__ double TJ[216];  // 9 x 24  (rho1 Omega) J
    __shared__ double ev[9];
    __shared__ double Oe[9];    // Omega e
    __shared__ double misc[2];  // [0] rho1
    const int lane = threadIdx.x & 31;
    const int cur = w.lm->cur;
    for (int e = blockIdx.x; e < w.NI; e += gridDim.x) {
        const int ki = w.imu_i[e], kj = w.imu_j[e];
        const double* si = (cur ? w.kf_state[1] : w.kf_state[0]) + 22 * (size_t)ki;
        const double* sj = (cur ? w.kf_state[1] : w.kf_state[0]) + 22 * (size_t)kj;
        const double* M = w.imu_preint + 142 * (size_t)e;
        for (int i = lane; i < 81; i += 32) Om[i] = w.imu_info[81 * (size_t)e + i];
        for (int i = lane; i < 216; i += 32) J[i] = 0.0;
        __syncwarp();
        if (lane == 0) {
            V3 rP, rV, rPhi;
            pvr_error(w, si, sj, M, rP, rV, rPhi);
            ev[0] = rP.x, ev[1] = rP.y, ev[2] = rP.z, ev[3] = rV.x, ev[4] = rV.y, ev[5] = rV.z;
            ev[6] = rPhi.x, ev[7] = rPhi.y, ev[8] = rPhi.z;
            const V3 Pi = ld3(si), Vi = ld3(si + 3), Pj = ld3(sj), Vj = ld3(sj + 3);
            const M3 Ri = q_to_matrix(Q4{si[6], si[7], si[8], si[9]});
            const M3 Rj = q_to_matrix(Q4{sj[6], sj[7], sj[8], sj[9]});
            const V3 dbg = ld3(si + 16);
            const V3 g = ld3(w.g);
            const double T = M[VILBA_PI_DT], T2 = T * T;
            const M3 RiT = transpose(Ri);
            const M3 JrInv = jacobian_r_inv(rPhi);
            const M3 JRg = ldm3(M + VILBA_PI_JRG);
            auto put = [&](int r0, int c0, const M3& B) {
                J[(r0 + 0) * 24 + c0 + 0] = B.a00, J[(r0 + 0) * 24 + c0 + 1] = B.a01, J[(r0 + 0) * 24 + c0 + 2] = B.a02;
                J[(r0 + 1) * 24 + c0 + 0] = B.a10, J[(r0 + 1) * 24 + c0 + 1] = B.a11, J[(r0 + 1) * 24 + c0 + 2] = B.a12;
                J[(r0 + 2) * 24 + c0 + 0] = B.a20, J[(r0 + 2) * 24 + c0 + 1] = B.a21, J[(r0 + 2) * 24 + c0 + 2] = B.a22;
            };
            put(0, 0, -RiT);
            put(0, 3, RiT * (-T));
            put(0, 6, hat(RiT * (Pj - Pi - Vi * T - (0.5 * g) * T2)));
            put(3, 3, -RiT);
            put(3, 6, hat(RiT * (Vj - Vi - g * T)));
            put(6, 6, ((-JrInv) * transpose(Rj)) * Ri);
            const M3 ExpT = q_to_matrix(so3_inverse(so3_exp(rPhi)));
            const M3 JrCorr = jacobian_r(JRg * dbg);
            put(0, 9, -ldm3(M + VILBA_PI_JPG));
            put(0, 12, -ldm3(M + VILBA_PI_JPA));
            put(3, 9, -ldm3(M + VILBA_PI_JVG));
            put(3, 12, -ldm3(M + VILBA_PI_JVA));
            put(6, 9, (((-JrInv) * ExpT) * JrCorr) * JRg);
            put(0, 15, RiT);
            put(3, 18, RiT);
            put(6, 21, JrInv);
        }
        __syncwarp();
        if (lane < 9) {  // Omega e, then chi2 = e^T Omega e and the Huber weight
            double t = 0.0;
#pragma unroll
            for (int c = 0; c < 9; ++c) t += Om[9 * lane + c] * ev[c];
            Oe[lane] = t;
        }
        __syncwarp();
        if (lane == 0) {
            double c2 = 0.0;
#pragma unroll
            for (int r = 0; r < 9; ++r) c2 += ev[r] * Oe[r];
            double rho0, rho1;
            huber(c2, w.huber_pvr, rho0, rho1);
            misc[0] = rho1;
        }
        __syncwarp();
        const double wgt = misc[0];
        for (int i = lane; i < 216; i += 32) {
            const int r = i / 24, c = i - 24 * r;
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 9; ++k) s += Om[9 * r + k] * J[24 * k + c];
            TJ[i] = wgt * s;
        }
        __syncwarp();
        // bias edge (A = -I, B = +I): weights on the diagonal
        V3 rg, ra;
        bias_error(si, sj, rg, ra);
        const double wg = w.inv_gyr_rw2 / M[VILBA_PI_DT], wa = w.inv_acc_rw2 / M[VILBA_PI_DT];
        const double c2 = rg.x * (wg * rg.x) + rg.y * (wg * rg.y) + rg.z * (wg * rg.z) + ra.x * (wa * ra.x) +
                          ra.y * (wa * ra.y) + ra.z * (wa * ra.z);
        double brho0, brho1;
        huber(c2, w.huber_bias, brho0, brho1);
        const double eb[6] = {rg.x, rg.y, rg.z, ra.x, ra.y, ra.z};
        // slot layout: 30 x 30 (local order PVR_i 9 | Bias_i 6 | PVR_j 9 | Bias_j 6) followed by 30 rhs entries
        double* slot = w.imu_slot + 930 * (size_t)e;
        for (int i = lane; i < 900; i += 32) {
            const int r = i / 30, c = i - 30 * r;
            double v = 0.0;
            if (r < 24 && c < 24) {
#pragma unroll
                for (int k = 0; k < 9; ++k) v += J[24 * k + r] * TJ[24 * k + c];
            }
            // bias-edge blocks: (Bias_i,Bias_i) += w, (Bias_j,Bias_j) += w, (Bias_i,Bias_j) and transpose -= w
            const int rb = (r >= 9 && r < 15) ? r - 9 : (r >= 24 ? r - 24 : -1);
            const int cb = (c >= 9 && c < 15) ? c - 9 : (c >= 24 ? c - 24 : -1);
            if (rb >= 0 && rb == cb) {
                const double om = brho1 * (rb < 3 ? wg : wa);
                const bool r_is_i = r < 15, c_is_i = c < 15;
                v += (r_is_i == c_is_i) ? om : -om;
            }
            slot[i] = v;
        }
        if (lane < 30) {
            const int r = lane;
            double v = 0.0;
            if (r < 24) {  // A^T * (-rho1 Omega e)
#pragma unroll
                for (int k = 0; k < 9; ++k) v -= J[24 * k + r] * (wgt * Oe[k]);
            }
            const int rb = (r >= 9 && r < 15) ? r - 9 : (r >= 24 ? r - 24 : -1);
            if (rb >= 0) {
                const double om = brho1 * (rb < 3 ? wg : wa);
                const double omega_r = -om * eb[rb];
                v += (r < 15) ? -omega_r : omega_r;  // A^T omega_r with A = -I ; B^T omega_r with B = +I
            }
            slot[900 + r] = v;
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(kPointThreads, 2) linearize_v2_kernel(const DevWindow* __restrict__ wp) {
    const DevWindow w = wp[blockIdx.y];  // one window per grid row
    if (w.lm->phase != PH_LINEARIZE) return;
    const int point_ctas = gridDim.x;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int cur = w.lm->cur;

    // ======================= mono edges: one warp per map point =======================
    const KfSmem ks = kf_smem_carve(smem, w.K);
    const int warps_per_cta = blockDim.x >> 5;
    double* acc_base = reinterpret_cast<double*>(reinterpret_cast<char*>(smem) + ((kf_smem_bytes(w.K) + 15) / 16) * 16);
    const int nf = w.n_free;
    double* acc = acc_base + (size_t)warp * nf * kAccStride;
    for (int i = threadIdx.x; i < warps_per_cta * nf * kAccStride; i += blockDim.x) acc_base[i] = 0.0;
    kf_stage<false>(w, ks, cur);
    __syncthreads();
    const int gwarp = blockIdx.x * warps_per_cta + warp;
    const int nwarps = point_ctas * warps_per_cta;
    const double* pts = (cur ? w.pts[1] : w.pts[0]);
    const M3 Rcb = ldm3(w.Rcb);
    double maxd = 0.0;

    // 8 lanes per map point, 4 points per warp: a point has ~8 observations, so the lanes stay busy.  The pose-block
    // sums go to the warp-private accumulator in four phases (one point at a time: the key-frames of ONE point
    // are distinct, those of neighbouring points are not), which keeps the summation order fixed.
    const int gl = lane & 7, grp = lane >> 3;
    for (int base = gwarp * 4; base < w.P; base += nwarps * 4) {
        const int p = base + grp;
        const bool valid = p < w.P;
        int e0i = 0, e1i = 0;
        V3 Pw = v3(0, 0, 0);
        if (valid) {
            e0i = w.pt_obs_begin[p], e1i = w.pt_obs_begin[p + 1];
            Pw = ld3(pts + 3 * (size_t)p);
        }
        const int rounds = __reduce_max_sync(0xffffffffu, (e1i - e0i + 7) >> 3);
        double hxx = 0, hxy = 0, hxz = 0, hyy = 0, hyz = 0, hzz = 0, bx = 0, by = 0, bz = 0;
        for (int round = 0; round < rounds; ++round) {
            const int e = e0i + 8 * round + gl;
            const bool have = e < e1i;
            bool to_pose = false;
            int blk = -1;
            // contributions of this edge to its key-frame's [P,Phi] block: 21 upper entries + 6 rhs entries
            double ctr[kAccStride];
#pragma unroll
            for (int i = 0; i < kAccStride; ++i) ctr[i] = 0.0;
            if (have) {
                const MonoObs o = load_obs(w.obs, e);
                double* Wp = w.W + 18 * (size_t)e;
                blk = ks.blk[o.kf];
                bool wrote_w = false;
                if (!o.culled) {
                    const double* cam = ks.cam + 12 * o.kf;
                    double r0, r1;
                    V3 Paux, Pc;
                    mono_error(w, cam, Pw, o, r0, r1, Paux, Pc);
                    const double is2 = (double)o.is2;
                    double wgt = is2;
                    if (o.robust) {
                        double rho0, rho1;
                        huber(r0 * (is2 * r0) + r1 * (is2 * r1), w.huber_mono, rho0, rho1);
                        wgt = rho1 * is2;
                    }
                    const M3 Rcw = ldm3(cam);
                    const double iz = 1.0 / Pc.z;  // see mono_error: one reciprocal for the six divisions by z
                    const double ja = w.fx * iz, jb = (-(Pc.x * iz) * w.fx) * iz;
                    const double jc = w.fy * iz, jd = (-(Pc.y * iz) * w.fy) * iz;
                    // J_l (2x3, w.r.t. the point) and F (2x3, w.r.t. dPhi); the pose Jacobian is J_p = [-J_l | F]
                    const double l00 = -(ja * Rcw.a00 + jb * Rcw.a20), l01 = -(ja * Rcw.a01 + jb * Rcw.a21),
                                 l02 = -(ja * Rcw.a02 + jb * Rcw.a22);
                    const double l10 = -(jc * Rcw.a10 + jd * Rcw.a20), l11 = -(jc * Rcw.a11 + jd * Rcw.a21),
                                 l12 = -(jc * Rcw.a12 + jd * Rcw.a22);
                    const M3 HR = hat(Paux) * Rcb;
                    const double f00 = -(ja * HR.a00 + jb * HR.a20), f01 = -(ja * HR.a01 + jb * HR.a21),
                                 f02 = -(ja * HR.a02 + jb * HR.a22);
                    const double f10 = -(jc * HR.a10 + jd * HR.a20), f11 = -(jc * HR.a11 + jd * HR.a21),
                                 f12 = -(jc * HR.a12 + jd * HR.a22);
                    // Because J_p = [-J_l | F], every block of the edge's quadratic form is one of three products:
                    //   A = J_l^T w J_l (H_ll part, = top-left of H_pp, = -rows 0..2 of H_pl),
                    //   B = F^T w J_l   (rows 3..5 of H_pl, = -top-right of H_pp transposed),  C = F^T w F.
                    const double m00 = wgt * l00, m01 = wgt * l01, m02 = wgt * l02;  // w J_l
                    const double m10 = wgt * l10, m11 = wgt * l11, m12 = wgt * l12;
                    const double axx = l00 * m00 + l10 * m10, axy = l00 * m01 + l10 * m11, axz = l00 * m02 + l10 * m12;
                    const double ayy = l01 * m01 + l11 * m11, ayz = l01 * m02 + l11 * m12, azz = l02 * m02 + l12 * m12;
                    hxx += axx, hxy += axy, hxz += axz, hyy += ayy, hyz += ayz, hzz += azz;
                    const double wr0 = -wgt * r0, wr1 = -wgt * r1;
                    const double ebx = l00 * wr0 + l10 * wr1, eby = l01 * wr0 + l11 * wr1, ebz = l02 * wr0 + l12 * wr1;
                    bx += ebx, by += eby, bz += ebz;
                    if (blk >= 0) {
                        const double b00 = f00 * m00 + f10 * m10, b01 = f00 * m01 + f10 * m11, b02 = f00 * m02 + f10 * m12;
                        const double b10 = f01 * m00 + f11 * m10, b11 = f01 * m01 + f11 * m11, b12 = f01 * m02 + f11 * m12;
                        const double b20 = f02 * m00 + f12 * m10, b21 = f02 * m01 + f12 * m11, b22 = f02 * m02 + f12 * m12;
                        const double g00 = wgt * f00, g01 = wgt * f01, g02 = wgt * f02;  // w F
                        const double g10 = wgt * f10, g11 = wgt * f11, g12 = wgt * f12;
                        // H_pl block (6x3): rows [P] = -A, rows [Phi] = B
                        Wp[0] = -axx, Wp[1] = -axy, Wp[2] = -axz;
                        Wp[3] = -axy, Wp[4] = -ayy, Wp[5] = -ayz;
                        Wp[6] = -axz, Wp[7] = -ayz, Wp[8] = -azz;
                        Wp[9] = b00, Wp[10] = b01, Wp[11] = b02;
                        Wp[12] = b10, Wp[13] = b11, Wp[14] = b12;
                        Wp[15] = b20, Wp[16] = b21, Wp[17] = b22;
                        // upper triangle of J_p^T w J_p, row-major: [A | -B^T ; . | C]
                        ctr[0] = axx, ctr[1] = axy, ctr[2] = axz, ctr[3] = -b00, ctr[4] = -b10, ctr[5] = -b20;
                        ctr[6] = ayy, ctr[7] = ayz, ctr[8] = -b01, ctr[9] = -b11, ctr[10] = -b21;
                        ctr[11] = azz, ctr[12] = -b02, ctr[13] = -b12, ctr[14] = -b22;
                        ctr[15] = f00 * g00 + f10 * g10, ctr[16] = f00 * g01 + f10 * g11, ctr[17] = f00 * g02 + f10 * g12;
                        ctr[18] = f01 * g01 + f11 * g11, ctr[19] = f01 * g02 + f11 * g12;
                        ctr[20] = f02 * g02 + f12 * g12;
                        // rhs: J_p^T (-w r) = [-b_l part | F^T (-w r)]
                        ctr[21] = -ebx, ctr[22] = -eby, ctr[23] = -ebz;
                        ctr[24] = f00 * wr0 + f10 * wr1, ctr[25] = f01 * wr0 + f11 * wr1, ctr[26] = f02 * wr0 + f12 * wr1;
                        wrote_w = true;
                        to_pose = true;
                    }
                }
                if (!wrote_w) {
#pragma unroll
                    for (int i = 0; i < 18; ++i) Wp[i] = 0.0;
                }
            }
#pragma unroll 1
            for (int ph = 0; ph < 4; ++ph) {
                if (grp == ph && to_pose) {
                    double* a = acc + (size_t)blk * kAccStride;
                    // read - add - write in chunks of 9 so that the shared-memory round trips overlap
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        double t[9];
#pragma unroll
                        for (int i = 0; i < 9; ++i) t[i] = a[9 * ch + i];
#pragma unroll
                        for (int i = 0; i < 9; ++i) t[i] += ctr[9 * ch + i];
#pragma unroll
                        for (int i = 0; i < 9; ++i) a[9 * ch + i] = t[i];
                    }
                }
                __syncwarp();
            }
        }
        hxx = group8_sum(hxx), hxy = group8_sum(hxy), hxz = group8_sum(hxz);
        hyy = group8_sum(hyy), hyz = group8_sum(hyz), hzz = group8_sum(hzz);
        bx = group8_sum(bx), by = group8_sum(by), bz = group8_sum(bz);
        if (gl == 0 && valid) {
            double* H = w.Hll + 6 * (size_t)p;
            H[0] = hxx, H[1] = hxy, H[2] = hxz, H[3] = hyy, H[4] = hyz, H[5] = hzz;
            st3(w.bl + 3 * (size_t)p, v3(bx, by, bz));
            maxd = fmax(maxd, fmax(fabs(hxx), fmax(fabs(hyy), fabs(hzz))));
        }
    }
    maxd = warp_max(maxd);
    if (lane == 0 && maxd > 0.0) atomic_max_nonneg(&w.lm->maxdiag_bits, maxd);  // max is order-independent
    __syncthreads();
    // CTA partial = fixed-order sum over its warps
    double* part = w.lin_partial + (size_t)blockIdx.x * nf * kAccStride;
    for (int i = threadIdx.x; i < nf * kAccStride; i += blockDim.x) {
        double s = 0.0;
        for (int ww = 0; ww < warps_per_cta; ++ww) s += acc_base[(size_t)ww * nf * kAccStride + i];
        part[i] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// assemble: Hpp(r,c), r <= c, and bp from the CTA partials and the IMU slots (fixed order)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int pose6_index(int o) {  // offset inside a 15-block -> index in the [P,Phi] 6-block
    return (o < 3) ? o : ((o >= 6 && o < 9) ? o - 3 : -1);
}

// fixed-order reduction of the CTA partials: one warp per entry (lane-strided partial sums, xor tree)
__global__ void __launch_bounds__(256) reduce_partials_kernel(const DevWindow* __restrict__ wp, int point_ctas) {
    const DevWindow w = wp[blockIdx.y];  // one window per grid row
    if (w.lm->phase != PH_LINEARIZE) return;
    const int entries = w.n_free * kAccStride;
    const int lane = threadIdx.x & 31;
    const int nw = gridDim.x * (blockDim.x >> 5);
    for (int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); gw < entries; gw += nw) {
        double s = 0.0;
        for (int cta = lane; cta < point_ctas; cta += 32) s += w.lin_partial[(size_t)cta * entries + gw];
        s = warp_sum(s);
        if (lane == 0) w.mono_sum[gw] = s;
    }
}

__global__ void __launch_bounds__(256) assemble_hpp_kernel(const DevWindow* __restrict__ wp) {
    const DevWindow w = wp[blockIdx.y];  // one window per grid row
    if (w.lm->phase != PH_LINEARIZE) return;
    const int n = w.n;
    const size_t total = (size_t)n * n + n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const bool is_rhs = i >= (size_t)n * n;
        int r, c;
        if (is_rhs) {
            r = (int)(i - (size_t)n * n);
            c = r;
        } else {
            r = (int)(i / n);
            c = (int)(i - (size_t)r * n);
            if (r > c) {
                w.Hpp_w[i] = 0.0;
                continue;
            }
        }
        const int a = r / 15, b = c / 15, oa = r - 15 * a, ob = c - 15 * b;
        double v = 0.0;
        // mono partials
        const int pa = pose6_index(oa), pb = pose6_index(ob);
        if (a == b && pa >= 0 && (is_rhs || pb >= 0)) {
            int idx;
            if (is_rhs)
                idx = 21 + pa;
            else  // upper index of (pa, pb), pa <= pb: rows of length 6,5,4,...
                idx = pa * 6 - (pa * (pa - 1)) / 2 + (pb - pa);
            v += w.mono_sum[(size_t)a * kAccStride + idx];
        }
        // IMU slots: the edge where block a is the "i" key-frame, then the one where it is the "j" key-frame
        // (added by one rank only when the window is sharded)
        const int ea[2] = {w.shard_owner ? w.blk_edge_i[a] : -1, w.shard_owner ? w.blk_edge_j[a] : -1};
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int e = ea[s];
            if (e < 0) continue;
            const int bi = w.kf_block[w.imu_i[e]], bj = w.kf_block[w.imu_j[e]];
            const int la = (a == bi) ? oa : 15 + oa;
            int lb;
            if (b == bi)
                lb = ob;
            else if (b == bj)
                lb = 15 + ob;
            else
                continue;
            if (a == b && ((a == bi) != (s == 0))) continue;  // safety: slot s must match the role of a
            v += is_rhs ? w.imu_slot[930 * (size_t)e + 900 + la] : w.imu_slot[930 * (size_t)e + 30 * la + lb];
        }
        if (is_rhs)
            w.bp_w[r] = v;
        else
            w.Hpp_w[i] = v;
    }
    if (w.sharded) return;  // diag H crosses the ranks first: lm_iter_begin_kernel behind the reduction
    // the last CTA to arrive starts the LM iteration (H_pp is complete then): no kernel of its own for a few loads
    __shared__ bool is_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = atomicAdd(w.chi_counter + 1, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        if (threadIdx.x == 0) w.chi_counter[1] = 0u;
        lm_iter_begin_cta(w, w.lm);
    }
}

// ------------------------------------------------------------------------------------------------
// Schur gather: one CTA per key-frame block pair (a <= b)
// ------------------------------------------------------------------------------------------------
constexpr int kSchurThreads = 1024;  // 32 warps x 5 (edge_a, edge_b) pairs per pass

// per trial: Y_e = W_e D_l^-1 (BDinv, block_solver.hpp:407) and Wc_e = W_e (D_l^-1 b_l) for every mono edge
__global__ void __launch_bounds__(256) schur_prep_kernel(const DevWindow* __restrict__ wp) {
    const DevWindow w = wp[blockIdx.y];  // one window per grid row
    if (w.lm->phase != PH_TRIAL) return;
    const double lambda = w.lm->lambda;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < w.E; e += gridDim.x * blockDim.x) {
        const int p = w.edge_pt[e];
        const double* H = w.Hll + 6 * (size_t)p;
        bool ok;
        const S3 Dinv = s3_inverse(S3{H[0] + lambda, H[1], H[2], H[3] + lambda, H[4], H[5] + lambda}, ok);
        const V3 db = s3_mul(Dinv, ld3(w.bl + 3 * (size_t)p));
        const double* Wp = w.W + 18 * (size_t)e;
        double* Yp = w.Y + 24 * (size_t)e;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            const V3 wr = v3(Wp[3 * r], Wp[3 * r + 1], Wp[3 * r + 2]);
            const V3 yv = s3_mul(Dinv, wr);
            Yp[3 * r] = yv.x, Yp[3 * r + 1] = yv.y, Yp[3 * r + 2] = yv.z;
            Yp[18 + r] = dot(wr, db);
        }
    }
}

// Lane mapping: a warp handles 5 list entries per pass, 6 lanes each; lane (slot, r) owns row r of the
// 6x6 product Y_a W_b^T of its entry, so the 6 lanes of an entry read Y_a (144 B) and W_b (144 B) as
// coalesced, L1-broadcast lines instead of 42 scattered 8-byte loads per thread.
__global__ void __launch_bounds__(kSchurThreads) schur_gather_kernel(const DevWindow* __restrict__ wp) {
    const DevWindow w = wp[blockIdx.y];  // one window per grid row
    if (w.lm->phase != PH_TRIAL) return;
    __shared__ double red[kSchurThreads / 32][42];
    __shared__ double blockacc[42];
    for (int pair = blockIdx.x; pair < w.n_pairs; pair += gridDim.x) {
    const int a = w.pair_a[pair], b = w.pair_b[pair];
    const int t0 = w.pair_begin[pair], t1 = w.pair_begin[pair + 1];
    const double lambda = w.lm->lambda;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool diag = (a == b);
    const int slot = lane / 6, r = lane - 6 * slot;  // lanes 30, 31 idle
    const bool live = lane < 30;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    double rb = 0.0;
    constexpr int kPerPass = (kSchurThreads / 32) * 5;
    if (live) {
        // two list entries in flight per lane group: the index -> row loads of both overlap
        int t = t0 + warp * 5 + slot;
        for (; t + kPerPass < t1; t += 2 * kPerPass) {
            const int ea0 = w.pair_ea[t], eb0 = w.pair_eb[t];
            const int ea1 = w.pair_ea[t + kPerPass], eb1 = w.pair_eb[t + kPerPass];
            const double* Ya0 = w.Y + 24 * (size_t)ea0;
            const double* Wb0 = w.W + 18 * (size_t)eb0;
            const double* Ya1 = w.Y + 24 * (size_t)ea1;
            const double* Wb1 = w.W + 18 * (size_t)eb1;
            const double y00 = Ya0[3 * r], y01 = Ya0[3 * r + 1], y02 = Ya0[3 * r + 2];
            const double y10 = Ya1[3 * r], y11 = Ya1[3 * r + 1], y12 = Ya1[3 * r + 2];
            double b0[18], b1[18];
#pragma unroll
            for (int i = 0; i < 18; ++i) b0[i] = Wb0[i], b1[i] = Wb1[i];
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                acc[c] += y00 * b0[3 * c] + y01 * b0[3 * c + 1] + y02 * b0[3 * c + 2];
                acc[c] += y10 * b1[3 * c] + y11 * b1[3 * c + 1] + y12 * b1[3 * c + 2];
            }
            if (diag) rb += Ya0[18 + r] + Ya1[18 + r];
        }
        for (; t < t1; t += kPerPass) {
            const int ea = w.pair_ea[t], eb = w.pair_eb[t];
            const double* Ya = w.Y + 24 * (size_t)ea;
            const double* Wb = w.W + 18 * (size_t)eb;
            const double y0 = Ya[3 * r], y1 = Ya[3 * r + 1], y2 = Ya[3 * r + 2];
#pragma unroll
            for (int c = 0; c < 6; ++c) acc[c] += y0 * Wb[3 * c] + y1 * Wb[3 * c + 1] + y2 * Wb[3 * c + 2];
            if (diag) rb += Ya[18 + r];  // ea == eb: rhs  bs(a) -= W_a (Dinv b_l)
        }
    }
    // fixed reduction tree: the 5 slots of a warp (slot order), then the warps in order
    const int rr0 = lane % 6;
#pragma unroll
    for (int c = 0; c < 7; ++c) {
        double v = live ? (c < 6 ? acc[c < 6 ? c : 0] : rb) : 0.0;
        double s = __shfl_sync(0xffffffffu, v, rr0);
#pragma unroll
        for (int k = 1; k < 5; ++k) s += __shfl_sync(0xffffffffu, v, rr0 + 6 * k);
        if (c < 6)
            acc[c < 6 ? c : 0] = s;
        else
            rb = s;
    }
    if (lane < 6) {
#pragma unroll
        for (int c = 0; c < 6; ++c) red[warp][6 * lane + c] = acc[c];
        red[warp][36 + lane] = rb;
    }
    __syncthreads();
    if (threadIdx.x < 42) {
        double s = 0.0;
        for (int ww = 0; ww < kSchurThreads / 32; ++ww) s += red[ww][threadIdx.x];
        blockacc[threadIdx.x] = s;
    }
    __syncthreads();
    // write the 15x15 block of S (upper part of the matrix): S = Hpp + lambda I - scatter(acc)
    const int n = w.n;
    for (int i = threadIdx.x; i < 225; i += kSchurThreads) {
        const int rr = i / 15, c = i - 15 * rr;
        const int gr = 15 * a + rr, gc = 15 * b + c;
        if (gr > gc) continue;
        double v = w.Hpp[(size_t)gr * n + gc];  // (a sharded rank: its own partial sum)
        if (gr == gc && w.shard_owner) v += lambda;  // setLambda on the pose blocks (block_solver.hpp:570-577)
        const int pr = pose6_index(rr), pc = pose6_index(c);
        if (pr >= 0 && pc >= 0) v -= blockacc[6 * pr + pc];
        w.S_w[(size_t)gr * w.lds + gc] = v;
    }
    if (diag && threadIdx.x < 15) {
        const int rr = threadIdx.x;
        double v = w.bp[15 * a + rr];
        const int pr = pose6_index(rr);
        if (pr >= 0) v -= blockacc[36 + pr];
        w.bs_w[15 * a + rr] = v;
    }
    __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Schur complement, tile-scan formulation (windows with <= 32 key-frames: C1 / C3 / C5).
//
// Every thread OWNS one row of one key-frame block pair of S (pair (a <= b), row r: 6 accumulators in
// registers), like the gather -- but instead of chasing precomputed (edge_a, edge_b) lists through L2 it
// scans the window's map points tile by tile.  A tile (a few points, all their edges: one contiguous chunk
// of W) is staged in shared memory together with Y = W D^-1 and W (D^-1 b_l); a 32-bit mask of the free
// key-frames observing each point tells a thread in two instructions whether the point touches its pair,
// and a popcount of the observer mask gives the local edge index.  W is read once per CTA, coalesced; no
// atomics, fixed summation order => S is reproducible.  The pairs are cut into `sets` contiguous subsets
// (consecutive pairs share key-frame a, so the lanes of a warp hit together) and the points into `gridDim.x /
// sets` subsets whose partial sums the finish kernel adds in a fixed order.
// (landmark loop of BlockSolver::solve, block_solver.hpp:381-439)
// ------------------------------------------------------------------------------------------------
// Per trial a small kernel writes one 96-byte record per map point (D^-1, D^-1 b_l, observer masks, first
// edge inside its tile) and one 128-byte header per tile (for every key-frame the tile's points it observes),
// so that a tile is three contiguous chunks of global memory: W of its edges, its records, its header.  The
// tile kernel fetches them with bulk asynchronous copies (TMA, cp.async.bulk + mbarrier) one tile ahead.
constexpr int kTsRecDoubles = 12;  // D^-1 (6) | D^-1 b_l (3) | {observers, free observers} | {first edge, -} | pad
constexpr int kTsHdrWords = 32;    // colmask[k]: bit l = point l of the tile has an active edge to key-frame k

struct TsLayout {
    size_t w, rec, hdr, buf_bytes, total;
};
__host__ __device__ static inline TsLayout ts_layout(int K, int tile_pts) {
    TsLayout L;
    const size_t edges = (size_t)tile_pts * (K < 32 ? K : 32);
    L.w = 0;
    L.rec = edges * 144;
    L.hdr = L.rec + (size_t)tile_pts * 8 * kTsRecDoubles;
    L.buf_bytes = L.hdr + 4 * kTsHdrWords;
    L.total = 2 * L.buf_bytes + 64;
    return L;
}

__host__ __device__ static inline size_t sp_acc_doubles(int nf) { return (size_t)(nf * (nf + 1) / 2) * 36 + (size_t)nf * 6; }
size_t schur_partial_doubles(int n_free) { return sp_acc_doubles(n_free); }
bool schur_tile_fits(int K, int n_free) { return K <= 32 && n_free <= 32; }
size_t schur_tile_smem_bytes(int max_K, int tile_pts) { return ts_layout(max_K, tile_pts).total; }
size_t schur_tile_rec_doubles(int P) { return (size_t)kTsRecDoubles * P; }
size_t schur_tile_hdr_words(int P, int tile_pts) { return (size_t)kTsHdrWords * ((P + tile_pts - 1) / tile_pts) + 4; }

// one warp per tile, one lane per map point
__global__ void __launch_bounds__(256) schur_rec_kernel(const DevWindow* __restrict__ wp, int tile_pts, int factor_form) {
    const DevWindow w = wp[blockIdx.y];  // one window per grid row
    if (w.lm->phase != PH_TRIAL) return;
    const int lane = threadIdx.x & 31;
    const int ntile = (w.P + tile_pts - 1) / tile_pts;
    const double lambda = w.lm->lambda;
    const int nw = gridDim.x * (blockDim.x >> 5);
    for (int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < ntile; g += nw) {
        const int p = g * tile_pts + lane;
        const bool valid = lane < tile_pts && p < w.P;
        unsigned am = 0, fm = 0;
        if (valid) {
            const int e0 = w.pt_obs_begin[p], e1 = w.pt_obs_begin[p + 1];
            for (int e = e0; e < e1; ++e) {
                const int rw = w.obs[e].w;
                const int kf = rw & OBS_KF_MASK;
                am |= 1u << kf;
                if (!(rw & OBS_CULLED) && w.kf_block[kf] >= 0) fm |= 1u << kf;
            }
            const double* H = w.Hll + 6 * (size_t)p;
            double* d = w.ts_rec + (size_t)kTsRecDoubles * p;
            const unsigned eb = (unsigned)(e0 - w.pt_obs_begin[g * tile_pts]);
            if (factor_form) {
                // D = H_ll + lambda I = L L^T (positive definite: H_ll is a sum of J^T w J with w >= 0, lambda > 0), so
                // D^-1 = G G^T with G = L^-T upper triangular.  With Z = W G the landmark term of the reduced system is
                // W_a D^-1 W_b^T = Z_a Z_b^T and W_a D^-1 b_l = Z_a (G^T b_l): one operand array instead of W and W D^-1.
                const double xx = H[0] + lambda, xy = H[1], xz = H[2], yy = H[3] + lambda, yz = H[4], zz = H[5] + lambda;
                const double m00 = rsqrt(xx);
                const double l10 = xy * m00, l20 = xz * m00;
                const double m11 = rsqrt(yy - l10 * l10);
                const double l21 = (yz - l20 * l10) * m11;
                const double m22 = rsqrt(zz - l20 * l20 - l21 * l21);
                double g00 = m00, g11 = m11, g22 = m22;
                double g01 = -l10 * m00 * m11;            // (L^-1)(1,0)
                double g12 = -l21 * m11 * m22;            // (L^-1)(2,1)
                double g02 = -(l20 * m00 + l21 * g01) * m22;  // (L^-1)(2,0)
                if (!(isfinite(g00) && isfinite(g11) && isfinite(g22)))  // not positive definite: the point drops out
                    g00 = g01 = g02 = g11 = g12 = g22 = 0.0;
                const V3 b = ld3(w.bl + 3 * (size_t)p);
                unsigned* u = reinterpret_cast<unsigned*>(d);
                u[0] = am, u[1] = fm, u[2] = eb, u[3] = (unsigned)(e1 - e0);
                d[2] = g00, d[3] = g01, d[4] = g02, d[5] = g11, d[6] = g12, d[7] = g22;
                d[8] = g00 * b.x, d[9] = fma(g11, b.y, g01 * b.x), d[10] = fma(g22, b.z, fma(g12, b.y, g02 * b.x));  // G^T b_l
                d[11] = 0.0;
            } else {
                bool ok;
                const S3 Dinv = s3_inverse(S3{H[0] + lambda, H[1], H[2], H[3] + lambda, H[4], H[5] + lambda}, ok);
                const V3 db = s3_mul(Dinv, ld3(w.bl + 3 * (size_t)p));
                d[0] = Dinv.xx, d[1] = Dinv.xy, d[2] = Dinv.xz, d[3] = Dinv.yy, d[4] = Dinv.yz, d[5] = Dinv.zz;
                d[6] = db.x, d[7] = db.y, d[8] = db.z;
                unsigned* u = reinterpret_cast<unsigned*>(d + 9);
                u[0] = am, u[1] = fm, u[2] = eb, u[3] = 0u;
            }
        }
        unsigned mine = 0;
#pragma unroll 4
        for (int k = 0; k < 32; ++k) {
            const unsigned bal = __ballot_sync(0xffffffffu, valid && ((fm >> k) & 1u));
            if (lane == k) mine = bal;
        }
        w.ts_hdr[(size_t)kTsHdrWords * g + lane] = mine;
    }
}

// Work balance: a block pair (a, b) is touched by a point only if both key-frames observe it, which in a sliding
// window is likely for close key-frames and rare for distant ones.  Every lane group therefore owns TWO pairs:
// with the pairs ordered by distance d = b - a, group j takes the j-th closest and the j-th farthest, so the
// hit counts of the groups (and of the warps) are nearly equal.
__device__ __forceinline__ void ts_pair_by_distance(int q, int nf, int& a, int& b) {
    int d = 0, off = 0;  // pairs with distance < d: d nf - d (d - 1) / 2
    while (d < nf && off + (nf - d) <= q) off += nf - d, ++d;
    a = q - off;
    b = a + d;
}

__global__ void __launch_bounds__(512) schur_tile_kernel(const DevWindow* __restrict__ wp, int sets, int tile_pts) {
    const DevWindow w = wp[blockIdx.y];  // one window per grid row
    if (w.lm->phase != PH_TRIAL) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar[2];
    const TsLayout L = ts_layout(w.K, tile_pts);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int set = blockIdx.x % sets, psub = blockIdx.x / sets, npsub = gridDim.x / sets;
    // the two rows of S this thread owns
    const int nf = w.n_free;
    const int ngroups = (w.n_pairs + 1) / 2;
    const int gpc = (ngroups + sets - 1) / sets;  // lane groups per set
    const int slot = lane / 6, r = lane - 6 * slot;
    const int grp_local = warp * 5 + slot;
    const int grp = set * gpc + grp_local;
    const bool live = lane < 30 && grp_local < gpc && grp < ngroups;
    int pa[2] = {0, 0}, pb[2] = {0, 0}, ka[2] = {0, 0}, kb[2] = {0, 0};
    bool own[2] = {false, false};
    if (live) {
        const int q2 = w.n_pairs - 1 - grp;
        ts_pair_by_distance(grp, nf, pa[0], pb[0]);
        own[0] = true;
        if (q2 > grp) {
            ts_pair_by_distance(q2, nf, pa[1], pb[1]);
            own[1] = true;
        }
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) ka[s2] = w.blk_kf[pa[s2]], kb[s2] = w.blk_kf[pb[s2]];
    }
    double acc[2][6] = {{0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0}};
    double rb[2] = {0.0, 0.0};
    const int ntile = (w.P + tile_pts - 1) / tile_pts;
    const int g_begin = (int)((long long)ntile * psub / npsub), g_end = (int)((long long)ntile * (psub + 1) / npsub);
    const int nt = g_end - g_begin;
    const bool leader = threadIdx.x == 0;
    if (leader) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    // leader: bulk copies of tile g (edge range [e_lo, e_hi)) into buffer `b2`
    auto issue = [&](int g, int b2, int e_lo, int e_hi) {
        unsigned char* buf = smem_raw + (size_t)b2 * L.buf_bytes;
        const int p0 = g * tile_pts, np = min(tile_pts, w.P - p0);
        const unsigned wbytes = 144u * (unsigned)(e_hi - e_lo), rbytes = 8u * kTsRecDoubles * (unsigned)np, hbytes = 4u * kTsHdrWords;
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // earlier generic-proxy reads of the buffer
        mbar_expect_tx(&bar[b2], wbytes + rbytes + hbytes);
        if (wbytes) tma_load_1d(buf + L.w, w.W + 18 * (size_t)e_lo, wbytes, &bar[b2]);
        tma_load_1d(buf + L.rec, w.ts_rec + (size_t)kTsRecDoubles * p0, rbytes, &bar[b2]);
        tma_load_1d(buf + L.hdr, w.ts_hdr + (size_t)kTsHdrWords * g, hbytes, &bar[b2]);
    };
    auto tile_edges = [&](int t, int& e_lo, int& e_hi) {  // edge range of the t-th tile of this CTA
        e_lo = e_hi = 0;
        if (leader && t < nt) {
            const int p0 = (g_begin + t) * tile_pts;
            e_lo = w.pt_obs_begin[p0];
            e_hi = w.pt_obs_begin[min(p0 + tile_pts, w.P)];
        }
    };
    int e_lo_n, e_hi_n, e_lo_nn, e_hi_nn;
    tile_edges(0, e_lo_n, e_hi_n);
    if (leader && nt > 0) issue(g_begin, 0, e_lo_n, e_hi_n);
    tile_edges(1, e_lo_n, e_hi_n);

    for (int t = 0; t < nt; ++t) {
        if (leader && t + 1 < nt) issue(g_begin + t + 1, (t + 1) & 1, e_lo_n, e_hi_n);  // its buffer was released by the barrier below
        tile_edges(t + 2, e_lo_nn, e_hi_nn);  // in flight until the next iteration needs them
        const unsigned char* buf = smem_raw + (size_t)(t & 1) * L.buf_bytes;
        const double* bW = reinterpret_cast<const double*>(buf + L.w);
        const double* bRec = reinterpret_cast<const double*>(buf + L.rec);
        const unsigned* colmask = reinterpret_cast<const unsigned*>(buf + L.hdr);
        mbar_wait(&bar[t & 1], (unsigned)((t >> 1) & 1));
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
            // the points of the tile that touch this block pair: two loads and an AND, then only the hits
            unsigned hits = own[s2] ? (colmask[ka[s2]] & colmask[kb[s2]]) : 0u;
            const unsigned below_a = (1u << ka[s2]) - 1u, below_b = (1u << kb[s2]) - 1u;
            const bool diag = pa[s2] == pb[s2];
            while (hits) {
                const int l = __ffs(hits) - 1;
                hits &= hits - 1;
                const double* d = bRec + kTsRecDoubles * l;
                const uint2 mk = *reinterpret_cast<const uint2*>(d + 9);
                const unsigned eb = *reinterpret_cast<const unsigned*>(d + 10);
                const double* Wi = bW + 18 * (eb + __popc(mk.x & below_a)) + 3 * r;
                const double* Wj = bW + 18 * (eb + __popc(mk.x & below_b));
                const double w0 = Wi[0], w1 = Wi[1], w2 = Wi[2];
                // row r of W_i D^-1 (BDinv, block_solver.hpp:407)
                const double y0 = fma(d[2], w2, fma(d[1], w1, d[0] * w0));
                const double y1 = fma(d[4], w2, fma(d[3], w1, d[1] * w0));
                const double y2 = fma(d[5], w2, fma(d[4], w1, d[2] * w0));
                // the 6 row lanes of a pair read the same W_j (a broadcast)
                double wj[18];
#pragma unroll
                for (int k = 0; k < 18; ++k) wj[k] = Wj[k];
#pragma unroll
                for (int c = 0; c < 6; ++c)
                    acc[s2][c] = fma(y2, wj[3 * c + 2], fma(y1, wj[3 * c + 1], fma(y0, wj[3 * c], acc[s2][c])));
                if (diag) rb[s2] = fma(w2, d[8], fma(w1, d[7], fma(w0, d[6], rb[s2])));  // rhs: b_s(a) -= W_a (D^-1 b_l)
            }
        }
        e_lo_n = e_lo_nn, e_hi_n = e_hi_nn;
        __syncthreads();  // tile t consumed: its buffer may be refilled
    }
    double* part = w.schur_partial + (size_t)psub * sp_acc_doubles(nf);
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
        if (!own[s2]) continue;
        const int a = pa[s2], b = pb[s2];
        double* d = part + (size_t)(a * nf - a * (a - 1) / 2 + (b - a)) * 36 + 6 * r;
#pragma unroll
        for (int c = 0; c < 6; ++c) d[c] = acc[s2][c];
        if (a == b) part[(size_t)w.n_pairs * 36 + 6 * a + r] = rb[s2];
    }
}

// ------------------------------------------------------------------------------------------------
// Tensor-pipe variant: ONE WARP PER HIT.  A block pair's contribution of one point is the 6x6x3 product
// Z_a Z_b^T, which is one m8n8k4 FP64 MMA with the operands padded by zeros: every lane loads ONE double of Z_a
// and ONE of Z_b (18 distinct words of a 144-byte block, no replication across lanes) where the row-per-lane
// kernels load 21 and issue 21 DFMA per lane.  The FP64 pipe does the same number of multiply-adds per clock
// either way (tools/ubench_fp64.cu); what the MMA buys is 14 instead of 66 issue slots per warp-level hit, half
// the shared-memory wavefronts, and no idle lane groups.  Column 6 of the B operand carries G^T b_l, so that the
// diagonal pairs produce the right-hand side Z_a (G^T b_l) in column 6 of their accumulator for free.
// Measured (profiles/r2_schur_variants.md): parity-identical, 1.04 ms per 64-window launch against 0.96 ms for the
// row-per-lane-group kernel -- the per-(pair, tile) set-up and the list round trip cost what the MMA saves, so it is
// an option (env VILBA_SP_MMA=1), not the default.
// A warp owns up to kMmaPairs block pairs (two accumulator registers per lane and pair), interleaved by distance
// so that close (many hits) and distant pairs mix.  Per tile the lanes first compute, each for one map point, the
// packed edge indices of (a, b) -- the bookkeeping of ALL hits of the pair in the tile in one go -- and the hit loop
// takes them with a shuffle.
// ------------------------------------------------------------------------------------------------
constexpr int kMmaPairs = 6;
struct TmLayout {
    size_t rec, hdr, buf_bytes, cpad, zero, ents, total;
};
__host__ __device__ static inline TmLayout tm_layout(int tile_edges, int tile_pts) {
    TmLayout L;
    L.rec = (size_t)tile_edges * 144;
    L.hdr = L.rec + (size_t)tile_pts * 8 * kTsRecDoubles;
    L.buf_bytes = (L.hdr + 4 * kTsHdrWords + 127) / 128 * 128;
    L.cpad = 2 * L.buf_bytes;                         // G^T b_l per edge (24 bytes), current tile only
    L.zero = L.cpad + ((size_t)tile_edges * 24 + 15) / 16 * 16;
    L.ents = L.zero + 16;                             // per warp: the packed hit list of the pair at hand (32 words)
    L.total = L.ents + 16 * 128 + 64;
    return L;
}
size_t schur_mma_smem_bytes(int tile_edges, int tile_pts) { return tm_layout(tile_edges, tile_pts).total; }
int schur_mma_units(int n_free) { return (n_free * (n_free + 1) / 2 + kMmaPairs - 1) / kMmaPairs; }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(512, 2) schur_mma_kernel(const DevWindow* __restrict__ wp, int sets, int tile_pts, int tile_edges) {
    const DevWindow w = wp[blockIdx.y];  // one window per grid row
    if (w.lm->phase != PH_TRIAL) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar[2];
    const TmLayout L = tm_layout(tile_edges, tile_pts);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const int set = blockIdx.x % sets, psub = blockIdx.x / sets, npsub = gridDim.x / sets;
    const int nf = w.n_free;
    const int units = sets * warps, unit = set * warps + warp;
    // the block pairs of this warp: q = i * units + unit in the order of increasing distance b - a
    unsigned pk[kMmaPairs];  // key-frame index of a | key-frame index of b << 8
    int my_pairs = 0;
#pragma unroll
    for (int i = 0; i < kMmaPairs; ++i) {
        pk[i] = 0;
        const int q = i * units + unit;
        if (q < w.n_pairs) {
            int a, b;
            ts_pair_by_distance(q, nf, a, b);
            pk[i] = (unsigned)w.blk_kf[a] | ((unsigned)w.blk_kf[b] << 8);
            my_pairs = i + 1;
        }
    }
    double acc[kMmaPairs][2];
#pragma unroll
    for (int i = 0; i < kMmaPairs; ++i) acc[i][0] = acc[i][1] = 0.0;
    // operand fragments of mma.m8n8k4: A(row = lane / 4, k = lane % 4), B(k = lane % 4, col = lane / 4)
    const int frow = lane >> 2, fk = lane & 3;
    const bool zlane = frow < 6 && fk < 3;       // an entry of the 6x3 block
    const bool clane = frow == 6 && fk < 3;      // B only: G^T b_l in column 6
    const unsigned zoff = (unsigned)(3 * frow + fk) * 8u;
    const unsigned strideA = zlane ? 144u : 0u, strideB = zlane ? 144u : (clane ? 24u : 0u);
    const unsigned smem0 = smem_u32(smem_raw);
    const unsigned zero_addr = smem0 + (unsigned)L.zero;
    if (threadIdx.x < 2) reinterpret_cast<double*>(smem_raw + L.zero)[threadIdx.x] = 0.0;
    unsigned* ents = reinterpret_cast<unsigned*>(smem_raw + L.ents) + 32 * warp;
    if (lane < 32) ents[lane] = 0u;

    const int ntile = (w.P + tile_pts - 1) / tile_pts;
    const int g_begin = (int)((long long)ntile * psub / npsub), g_end = (int)((long long)ntile * (psub + 1) / npsub);
    const int nt = g_end - g_begin;
    const bool leader = threadIdx.x == 0;
    if (leader) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int g, int b2, int e_lo, int e_hi) {
        unsigned char* buf = smem_raw + (size_t)b2 * L.buf_bytes;
        const int p0 = g * tile_pts, np = min(tile_pts, w.P - p0);
        const unsigned wbytes = 144u * (unsigned)(e_hi - e_lo), rbytes = 8u * kTsRecDoubles * (unsigned)np, hbytes = 4u * kTsHdrWords;
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // earlier generic-proxy accesses of the buffer
        mbar_expect_tx(&bar[b2], wbytes + rbytes + hbytes);
        if (wbytes) tma_load_1d(buf, w.W + 18 * (size_t)e_lo, wbytes, &bar[b2]);
        tma_load_1d(buf + L.rec, w.ts_rec + (size_t)kTsRecDoubles * p0, rbytes, &bar[b2]);
        tma_load_1d(buf + L.hdr, w.ts_hdr + (size_t)kTsHdrWords * g, hbytes, &bar[b2]);
    };
    auto tile_edges_of = [&](int t, int& e_lo, int& e_hi) {
        e_lo = e_hi = 0;
        if (leader && t < nt) {
            const int p0 = (g_begin + t) * tile_pts;
            e_lo = w.pt_obs_begin[p0];
            e_hi = w.pt_obs_begin[min(p0 + tile_pts, w.P)];
        }
    };
    int e_lo_n, e_hi_n, e_lo_nn, e_hi_nn;
    tile_edges_of(0, e_lo_n, e_hi_n);
    if (leader && nt > 0) issue(g_begin, 0, e_lo_n, e_hi_n);
    tile_edges_of(1, e_lo_n, e_hi_n);

    for (int t = 0; t < nt; ++t) {
        if (leader && t + 1 < nt) issue(g_begin + t + 1, (t + 1) & 1, e_lo_n, e_hi_n);
        tile_edges_of(t + 2, e_lo_nn, e_hi_nn);
        unsigned char* buf = smem_raw + (size_t)(t & 1) * L.buf_bytes;
        double* bZ = reinterpret_cast<double*>(buf);
        const double* bRec = reinterpret_cast<const double*>(buf + L.rec);
        const unsigned* colmask = reinterpret_cast<const unsigned*>(buf + L.hdr);
        double* cpad = reinterpret_cast<double*>(smem_raw + L.cpad);
        const int np = min(tile_pts, w.P - (g_begin + t) * tile_pts);
        mbar_wait(&bar[t & 1], (unsigned)((t >> 1) & 1));
        // ---- W -> Z = W G in place, G^T b_l beside every edge: task = (point, row of the 6x3 blocks, one edge in four) ----
        for (int q = threadIdx.x; q < np * 24; q += blockDim.x) {
            const int l = q / 24, rem = q - 24 * l, j = rem >> 2, sl = rem & 3;
            const double* d = bRec + kTsRecDoubles * l;
            const uint4 hd = *reinterpret_cast<const uint4*>(d);
            const double2 ga = *reinterpret_cast<const double2*>(d + 2), gb = *reinterpret_cast<const double2*>(d + 4),
                          gc = *reinterpret_cast<const double2*>(d + 6);
            const double c0 = d[8], c1 = d[9], c2 = d[10];
            for (unsigned e = sl; e < hd.w; e += 4) {
                double* z = bZ + 18 * (hd.z + e) + 3 * j;
                const double w0 = z[0], w1 = z[1], w2 = z[2];
                z[0] = w0 * ga.x;
                z[1] = fma(w1, gb.y, w0 * ga.y);
                z[2] = fma(w2, gc.y, fma(w1, gc.x, w0 * gb.x));
                if (j == 0) {
                    double* c = cpad + 3 * (hd.z + e);
                    c[0] = c0, c[1] = c1, c[2] = c2;
                }
            }
        }
        __syncthreads();
        // this lane's map point: observers and first edge
        unsigned am_l = 0, eb_l = 0;
        if (lane < np) {
            const uint4 hd = *reinterpret_cast<const uint4*>(bRec + kTsRecDoubles * lane);
            am_l = hd.x, eb_l = hd.z;
        }
        const unsigned bufaddr = smem0 + (unsigned)((size_t)(t & 1) * L.buf_bytes);
        const unsigned baseA = zlane ? bufaddr + zoff : zero_addr;
        const unsigned baseB = zlane ? bufaddr + zoff : (clane ? smem0 + (unsigned)L.cpad + 8u * fk : zero_addr);
        const unsigned lt = (1u << lane) - 1u;
#pragma unroll
        for (int i = 0; i < kMmaPairs; ++i) {
            if (i >= my_pairs) break;  // warp-uniform
            const unsigned ka = pk[i] & 255u, kb = pk[i] >> 8;
            const unsigned hits = colmask[ka] & colmask[kb];
            if (!hits) continue;
            // edge of a | edge of b << 16 for the point of this lane; the hits are packed to the front of the warp's list
            const unsigned ent = (eb_l + __popc(am_l & ((1u << ka) - 1u))) | ((eb_l + __popc(am_l & ((1u << kb) - 1u))) << 16);
            __syncwarp();  // the list of the previous pair has been read
            if ((hits >> lane) & 1u) ents[__popc(hits & lt)] = ent;
            __syncwarp();
            const int n = __popc(hits);
            double c0 = acc[i][0], c1 = acc[i][1];
            for (int j = 0; j < n; j += 2) {  // two hits per round: four independent loads in flight
                const uint2 e2 = *reinterpret_cast<const uint2*>(ents + j);
                double a0, b0, a1, b1;
                asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(a0) : "r"(baseA + (e2.x & 0xffffu) * strideA));
                asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(b0) : "r"(baseB + (e2.x >> 16) * strideB));
                asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(a1) : "r"(baseA + (e2.y & 0xffffu) * strideA));
                asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(b1) : "r"(baseB + (e2.y >> 16) * strideB));
                dmma884(c0, c1, a0, b0);
                if (j + 1 < n) dmma884(c0, c1, a1, b1);  // (the entry beyond the list is stale but a valid edge of the tile)
            }
            acc[i][0] = c0, acc[i][1] = c1;
        }
        e_lo_n = e_lo_nn, e_hi_n = e_hi_nn;
        __syncthreads();  // tile t consumed: its buffer may be refilled
    }
    // accumulator fragment: C(row = lane / 4, col = 2 (lane % 4) + {0, 1}); column 6 of a diagonal pair is the rhs
    double* part = w.schur_partial + (size_t)psub * sp_acc_doubles(nf);
#pragma unroll
    for (int i = 0; i < kMmaPairs; ++i) {
        const int q = i * units + unit;
        if (i >= my_pairs || q >= w.n_pairs) break;
        int a, b;
        ts_pair_by_distance(q, nf, a, b);
        if (frow < 6) {
            double* d = part + (size_t)(a * nf - a * (a - 1) / 2 + (b - a)) * 36 + 6 * frow;
            if (fk < 3) d[2 * fk] = acc[i][0], d[2 * fk + 1] = acc[i][1];
            else if (a == b) part[(size_t)w.n_pairs * 36 + 6 * a + frow] = acc[i][0];
        }
    }
}

// Variant with ONE LANE PER BLOCK PAIR: the lane keeps the whole 6x6 block (36 accumulators) in registers, so the
// per-hit bookkeeping (mask, popcounts, addresses, record loads) is paid once per 108 + 54 FMAs instead of once per
// 18 + 9, and the 36 chains are independent.  One CTA covers all pairs of a window (W is read once); the points are
// split over gridDim.x CTAs.  Balance: pairs of close key-frames are hit by many more points than distant ones, so a
// pair at distance d gets R(d) = 3, 2 or 1 lanes that share its hits (point l of a tile goes to lane l mod R); the lanes
// are ordered by distance, so the lanes of a warp have similar hit counts.
__host__ __device__ static inline size_t schur_pair_partial_doubles_dev(int nf);
__host__ __device__ static inline int ts_replicas(int d, int nf) {
    const int x = 100 * d / (nf + 1);  // distance in percent of the window length
    return x < 14 ? 3 : (x < 41 ? 2 : 1);
}
__host__ __device__ static inline int ts_pair_lanes(int nf) {
    int lanes = 0;
    for (int d = 0; d < nf; ++d) lanes += (nf - d) * ts_replicas(d, nf);
    return lanes;
}
// lane -> (a, b, replica k of R); returns false beyond the last lane
__device__ __forceinline__ bool ts_lane_to_pair(int lane, int nf, int& a, int& b, int& k, int& R) {
    int off = 0;
    for (int d = 0; d < nf; ++d) {
        R = ts_replicas(d, nf);
        const int cnt = (nf - d) * R;
        if (lane < off + cnt) {
            const int idx = lane - off;
            a = idx / R;
            k = idx - a * R;
            b = a + d;
            return true;
        }
        off += cnt;
    }
    return false;
}
__device__ __forceinline__ int ts_pair_first_lane(int a, int b, int nf) {
    const int d = b - a;
    int off = 0;
    for (int dd = 0; dd < d; ++dd) off += (nf - dd) * ts_replicas(dd, nf);
    return off + a * ts_replicas(d, nf);
}
__host__ __device__ static inline size_t schur_pair_partial_doubles_dev(int nf) { return (size_t)ts_pair_lanes(nf) * 36 + (size_t)nf * 6 * 3; }
int schur_pair_lanes(int n_free) { return ts_pair_lanes(n_free); }
size_t schur_pair_partial_doubles(int n_free) { return schur_pair_partial_doubles_dev(n_free); }

__global__ void __launch_bounds__(384) schur_tile_pair_kernel(const DevWindow* __restrict__ wp, int tile_pts) {
    const DevWindow w = wp[blockIdx.y];  // one window per grid row
    if (w.lm->phase != PH_TRIAL) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar[2];
    const TsLayout L = ts_layout(w.K, tile_pts);
    const int psub = blockIdx.x, npsub = gridDim.x;
    const int nf = w.n_free;
    int a = 0, b = 0, rk = 0, R = 1;
    const bool live = ts_lane_to_pair(threadIdx.x, nf, a, b, rk, R);
    int ka = 0, kb = 0;
    if (live) ka = w.blk_kf[a], kb = w.blk_kf[b];
    const bool diag = live && a == b;
    const unsigned below_a = (1u << ka) - 1u, below_b = (1u << kb) - 1u;
    // points l of a tile with l mod R == rk
    const unsigned mine = R == 1 ? 0xffffffffu : (R == 2 ? (0x55555555u << rk) : (0x49249249u << rk));
    double acc[36];
    double rb[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 36; ++i) acc[i] = 0.0;
    const int ntile = (w.P + tile_pts - 1) / tile_pts;
    const int g_begin = (int)((long long)ntile * psub / npsub), g_end = (int)((long long)ntile * (psub + 1) / npsub);
    const int nt = g_end - g_begin;
    const bool leader = threadIdx.x == 0;
    if (leader) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int g, int b2, int e_lo, int e_hi) {
        unsigned char* buf = smem_raw + (size_t)b2 * L.buf_bytes;
        const int p0 = g * tile_pts, np = min(tile_pts, w.P - p0);
        const unsigned wbytes = 144u * (unsigned)(e_hi - e_lo), rbytes = 8u * kTsRecDoubles * (unsigned)np, hbytes = 4u * kTsHdrWords;
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        mbar_expect_tx(&bar[b2], wbytes + rbytes + hbytes);
        if (wbytes) tma_load_1d(buf + L.w, w.W + 18 * (size_t)e_lo, wbytes, &bar[b2]);
        tma_load_1d(buf + L.rec, w.ts_rec + (size_t)kTsRecDoubles * p0, rbytes, &bar[b2]);
        tma_load_1d(buf + L.hdr, w.ts_hdr + (size_t)kTsHdrWords * g, hbytes, &bar[b2]);
    };
    auto tile_edges = [&](int t, int& e_lo, int& e_hi) {
        e_lo = e_hi = 0;
        if (leader && t < nt) {
            const int p0 = (g_begin + t) * tile_pts;
            e_lo = w.pt_obs_begin[p0];
            e_hi = w.pt_obs_begin[min(p0 + tile_pts, w.P)];
        }
    };
    int e_lo_n, e_hi_n, e_lo_nn, e_hi_nn;
    tile_edges(0, e_lo_n, e_hi_n);
    if (leader && nt > 0) issue(g_begin, 0, e_lo_n, e_hi_n);
    tile_edges(1, e_lo_n, e_hi_n);
    for (int t = 0; t < nt; ++t) {
        if (leader && t + 1 < nt) issue(g_begin + t + 1, (t + 1) & 1, e_lo_n, e_hi_n);
        tile_edges(t + 2, e_lo_nn, e_hi_nn);
        const unsigned char* buf = smem_raw + (size_t)(t & 1) * L.buf_bytes;
        const double* bW = reinterpret_cast<const double*>(buf + L.w);
        const double* bRec = reinterpret_cast<const double*>(buf + L.rec);
        const unsigned* colmask = reinterpret_cast<const unsigned*>(buf + L.hdr);
        mbar_wait(&bar[t & 1], (unsigned)((t >> 1) & 1));
        unsigned hits = live ? (colmask[ka] & colmask[kb] & mine) : 0u;
#ifdef VILBA_TS_ABLATE
        if (w.dbg_flags & 2) hits = 0u;
#endif
        while (hits) {
            const int l = __ffs(hits) - 1;
            hits &= hits - 1;
            const double* d = bRec + kTsRecDoubles * l;
            const uint2 mk = *reinterpret_cast<const uint2*>(d + 9);
            const unsigned eb = *reinterpret_cast<const unsigned*>(d + 10);
            const double* Wi = bW + 18 * (eb + __popc(mk.x & below_a));
            const double2* Wj = reinterpret_cast<const double2*>(bW + 18 * (eb + __popc(mk.x & below_b)));
            const double dxx = d[0], dxy = d[1], dxz = d[2], dyy = d[3], dyz = d[4], dzz = d[5];
            double wj[18];
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const double2 v = Wj[i];
                wj[2 * i] = v.x, wj[2 * i + 1] = v.y;
            }
            const double b0 = d[6], b1 = d[7], b2 = d[8];
#pragma unroll
            for (int r = 0; r < 6; ++r) {  // row r of W_i D^-1 (BDinv, block_solver.hpp:407), then row r of the block
                const double w0 = Wi[3 * r], w1 = Wi[3 * r + 1], w2 = Wi[3 * r + 2];
                const double y0 = fma(dxz, w2, fma(dxy, w1, dxx * w0));
                const double y1 = fma(dyz, w2, fma(dyy, w1, dxy * w0));
                const double y2 = fma(dzz, w2, fma(dyz, w1, dxz * w0));
#pragma unroll
                for (int c = 0; c < 6; ++c)
                    acc[6 * r + c] = fma(y2, wj[3 * c + 2], fma(y1, wj[3 * c + 1], fma(y0, wj[3 * c], acc[6 * r + c])));
                if (diag) rb[r] = fma(w2, b2, fma(w1, b1, fma(w0, b0, rb[r])));  // rhs: b_s(a) -= W_a (D^-1 b_l)
            }
        }
        e_lo_n = e_lo_nn, e_hi_n = e_hi_nn;
        __syncthreads();  // tile t consumed: its buffer may be refilled
    }
    if (live) {
        double* part = w.schur_partial + (size_t)psub * schur_pair_partial_doubles_dev(nf);
        double* dst = part + (size_t)threadIdx.x * 36;
#pragma unroll
        for (int i = 0; i < 36; ++i) dst[i] = acc[i];
        if (diag) {
#pragma unroll
            for (int r = 0; r < 6; ++r) part[(size_t)ts_pair_lanes(nf) * 36 + (6 * a + r) * 3 + rk] = rb[r];
        }
    }
}

// finish for the lane-per-pair layout: S = H_pp + lambda I - sum over point subsets and replicas (fixed order)
__global__ void __launch_bounds__(256) schur_finish_pair_kernel(const DevWindow* __restrict__ wp, int point_ctas) {
    const DevWindow w = wp[blockIdx.y];  // one window per grid row
    if (w.lm->phase != PH_TRIAL) return;
    const int n = w.n, nf = w.n_free;
    const double lambda = w.lm->lambda;
    const size_t accN = schur_pair_partial_doubles_dev(nf);
    const size_t rhs0 = (size_t)ts_pair_lanes(nf) * 36;
    const size_t total = (size_t)n * n + n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const bool is_rhs = i >= (size_t)n * n;
        int gr, gc;
        if (is_rhs) {
            gr = gc = (int)(i - (size_t)n * n);
        } else {
            gr = (int)(i / n);
            gc = (int)(i - (size_t)gr * n);
            if (gr > gc) continue;
        }
        const int a = gr / 15, b = gc / 15, rr = gr - 15 * a, cc = gc - 15 * b;
        const int pr = pose6_index(rr), pc = pose6_index(cc);
        double v = is_rhs ? w.bp[gr] : w.Hpp[(size_t)gr * n + gc];  // (a sharded rank: its own partial sums)
        if (!is_rhs && gr == gc && w.shard_owner) v += lambda;  // setLambda on the pose blocks (block_solver.hpp:570-577)
        if (pr >= 0 && (is_rhs || pc >= 0)) {
            const int R = ts_replicas(b - a, nf);
            const size_t off = is_rhs ? rhs0 + (size_t)(6 * a + pr) * 3 : (size_t)ts_pair_first_lane(a, b, nf) * 36 + 6 * pr + pc;
            const size_t step = is_rhs ? 1 : 36;
            double sum = 0.0;
            for (int c = 0; c < point_ctas; ++c)
                for (int k = 0; k < R; ++k) sum += w.schur_partial[(size_t)c * accN + off + step * k];
            v -= sum;
        }
        if (is_rhs)
            w.bs_w[gr] = v;
        else
            w.S_w[(size_t)gr * w.lds + gc] = v;
    }
}

// S = H_pp + lambda I - sum of the CTA partials (fixed order), b_s = b_p - sum; upper triangle only.
// Only the [P,Phi] x [P,Phi] entries of a block pair (36 of 225) and 6 of 15 rhs entries carry a landmark term: four
// lanes share the partial sums of one such entry (lane k adds the point subsets k, k + 4, ... in order, then
// (l0 + l1) + (l2 + l3): a fixed tree), every other entry is a copy.  (One thread per entry of S looping over all
// point subsets took 29 us for a single window with 74 subsets -- longer than the tile kernel that produced them.)
__global__ void __launch_bounds__(256) schur_finish_kernel(const DevWindow* __restrict__ wp, int point_ctas) {
    const DevWindow w = wp[blockIdx.y];  // one window per grid row
    if (w.lm->phase != PH_TRIAL) return;
    const int n = w.n, nf = w.n_free;
    const double lambda = w.lm->lambda;
    const size_t accN = sp_acc_doubles(nf);
    const size_t pairsN = (size_t)w.n_pairs * 36;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    // ---- entries with a landmark term: accumulator entry j -> (row, column) of S, or a rhs entry ----
    const int sub = threadIdx.x & 3;
    const int quad = (threadIdx.x & 31) >> 2;  // the four lanes of an entry are neighbours; a warp takes 8 entries per trip
    for (size_t j0 = (tid >> 2) - quad; j0 < accN; j0 += nthreads >> 2) {  // (warp-uniform trip count: shuffles inside)
        const size_t j = j0 + quad;
        const bool valid = j < accN;
        double sum = 0.0;
        if (valid)
            for (int c = sub; c < point_ctas; c += 4) sum += w.schur_partial[(size_t)c * accN + j];
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        if (sub || !valid) continue;
        if (j < pairsN) {
            const int pair = (int)(j / 36), e = (int)(j - 36 * (size_t)pair), pr = e / 6, pc = e - 6 * pr;
            int a = 0, off = 0;  // pair -> (a, b): pairs of row a start at a nf - a (a - 1) / 2
            while (a + 1 < nf && off + (nf - a) <= pair) off += nf - a, ++a;
            const int b = a + (pair - off);
            const int gr = 15 * a + (pr < 3 ? pr : pr + 3), gc = 15 * b + (pc < 3 ? pc : pc + 3);
            if (gr > gc) continue;  // lower part of a diagonal block
            double v = w.Hpp[(size_t)gr * n + gc];
            if (gr == gc && w.shard_owner) v += lambda;  // setLambda on the pose blocks (block_solver.hpp:570-577)
            w.S_w[(size_t)gr * w.lds + gc] = v - sum;
        } else {
            const int e = (int)(j - pairsN), a = e / 6, pr = e - 6 * a;
            const int gr = 15 * a + (pr < 3 ? pr : pr + 3);
            w.bs_w[gr] = w.bp[gr] - sum;
        }
    }
    // ---- every other entry of the upper triangle and of the rhs: H_pp + lambda I, b_p ----
    const size_t total = (size_t)n * n + n;
    for (size_t i = tid; i < total; i += nthreads) {
        const bool is_rhs = i >= (size_t)n * n;
        int gr, gc;
        if (is_rhs) {
            gr = gc = (int)(i - (size_t)n * n);
        } else {
            gr = (int)(i / n);
            gc = (int)(i - (size_t)gr * n);
            if (gr > gc) continue;
        }
        const int rr = gr % 15, cc = gc % 15;
        const int pr = pose6_index(rr), pc = pose6_index(cc);
        if (pr >= 0 && (is_rhs || pc >= 0)) continue;  // done above
        if (is_rhs) {
            w.bs_w[gr] = w.bp[gr];
        } else {
            double v = w.Hpp[(size_t)gr * n + gc];
            if (gr == gc && w.shard_owner) v += lambda;
            w.S_w[(size_t)gr * w.lds + gc] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
size_t linearize_smem_bytes(int K, int n_free, int threads) { return linearize_v2_smem_bytes(K, n_free, threads / 32); }

cudaError_t launch_update_eval_apply(cudaStream_t s, const DevWindow* wp, const LaunchDims& d);
cudaError_t launch_lm_iter_begin(cudaStream_t s, const DevWindow* wp, const LaunchDims& d);
cudaError_t launch_lm_decide(cudaStream_t s, const DevWindow* wp, const LaunchDims& d);
cudaError_t launch_shard_diag(cudaStream_t s, const DevWindow* wp, const LaunchDims& d);
cudaError_t launch_shard_scale(cudaStream_t s, const DevWindow* wp, const LaunchDims& d);
cudaError_t configure_point_kernels(const LaunchDims& d);

cudaError_t configure_kernels(const LaunchDims& d) {
    cudaError_t e = opt_in_max_smem(linearize_v2_kernel);
    if (e != cudaSuccess) return e;
    e = opt_in_max_smem(schur_tile_kernel);
    if (e != cudaSuccess) return e;
    e = opt_in_max_smem(schur_mma_kernel);
    if (e != cudaSuccess) return e;
    e = opt_in_max_smem(schur_tile_pair_kernel);
    if (e != cudaSuccess) return e;
    e = configure_point_kernels(d);
    if (e != cudaSuccess) return e;
    if ((e = configure_chol_big(0)) != cudaSuccess) return e;
    if ((e = configure_chol_la()) != cudaSuccess) return e;
    return cudaSuccess;
}

cudaError_t launch_slot(cudaStream_t s, cudaStream_t side, cudaEvent_t fork, cudaEvent_t join, const DevWindow* wp,
                        const LaunchDims& d, cudaEvent_t* probe, const SlotComm* comm) {
    cudaError_t e;
    if (probe && (e = cudaEventRecord(probe[0], s)) != cudaSuccess) return e;
    // ---- linearise (skipped on the device unless phase == LINEARIZE): IMU edges beside the mono edges ----
    if ((e = cudaEventRecord(fork, s)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(side, fork, 0)) != cudaSuccess) return e;
    linearize_imu_v2_kernel<<<dim3(d.imu_grid, d.n_windows), 32, 0, side>>>(wp);
    if ((e = cudaEventRecord(join, side)) != cudaSuccess) return e;
    linearize_v2_kernel<<<dim3(d.point_grid, d.n_windows), d.lin_threads, d.smem_lin, s>>>(wp);
    if (probe && (e = cudaEventRecord(probe[7], s)) != cudaSuccess) return e;
    reduce_partials_kernel<<<dim3(d.reduce_grid, d.n_windows), 256, 0, s>>>(wp, d.point_grid);
    if ((e = cudaStreamWaitEvent(s, join, 0)) != cudaSuccess) return e;
    assemble_hpp_kernel<<<dim3(d.assemble_grid, d.n_windows), 256, 0, s>>>(wp);
    if (comm) {  // sharded window: only what computeLambdaInit needs crosses the wire (diag H_pp summed, max |diag H_ll|)
        if ((e = launch_shard_diag(s, wp, d)) != cudaSuccess) return e;
        if ((e = comm->reduce(comm->self, RED_DIAG, s)) != cudaSuccess) return e;
    }
    if (probe && (e = cudaEventRecord(probe[1], s)) != cudaSuccess) return e;
    if (comm && (e = launch_lm_iter_begin(s, wp, d)) != cudaSuccess) return e;  // (not sharded: tail of assemble_hpp)
    // ---- one LM trial (skipped unless phase == TRIAL) ----
    if (probe && (e = cudaEventRecord(probe[2], s)) != cudaSuccess) return e;
    if (d.sp_warps > 0) {
        schur_rec_kernel<<<dim3(4 * d.reduce_grid, d.n_windows), 256, 0, s>>>(wp, d.sp_tile_pts, d.sp_mma);
        if (d.sp_mma)
            schur_mma_kernel<<<dim3(d.sp_grid * d.sp_sets, d.n_windows), 32 * d.sp_warps, d.smem_sp, s>>>(wp, d.sp_sets, d.sp_tile_pts,
                                                                                                      d.sp_tile_edges);
        else if (d.sp_pair_lanes)
            schur_tile_pair_kernel<<<dim3(d.sp_grid, d.n_windows), 32 * d.sp_warps, d.smem_sp, s>>>(wp, d.sp_tile_pts);
        else
            schur_tile_kernel<<<dim3(d.sp_grid * d.sp_sets, d.n_windows), 32 * d.sp_warps, d.smem_sp, s>>>(wp, d.sp_sets, d.sp_tile_pts);
        if (probe && (e = cudaEventRecord(probe[6], s)) != cudaSuccess) return e;
        if (d.sp_pair_lanes)
            schur_finish_pair_kernel<<<dim3(d.assemble_grid, d.n_windows), 256, 0, s>>>(wp, d.sp_grid);
        else
            schur_finish_kernel<<<dim3(d.assemble_grid, d.n_windows), 256, 0, s>>>(wp, d.sp_grid);
    } else {
        schur_prep_kernel<<<dim3(d.point_grid, d.n_windows), 256, 0, s>>>(wp);
        if (probe && (e = cudaEventRecord(probe[6], s)) != cudaSuccess) return e;
        schur_gather_kernel<<<dim3(d.gather_grid, d.n_windows), kSchurThreads, 0, s>>>(wp);
    }
    if (comm && (e = comm->reduce(comm->self, RED_S, s)) != cudaSuccess) return e;  // S | b_s = sum of the partial reduced systems
    if (d.dbg_stop_after_schur) return cudaGetLastError();
    if (probe && (e = cudaEventRecord(probe[3], s)) != cudaSuccess) return e;
    if (d.chol_big_tiles > 0) e = launch_chol_big(s, side, fork, join, wp, d);
    else e = launch_chol_la(s, wp, d.n_windows, d.chol_cluster, d.chol_n);
    if (e != cudaSuccess) return e;
    if (probe && (e = cudaEventRecord(probe[4], s)) != cudaSuccess) return e;
    if ((e = launch_update_eval_apply(s, wp, d)) != cudaSuccess) return e;
    if (comm) {  // chi2 and the gain scale: landmark part from update_eval, pose part from this rank's partial b_p
        if ((e = launch_shard_scale(s, wp, d)) != cudaSuccess) return e;
        if ((e = comm->reduce(comm->self, RED_CHI, s)) != cudaSuccess) return e;
    }
    if (probe && (e = cudaEventRecord(probe[5], s)) != cudaSuccess) return e;
    // the LM decision: at the tail of update_eval, or (sharded) its own kernel behind the reduction of chi2 | scale
    if (comm && (e = launch_lm_decide(s, wp, d)) != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace vilba
