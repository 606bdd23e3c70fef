// lba_kernels.cu -- K2..K5 of the VI local-BA path for sm_100a (FP64 CUDA-core work).
//
//   update_eval : SparseOptimizer::update + computeActiveErrors + activeRobustChi2
//                 (g2o/core/sparse_optimizer.cpp:422-435,61-76,100-114) and the landmark
//                 back-substitution of BlockSolver::solve (g2o/core/block_solver.hpp:459-485)
//   linearize   : linearizeOplus + constructQuadraticForm of the three edge types
//                 (src/IMU/g2otypes.cpp:587-699,724-734,738-788; g2o/core/base_binary_edge.hpp:55-120;
//                 g2o/core/base_multi_edge.hpp:36-48,171-222) = BlockSolver::buildSystem (:502-560)
//   schur       : setLambda + the landmark loop of BlockSolver::solve (:564-589, :381-439)
//   chol_solve  : LinearSolverEigen::solve (g2o/solvers/linear_solver_eigen.h:94-124) as a dense
//                 Cholesky of the reduced camera system
//   lm_*        : OptimizationAlgorithmLevenberg::solve bookkeeping (optimization_algorithm_levenberg.cpp:61-164)
//   flags       : the cull / outlier loops of Optimizer::LocalBundleAdjustmentNavState (Optimizer.cpp:2659-2701)
//
// Work decomposition: one warp per map point (its mono edges sit on the lanes), one warp per IMU
// edge pair; every CTA first stages the key-frame states it needs (camera rotation R_cw = R_cb R_wb^T,
// P_wb, full NavState) in shared memory.
#include "lba_common.cuh"

namespace vilba {

// ------------------------------------------------------------------------------------------------
// update + evaluate
// ------------------------------------------------------------------------------------------------
template <bool APPLY>
__global__ void __launch_bounds__(kPointThreads) update_eval_kernel(DevWindow w) {
    extern __shared__ double smem[];
    const KfSmem ks = kf_smem_carve(smem, w.K);
    const int cur = w.lm->cur;
    const double lambda = w.lm->lambda;
    kf_stage<APPLY>(w, ks, cur);
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warps_per_cta = blockDim.x >> 5;
    const int gwarp = blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * warps_per_cta;
    const double* pts_in = w.pts[cur];
    double* pts_out = w.pts[cur ^ 1];
    double chi = 0.0, scale = 0.0;

    for (int item = gwarp; item < w.P + w.NI; item += nwarps) {
        if (item < w.P) {
            const int p = item;
            const int e0i = w.pt_obs_begin[p], e1i = w.pt_obs_begin[p + 1];
            V3 Pw = ld3(pts_in + 3 * (size_t)p);
            if (APPLY) {
                // x_l = D^-1 (b_l - sum_i W_i^T x_p(i)),  D = H_ll + lambda I   (block_solver.hpp:459-481)
                V3 acc = v3(0, 0, 0);
                for (int e = e0i + lane; e < e1i; e += 32) {
                    const int4 r = w.obs[e];
                    const int blk = ks.blk[r.w & OBS_KF_MASK];
                    if (!(r.w & OBS_CULLED) && blk >= 0) {
                        const double* Wp = w.W + 18 * (size_t)e;
                        const double* x = w.x + 15 * (size_t)blk;
                        const double xs[6] = {x[0], x[1], x[2], x[6], x[7], x[8]};
#pragma unroll
                        for (int rr = 0; rr < 6; ++rr) {
                            acc.x += Wp[3 * rr + 0] * xs[rr];
                            acc.y += Wp[3 * rr + 1] * xs[rr];
                            acc.z += Wp[3 * rr + 2] * xs[rr];
                        }
                    }
                }
                acc.x = warp_sum(acc.x), acc.y = warp_sum(acc.y), acc.z = warp_sum(acc.z);
                const double* H = w.Hll + 6 * (size_t)p;
                const V3 b = ld3(w.bl + 3 * (size_t)p);
                bool ok;
                const S3 Dinv = s3_inverse(S3{H[0] + lambda, H[1], H[2], H[3] + lambda, H[4], H[5] + lambda}, ok);
                const V3 xl = s3_mul(Dinv, b - acc);
                Pw = Pw + xl;  // VertexSBAPointXYZ::oplusImpl (types_sba.h:52-56)
                if (lane == 0) {
                    st3(pts_out + 3 * (size_t)p, Pw);
                    scale += xl.x * (lambda * xl.x + b.x) + xl.y * (lambda * xl.y + b.y) + xl.z * (lambda * xl.z + b.z);
                }
            }
            for (int e = e0i + lane; e < e1i; e += 32) {
                const MonoObs o = load_obs(w.obs, e);
                if (o.culled) continue;  // not in the active set: its cached error stays stale
                double r0, r1;
                V3 Paux, Pc;
                mono_error(w, ks.cam + 12 * o.kf, Pw, o, r0, r1, Paux, Pc);
                const double is2 = (double)o.is2;
                const double c2 = r0 * (is2 * r0) + r1 * (is2 * r1);
                w.obs_chi2[e] = c2;
                double rho0 = c2, rho1;
                if (o.robust) huber(c2, w.huber_mono, rho0, rho1);
                chi += rho0;
            }
        } else {
            // one IMU edge pair per warp; lane 0 evaluates both residuals
            const int e = item - w.P;
            if (lane == 0) {
                const double* si = ks.full + 22 * w.imu_i[e];
                const double* sj = ks.full + 22 * w.imu_j[e];
                const double* M = w.imu_preint + 142 * (size_t)e;
                V3 rP, rV, rPhi, rg, ra;
                pvr_error(w, si, sj, M, rP, rV, rPhi);
                const double ev[9] = {rP.x, rP.y, rP.z, rV.x, rV.y, rV.z, rPhi.x, rPhi.y, rPhi.z};
                double rho0, rho1;
                huber(quad9(w.imu_info + 81 * (size_t)e, ev), w.huber_pvr, rho0, rho1);
                chi += rho0;
                bias_error(si, sj, rg, ra);
                const double wg = w.inv_gyr_rw2 / M[VILBA_PI_DT], wa = w.inv_acc_rw2 / M[VILBA_PI_DT];
                const double c2 = rg.x * (wg * rg.x) + rg.y * (wg * rg.y) + rg.z * (wg * rg.z) +
                                  ra.x * (wa * ra.x) + ra.y * (wa * ra.y) + ra.z * (wa * ra.z);
                huber(c2, w.huber_bias, rho0, rho1);
                chi += rho0;
            }
        }
    }
    // CTA reduction, then one atomic per CTA
    __shared__ double red[2][kPointThreads / 32];
    chi = warp_sum(chi);
    scale = warp_sum(scale);
    if (lane == 0) {
        red[0][threadIdx.x >> 5] = chi;
        red[1][threadIdx.x >> 5] = scale;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0.0, s = 0.0;
        for (int i = 0; i < warps_per_cta; ++i) c += red[0][i], s += red[1][i];
        atomicAdd(&w.lm->chi_acc, c);
        if (APPLY) atomicAdd(&w.lm->scale_acc, s);
    }
}

// ------------------------------------------------------------------------------------------------
// linearize + accumulate: mono edges (warp per point)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPointThreads) linearize_mono_kernel(DevWindow w) {
    extern __shared__ double smem[];
    const KfSmem ks = kf_smem_carve(smem, w.K);
    const int cur = w.lm->cur;
    kf_stage<false>(w, ks, cur);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_per_cta = blockDim.x >> 5;
    const int gwarp = blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * warps_per_cta;
    const double* pts = w.pts[cur];
    const M3 Rcb = ldm3(w.Rcb);
    const int n = w.n;
    double maxd = 0.0;

    for (int p = gwarp; p < w.P; p += nwarps) {
        const int e0i = w.pt_obs_begin[p], e1i = w.pt_obs_begin[p + 1];
        const V3 Pw = ld3(pts + 3 * (size_t)p);
        double hxx = 0, hxy = 0, hxz = 0, hyy = 0, hyz = 0, hzz = 0, bx = 0, by = 0, bz = 0;
        for (int e = e0i + lane; e < e1i; e += 32) {
            const MonoObs o = load_obs(w.obs, e);
            double* Wp = w.W + 18 * (size_t)e;
            if (o.culled) {
#pragma unroll
                for (int i = 0; i < 18; ++i) Wp[i] = 0.0;
                continue;
            }
            const double* cam = ks.cam + 12 * o.kf;
            double r0, r1;
            V3 Paux, Pc;
            mono_error(w, cam, Pw, o, r0, r1, Paux, Pc);
            const double is2 = (double)o.is2;
            double wgt = is2;
            if (o.robust) {
                double rho0, rho1;
                huber(r0 * (is2 * r0) + r1 * (is2 * r1), w.huber_mono, rho0, rho1);
                wgt = rho1 * is2;  // robustInformation (base_edge.h:96-102)
            }
            // linearizeOplus (g2otypes.cpp:738-788)
            const M3 Rcw = ldm3(cam);
            const double z = Pc.z;
            const double ja = w.fx / z, jb = (-Pc.x / z * w.fx) / z;
            const double jc = w.fy / z, jd = (-Pc.y / z * w.fy) / z;
            // J_point = -Jpi * Rcw ; J_P = -J_point
            const double l00 = -(ja * Rcw.a00 + jb * Rcw.a20), l01 = -(ja * Rcw.a01 + jb * Rcw.a21),
                         l02 = -(ja * Rcw.a02 + jb * Rcw.a22);
            const double l10 = -(jc * Rcw.a10 + jd * Rcw.a20), l11 = -(jc * Rcw.a11 + jd * Rcw.a21),
                         l12 = -(jc * Rcw.a12 + jd * Rcw.a22);
            // J_Phi = -Jpi * (hat(Paux) * Rcb)
            const M3 HR = hat(Paux) * Rcb;
            const double f00 = -(ja * HR.a00 + jb * HR.a20), f01 = -(ja * HR.a01 + jb * HR.a21),
                         f02 = -(ja * HR.a02 + jb * HR.a22);
            const double f10 = -(jc * HR.a10 + jd * HR.a20), f11 = -(jc * HR.a11 + jd * HR.a21),
                         f12 = -(jc * HR.a12 + jd * HR.a22);
            const double Jl[2][3] = {{l00, l01, l02}, {l10, l11, l12}};
            const double Jp[2][6] = {{-l00, -l01, -l02, f00, f01, f02}, {-l10, -l11, -l12, f10, f11, f12}};
            // landmark block and rhs
            hxx += wgt * (l00 * l00 + l10 * l10);
            hxy += wgt * (l00 * l01 + l10 * l11);
            hxz += wgt * (l00 * l02 + l10 * l12);
            hyy += wgt * (l01 * l01 + l11 * l11);
            hyz += wgt * (l01 * l02 + l11 * l12);
            hzz += wgt * (l02 * l02 + l12 * l12);
            const double wr0 = -wgt * r0, wr1 = -wgt * r1;  // omega_r = -rho1 * Omega * e
            bx += l00 * wr0 + l10 * wr1;
            by += l01 * wr0 + l11 * wr1;
            bz += l02 * wr0 + l12 * wr1;
            const int blk = ks.blk[o.kf];
            if (blk < 0) {  // fixed key-frame: no pose block, no H_pl
#pragma unroll
                for (int i = 0; i < 18; ++i) Wp[i] = 0.0;
                continue;
            }
            // H_pl block (6x3): W = Jp^T (wgt) Jl
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) Wp[3 * r + c] = wgt * (Jp[0][r] * Jl[0][c] + Jp[1][r] * Jl[1][c]);
            // pose diagonal block (upper triangle of the 6x6 at rows/cols [P, Phi]) and rhs
            const int off = 15 * blk;
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                const int gr = off + (r < 3 ? r : r + 3);
#pragma unroll
                for (int c = r; c < 6; ++c) {
                    const int gc = off + (c < 3 ? c : c + 3);
                    atomicAdd(&w.Hpp[(size_t)gr * n + gc], wgt * (Jp[0][r] * Jp[0][c] + Jp[1][r] * Jp[1][c]));
                }
                atomicAdd(&w.bp[gr], Jp[0][r] * wr0 + Jp[1][r] * wr1);
            }
        }
        hxx = warp_sum(hxx), hxy = warp_sum(hxy), hxz = warp_sum(hxz);
        hyy = warp_sum(hyy), hyz = warp_sum(hyz), hzz = warp_sum(hzz);
        bx = warp_sum(bx), by = warp_sum(by), bz = warp_sum(bz);
        if (lane == 0) {
            double* H = w.Hll + 6 * (size_t)p;
            H[0] = hxx, H[1] = hxy, H[2] = hxz, H[3] = hyy, H[4] = hyz, H[5] = hzz;
            st3(w.bl + 3 * (size_t)p, v3(bx, by, bz));
            maxd = fmax(maxd, fmax(fabs(hxx), fmax(fabs(hyy), fabs(hzz))));
        }
    }
    maxd = warp_max(maxd);
    if (lane == 0 && maxd > 0.0) atomic_max_nonneg(&w.lm->maxdiag_bits, maxd);
}

// ------------------------------------------------------------------------------------------------
// linearize + accumulate: IMU edges (one warp per EdgeNavStatePVR + EdgeNavStateBias pair)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) linearize_imu_kernel(DevWindow w) {
    __shared__ double J[9 * 24];   // columns: PVR_i (9) | Bias_i (6) | PVR_j (9)
    __shared__ double Om[81];
    __shared__ double TJ[9 * 24];  // Omega_w * J
    __shared__ double ev[9];
    __shared__ double Oe[9];       // Omega * e
    __shared__ double wgt_s;
    const int e = blockIdx.x;
    const int lane = threadIdx.x;
    const int cur = w.lm->cur;
    const int ki = w.imu_i[e], kj = w.imu_j[e];
    const double* si = w.kf_state[cur] + 22 * (size_t)ki;
    const double* sj = w.kf_state[cur] + 22 * (size_t)kj;
    const double* M = w.imu_preint + 142 * (size_t)e;
    const int bi = w.kf_block[ki], bj = w.kf_block[kj];
    const int n = w.n;

    for (int i = lane; i < 81; i += 32) Om[i] = w.imu_info[81 * (size_t)e + i];
    for (int i = lane; i < 9 * 24; i += 32) J[i] = 0.0;
    __syncwarp();
    if (lane == 0) {
        V3 rP, rV, rPhi;
        pvr_error(w, si, sj, M, rP, rV, rPhi);  // the cached _error of computeActiveErrors (same state)
        ev[0] = rP.x, ev[1] = rP.y, ev[2] = rP.z, ev[3] = rV.x, ev[4] = rV.y, ev[5] = rV.z;
        ev[6] = rPhi.x, ev[7] = rPhi.y, ev[8] = rPhi.z;
        double rho0, rho1;
        huber(quad9(Om, ev), w.huber_pvr, rho0, rho1);
        wgt_s = rho1;
        for (int r = 0; r < 9; ++r) {
            double t = 0.0;
            for (int c = 0; c < 9; ++c) t += Om[9 * r + c] * ev[c];
            Oe[r] = t;
        }
        // EdgeNavStatePVR::linearizeOplus (g2otypes.cpp:587-699)
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
        // vertex 0: PVR_i
        put(0, 0, -RiT);
        put(0, 3, RiT * (-T));
        put(0, 6, hat(RiT * (Pj - Pi - Vi * T - (0.5 * g) * T2)));
        put(3, 3, -RiT);
        put(3, 6, hat(RiT * (Vj - Vi - g * T)));
        put(6, 6, ((-JrInv) * transpose(Rj)) * Ri);
        // vertex 2: Bias_i (dbg, dba)
        const M3 ExpT = q_to_matrix(so3_inverse(so3_exp(rPhi)));
        const M3 JrCorr = jacobian_r(JRg * dbg);
        put(0, 9, -ldm3(M + VILBA_PI_JPG));
        put(0, 12, -ldm3(M + VILBA_PI_JPA));
        put(3, 9, -ldm3(M + VILBA_PI_JVG));
        put(3, 12, -ldm3(M + VILBA_PI_JVA));
        put(6, 9, (((-JrInv) * ExpT) * JrCorr) * JRg);
        // vertex 1: PVR_j
        put(0, 15, RiT);
        put(3, 18, RiT);
        put(6, 21, JrInv);
    }
    __syncwarp();
    const double wgt = wgt_s;
    // TJ = (rho1 * Omega) * J
    for (int i = lane; i < 9 * 24; i += 32) {
        const int r = i / 24, c = i - 24 * r;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 9; ++k) s += Om[9 * r + k] * J[24 * k + c];
        TJ[i] = wgt * s;
    }
    __syncwarp();
    // H = J^T TJ, scattered into the upper triangle of Hpp; b = -J^T (rho1 Omega e) = -TJ^T e
    for (int i = lane; i < 24 * 24; i += 32) {
        const int r = i / 24, c = i - 24 * r;
        const int br = (r < 15) ? bi : bj, bc = (c < 15) ? bi : bj;
        if (br < 0 || bc < 0) continue;
        const int gr = 15 * br + (r < 15 ? r : r - 15), gc = 15 * bc + (c < 15 ? c : c - 15);
        if (gr > gc) continue;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 9; ++k) s += J[24 * k + r] * TJ[24 * k + c];
        atomicAdd(&w.Hpp[(size_t)gr * n + gc], s);
    }
    if (lane < 24) {
        const int r = lane;
        const int br = (r < 15) ? bi : bj;
        if (br >= 0) {
            double s = 0.0;  // A^T * (-rho1 * Omega e)
#pragma unroll
            for (int k = 0; k < 9; ++k) s += J[24 * k + r] * (wgt * Oe[k]);
            atomicAdd(&w.bp[15 * br + (r < 15 ? r : r - 15)], -s);
        }
    }
    // EdgeNavStateBias: A = -I, B = +I (g2otypes.cpp:724-734), information diag(1/rw2)/dT
    if (lane < 6) {
        V3 rg, ra;
        bias_error(si, sj, rg, ra);
        const double wg = w.inv_gyr_rw2 / M[VILBA_PI_DT], wa = w.inv_acc_rw2 / M[VILBA_PI_DT];
        const double c2 = rg.x * (wg * rg.x) + rg.y * (wg * rg.y) + rg.z * (wg * rg.z) + ra.x * (wa * ra.x) +
                          ra.y * (wa * ra.y) + ra.z * (wa * ra.z);
        double rho0, rho1;
        huber(c2, w.huber_bias, rho0, rho1);
        const double eb[6] = {rg.x, rg.y, rg.z, ra.x, ra.y, ra.z};
        const double om = rho1 * (lane < 3 ? wg : wa);
        const double omega_r = -om * eb[lane];
        if (bi >= 0) {
            const int gi = 15 * bi + 9 + lane;
            atomicAdd(&w.Hpp[(size_t)gi * n + gi], om);
            atomicAdd(&w.bp[gi], -omega_r);  // A^T omega_r, A = -I
        }
        if (bj >= 0) {
            const int gj = 15 * bj + 9 + lane;
            atomicAdd(&w.Hpp[(size_t)gj * n + gj], om);
            atomicAdd(&w.bp[gj], omega_r);
        }
        if (bi >= 0 && bj >= 0) {
            const int gi = 15 * bi + 9 + lane, gj = 15 * bj + 9 + lane;
            const int r = min(gi, gj), c = max(gi, gj);
            atomicAdd(&w.Hpp[(size_t)r * n + c], -om);  // A^T Omega B
        }
    }
}

// ------------------------------------------------------------------------------------------------
// imu_prepare: information of EdgeNavStatePVR = inverse of cov_P_V_Phi (Optimizer.cpp:2510).
// Gauss-Jordan with partial pivoting, one warp per edge (lane = column of the augmented matrix).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) imu_prepare_kernel(DevWindow w) {
    __shared__ double A[9][18];
    const int e = blockIdx.x, lane = threadIdx.x;
    const double* cov = w.imu_preint + 142 * (size_t)e + VILBA_PI_COV;
    for (int i = lane; i < 9 * 18; i += 32) {
        const int r = i / 18, c = i - 18 * r;
        A[r][c] = (c < 9) ? cov[9 * r + c] : ((c - 9 == r) ? 1.0 : 0.0);
    }
    __syncwarp();
    for (int k = 0; k < 9; ++k) {
        int piv = k;
        double best = fabs(A[k][k]);
        for (int r = k + 1; r < 9; ++r) {
            const double v = fabs(A[r][k]);
            if (v > best) best = v, piv = r;
        }
        __syncwarp();
        if (piv != k && lane < 18) {
            const double t = A[k][lane];
            A[k][lane] = A[piv][lane];
            A[piv][lane] = t;
        }
        __syncwarp();
        const double d = A[k][k];
        __syncwarp();
        if (lane < 18) A[k][lane] /= d;
        __syncwarp();
        double f[9];
#pragma unroll
        for (int r = 0; r < 9; ++r) f[r] = A[r][k];
        __syncwarp();
        if (lane < 18) {
            const double akc = A[k][lane];
#pragma unroll
            for (int r = 0; r < 9; ++r)
                if (r != k) A[r][lane] -= f[r] * akc;
        }
        __syncwarp();
    }
    for (int i = lane; i < 81; i += 32) w.imu_info[81 * (size_t)e + i] = A[i / 9][9 + i % 9];
}

// ------------------------------------------------------------------------------------------------
// Schur complement: S = Hpp + lambda I - sum_l W_l D_l^-1 W_l^T,  bs = bp - sum_l W_l D_l^-1 b_l
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) schur_init_kernel(DevWindow w) {
    const double lambda = w.lm->lambda;
    const int n = w.n;
    const size_t total = (size_t)n * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / n), c = (int)(i - (size_t)r * n);
        double v = w.Hpp[i];
        if (r == c) v += lambda;  // setLambda on the pose blocks (block_solver.hpp:570-577)
        w.S[i] = v;
        if (c == 0) w.bs[r] = w.bp[r];
    }
}

__global__ void __launch_bounds__(kPointThreads) schur_points_kernel(DevWindow w) {
    const double lambda = w.lm->lambda;
    const int lane = threadIdx.x & 31;
    const int warps_per_cta = blockDim.x >> 5;
    const int gwarp = blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * warps_per_cta;
    const int n = w.n;
    for (int p = gwarp; p < w.P; p += nwarps) {
        const int e0i = w.pt_obs_begin[p], e1i = w.pt_obs_begin[p + 1];
        const int m = e1i - e0i;
        const double* H = w.Hll + 6 * (size_t)p;
        bool ok;
        const S3 Dinv = s3_inverse(S3{H[0] + lambda, H[1], H[2], H[3] + lambda, H[4], H[5] + lambda}, ok);
        const V3 db = s3_mul(Dinv, ld3(w.bl + 3 * (size_t)p));
        for (int base = 0; base < m; base += 32) {  // chunks of 32 observations (i side)
            const int ei = e0i + base + lane;
            double Wi[18], Yi[18];
            int bi = -1;
            if (base + lane < m) {
                const int4 r = w.obs[ei];
                if (!(r.w & OBS_CULLED)) bi = w.kf_block[r.w & OBS_KF_MASK];
            }
            if (bi >= 0) {
#pragma unroll
                for (int i = 0; i < 18; ++i) Wi[i] = w.W[18 * (size_t)ei + i];
            } else {
#pragma unroll
                for (int i = 0; i < 18; ++i) Wi[i] = 0.0;
            }
#pragma unroll
            for (int r = 0; r < 6; ++r) {  // Y = W * Dinv  (BDinv, block_solver.hpp:407)
                const V3 y = s3_mul(Dinv, v3(Wi[3 * r], Wi[3 * r + 1], Wi[3 * r + 2]));
                Yi[3 * r] = y.x, Yi[3 * r + 1] = y.y, Yi[3 * r + 2] = y.z;
            }
            if (bi >= 0) {  // bs -= W * (Dinv b_l)   (:409-413)
                const int off = 15 * bi;
#pragma unroll
                for (int r = 0; r < 6; ++r)
                    atomicAdd(&w.bs[off + (r < 3 ? r : r + 3)],
                              -(Wi[3 * r] * db.x + Wi[3 * r + 1] * db.y + Wi[3 * r + 2] * db.z));
            }
            // pairs (i, j), j >= i over the whole observation list of the point
            for (int jb = base; jb < m; jb += 32) {
                const int ej = e0i + jb + lane;
                double Wj[18];
                int bjl = -1;
                if (jb == base) {
                    bjl = bi;
#pragma unroll
                    for (int i = 0; i < 18; ++i) Wj[i] = Wi[i];
                } else {
                    if (jb + lane < m) {
                        const int4 r = w.obs[ej];
                        if (!(r.w & OBS_CULLED)) bjl = w.kf_block[r.w & OBS_KF_MASK];
                    }
#pragma unroll
                    for (int i = 0; i < 18; ++i) Wj[i] = (bjl >= 0) ? w.W[18 * (size_t)ej + i] : 0.0;
                }
                const int cntj = min(32, m - jb);
                for (int j = 0; j < cntj; ++j) {
                    const int bj = __shfl_sync(0xffffffffu, bjl, j);
                    double Wb[18];
#pragma unroll
                    for (int i = 0; i < 18; ++i) Wb[i] = __shfl_sync(0xffffffffu, Wj[i], j);
                    if (bj < 0 || bi < 0) continue;
                    const int gi_obs = base + lane, gj_obs = jb + j;
                    if (gi_obs > gj_obs) continue;
                    const bool diag = (gi_obs == gj_obs);
                    const int offi = 15 * bi, offj = 15 * bj;
#pragma unroll
                    for (int r = 0; r < 6; ++r) {
                        const int gr = offi + (r < 3 ? r : r + 3);
#pragma unroll
                        for (int c = 0; c < 6; ++c) {
                            if (diag && c < r) continue;
                            const int gc = offj + (c < 3 ? c : c + 3);
                            const double v =
                                Yi[3 * r] * Wb[3 * c] + Yi[3 * r + 1] * Wb[3 * c + 1] + Yi[3 * r + 2] * Wb[3 * c + 2];
                            const size_t idx = (gr <= gc) ? (size_t)gr * n + gc : (size_t)gc * n + gr;
                            atomicAdd(&w.S[idx], -v);  // Hschur(i1,i2) -= BDinv * Bj^T (:416-430)
                        }
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// dense Cholesky + solve of the reduced camera system (single CTA, blocked right-looking).
// The upper triangle of the row-major S is read as the lower triangle of a column-major matrix:
// L(i,j), i >= j, lives at S[j*n + i].  The right-hand side rides along as an extra row, so the
// forward substitution is a by-product of the panel solves; a single warp then back-substitutes.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) chol_solve_kernel(DevWindow w) {
    constexpr int NB = kCholNB;
    extern __shared__ double smem[];
    const int n = w.n;
    double* D = smem;                 // NB x NB diagonal block (lower, row = i)
    double* Pn = smem + NB * NB;      // (rows below + rhs row) x NB panel
    double* A = w.S;
    double* y = w.bs;                 // becomes L^-1 b
    __shared__ int fail;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) fail = 0;
    __syncthreads();

    for (int j0 = 0; j0 < n; j0 += NB) {
        const int jb = min(NB, n - j0);
        // 1. diagonal block -> shared, factor with one warp
        for (int i = tid; i < NB * NB; i += nt) {
            const int r = i / NB, c = i - NB * r;
            D[i] = (r < jb && c < jb && c <= r) ? A[(size_t)(j0 + c) * n + j0 + r] : 0.0;
        }
        __syncthreads();
        if (tid < 32) {
            for (int j = 0; j < jb; ++j) {
                const double djj = D[j * NB + j];
                if (!(djj > 0.0)) {
                    if (tid == 0) fail = 1;
                }
                const double d = sqrt(djj);
                __syncwarp();
                if (tid > j && tid < jb) D[tid * NB + j] /= d;
                if (tid == j) D[j * NB + j] = d;
                __syncwarp();
                // rank-1 update of the remaining lower triangle
                for (int idx = tid; idx < (jb - j - 1) * (jb - j - 1); idx += 32) {
                    const int r = j + 1 + idx / (jb - j - 1), c = j + 1 + idx % (jb - j - 1);
                    if (c <= r) D[r * NB + c] -= D[r * NB + j] * D[c * NB + j];
                }
                __syncwarp();
            }
        }
        __syncthreads();
        // write back the factored diagonal block
        for (int i = tid; i < NB * NB; i += nt) {
            const int r = i / NB, c = i - NB * r;
            if (r < jb && c <= r) A[(size_t)(j0 + c) * n + j0 + r] = D[i];
        }
        // 2. panel: rows below the block and the rhs row (index m_rows-1): X L11^T = A21
        const int rows_below = n - j0 - jb;
        const int m_rows = rows_below + 1;
        for (int rr = tid; rr < m_rows; rr += nt) {
            double xr[NB];
            const bool is_rhs = (rr == rows_below);
            const int gi = j0 + jb + rr;
#pragma unroll
            for (int c = 0; c < NB; ++c)
                xr[c] = (c < jb) ? (is_rhs ? y[j0 + c] : A[(size_t)(j0 + c) * n + gi]) : 0.0;
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                if (c < jb) {
                    double s = xr[c];
#pragma unroll
                    for (int k = 0; k < NB; ++k)
                        if (k < c) s -= xr[k] * D[c * NB + k];
                    xr[c] = s / D[c * NB + c];
                }
            }
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                Pn[(size_t)rr * NB + c] = xr[c];
                if (c < jb) {
                    if (is_rhs)
                        y[j0 + c] = xr[c];
                    else
                        A[(size_t)(j0 + c) * n + gi] = xr[c];
                }
            }
        }
        __syncthreads();
        // 3. trailing update: A22(i,c) -= P(i,:) . P(c,:) for i >= c, and rhs(c) -= P(rhs,:) . P(c,:)
        //    4x4 register tiles over the (rows_below+1) x rows_below lower-trapezoid
        {
            const int tr = (m_rows + 3) / 4, tc = (rows_below + 3) / 4;
            for (int t = tid; t < tr * tc; t += nt) {
                const int ti = t / tc, tj = t - tc * ti;
                if (tj > ti) continue;
                const int r0 = 4 * ti, c0 = 4 * tj;
                double acc[4][4];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll 4
                for (int k = 0; k < NB; ++k) {
                    double pr[4], pc[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        pr[a] = (r0 + a < m_rows) ? Pn[(size_t)(r0 + a) * NB + k] : 0.0;
                        pc[a] = (c0 + a < rows_below) ? Pn[(size_t)(c0 + a) * NB + k] : 0.0;
                    }
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) acc[a][b] += pr[a] * pc[b];
                }
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int r = r0 + a, c = c0 + b;
                        if (r >= m_rows || c >= rows_below || c > r) continue;
                        if (r == rows_below)
                            y[j0 + jb + c] -= acc[a][b];
                        else
                            A[(size_t)(j0 + jb + c) * n + (j0 + jb + r)] -= acc[a][b];
                    }
            }
        }
        __syncthreads();
    }
    // back substitution L^T x = y with one warp: x_i = (y_i - sum_{k>i} L(k,i) x_k) / L(i,i)
    if (tid < 32) {
        double* x = w.x;
        for (int i = tid; i < n; i += 32) x[i] = y[i];
        __syncwarp();
        for (int i = n - 1; i >= 0; --i) {
            const double xi = x[i] / A[(size_t)i * n + i];
            __syncwarp();
            if (tid == 0) x[i] = xi;
            // x[c] -= L(i,c) * xi for c < i ;  L(i,c) at A[c*n + i]
            for (int c = tid; c < i; c += 32) x[c] -= A[(size_t)c * n + i] * xi;
            __syncwarp();
        }
        if (tid == 0) w.lm->chol_fail = fail;
    }
}

// ------------------------------------------------------------------------------------------------
// LM bookkeeping (single CTA each)
// ------------------------------------------------------------------------------------------------
__global__ void lm_stage_begin_kernel(DevWindow w) {
    if (threadIdx.x == 0) {
        LmState* s = w.lm;
        s->current_chi = s->chi_acc;  // currentChi = activeRobustChi2() at the stage's first iteration
        s->chi_acc = 0.0;
        s->scale_acc = 0.0;
        s->maxdiag_bits = 0ull;
    }
}

__global__ void __launch_bounds__(256) lm_iter_begin_kernel(DevWindow w, int iteration) {
    __shared__ double red[8];
    LmState* s = w.lm;
    double m = 0.0;
    if (iteration == 0)
        for (int d = threadIdx.x; d < w.n; d += blockDim.x) m = fmax(m, fabs(w.Hpp[(size_t)d * w.n + d]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        if (iteration == 0) {  // computeLambdaInit (optimization_algorithm_levenberg.cpp:166-180)
            double mx = __longlong_as_double((long long)s->maxdiag_bits);
            for (int i = 0; i < (int)(blockDim.x >> 5); ++i) mx = fmax(mx, red[i]);
            s->lambda = w.lm_tau * mx;
            s->ni = 2.0;
            s->n_bad = 0;
        }
        s->ini_chi = s->current_chi;
        s->lambda_first = s->lambda;
        s->qmax = 0;
        s->iter_result = -1;
        s->maxdiag_bits = 0ull;
        s->chi_acc = 0.0;
        s->scale_acc = 0.0;
    }
}

__global__ void __launch_bounds__(32) lm_decide_kernel(DevWindow w) {
    LmState* s = w.lm;
    const double lambda = s->lambda;
    double sc = 0.0;  // computeScale over the pose part (:182-189); landmarks arrive in scale_acc
    for (int j = threadIdx.x; j < w.n; j += 32) sc += w.x[j] * (lambda * w.x[j] + w.bp[j]);
    sc = warp_sum(sc);
    if (threadIdx.x == 0) {
        double scale = sc + s->scale_acc;
        scale += 1e-3;
        double tempChi = s->chi_acc;
        if (s->chol_fail) tempChi = DBL_MAX;
        double rho = (s->current_chi - tempChi) / scale;
        s->temp_chi = tempChi;
        if (rho > 0 && isfinite(tempChi)) {  // :134-142
            double alpha = 1. - pow((2 * rho - 1), 3.0);
            alpha = fmin(alpha, w.lm_good_hi);
            const double scaleFactor = fmax(w.lm_good_lo, alpha);
            s->lambda = lambda * scaleFactor;
            s->ni = 2;
            s->current_chi = tempChi;
            s->cur ^= 1;  // discardTop: the trial buffer becomes the estimate
            s->accepted = 1;
        } else {  // :143-147  pop: the estimate buffer is untouched, cached errors stay stale
            s->lambda = lambda * s->ni;
            s->ni *= 2;
            s->accepted = 0;
        }
        s->rho = rho;
        s->qmax += 1;
        s->chi_acc = 0.0;
        s->scale_acc = 0.0;
        s->chol_fail = 0;
        const bool again = (rho < 0) && (s->qmax < w.max_trials) && !s->stop;
        if (!again) {
            int res = 0;
            if (s->qmax == w.max_trials || rho == 0)
                res = 1;
            else {
                if ((s->ini_chi - s->current_chi) * 1e3 < s->ini_chi)
                    s->n_bad++;
                else
                    s->n_bad = 0;
                if (s->n_bad >= 3) res = 1;
            }
            s->iter_result = res;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// cull (after stage 1) and final outlier flags: chi2 from the cached errors, depth from the estimates
// ------------------------------------------------------------------------------------------------
template <bool CULL>
__global__ void __launch_bounds__(kPointThreads) flags_kernel(DevWindow w, uint8_t* outlier, int* n_culled) {
    extern __shared__ double smem[];
    const KfSmem ks = kf_smem_carve(smem, w.K);
    const int cur = w.lm->cur;
    kf_stage<false>(w, ks, cur);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_per_cta = blockDim.x >> 5;
    const int gwarp = blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * warps_per_cta;
    int cnt = 0;
    for (int p = gwarp; p < w.P; p += nwarps) {
        const V3 Pw = ld3(w.pts[cur] + 3 * (size_t)p);
        for (int e = w.pt_obs_begin[p] + lane; e < w.pt_obs_begin[p + 1]; e += 32) {
            int4 r = w.obs[e];
            const MonoObs o = load_obs(w.obs, e);
            double r0, r1;
            V3 Paux, Pc;
            mono_error(w, ks.cam + 12 * o.kf, Pw, o, r0, r1, Paux, Pc);
            const bool bad = (w.obs_chi2[e] > w.chi2_gate) || !(Pc.z > 0.0);  // isDepthPositive
            if (CULL) {
                if (bad) {
                    r.w |= OBS_CULLED;
                    ++cnt;
                }
                r.w &= ~OBS_ROBUST;  // e->setRobustKernel(0) on every mono edge
                w.obs[e] = r;
            } else {
                outlier[e] = bad ? 1 : 0;
            }
        }
    }
    if (CULL) {
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0 && cnt) atomicAdd(n_culled, cnt);
    }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
cudaError_t launch_imu_prepare(cudaStream_t s, const DevWindow& w) {
    if (w.NI > 0) imu_prepare_kernel<<<w.NI, 32, 0, s>>>(w);
    return cudaGetLastError();
}

cudaError_t launch_update_eval(cudaStream_t s, const DevWindow& w, const LaunchCfg& cfg, bool apply) {
    const size_t sm = kf_smem_bytes(w.K);
    if (apply)
        update_eval_kernel<true><<<cfg.point_grid, kPointThreads, sm, s>>>(w);
    else
        update_eval_kernel<false><<<cfg.point_grid, kPointThreads, sm, s>>>(w);
    return cudaGetLastError();
}

cudaError_t launch_linearize(cudaStream_t s, const DevWindow& w, const LaunchCfg& cfg) {
    cudaError_t e = cudaMemsetAsync(w.Hpp, 0, sizeof(double) * (size_t)w.n * w.n, s);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(w.bp, 0, sizeof(double) * (size_t)w.n, s);
    if (e != cudaSuccess) return e;
    const size_t sm = kf_smem_bytes(w.K);
    linearize_mono_kernel<<<cfg.point_grid, kPointThreads, sm, s>>>(w);
    if (w.NI > 0) linearize_imu_kernel<<<w.NI, 32, 0, s>>>(w);
    return cudaGetLastError();
}

cudaError_t launch_schur(cudaStream_t s, const DevWindow& w, const LaunchCfg& cfg) {
    const size_t total = (size_t)w.n * w.n;
    size_t g = (total + 255) / 256;
    if (g > (size_t)(4 * cfg.sm_count)) g = (size_t)(4 * cfg.sm_count);
    if (g < 1) g = 1;
    schur_init_kernel<<<(int)g, 256, 0, s>>>(w);
    schur_points_kernel<<<cfg.point_grid, kPointThreads, 0, s>>>(w);
    return cudaGetLastError();
}

static size_t chol_smem_bytes(int n) { return sizeof(double) * ((size_t)kCholNB * kCholNB + (size_t)(n + 1) * kCholNB); }

cudaError_t launch_chol_solve(cudaStream_t s, const DevWindow& w) {
    const size_t sm = chol_smem_bytes(w.n);
    static size_t configured = 0;
    if (sm > configured) {
        cudaError_t e = cudaFuncSetAttribute(chol_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
        configured = sm;
    }
    chol_solve_kernel<<<1, 1024, sm, s>>>(w);
    return cudaGetLastError();
}

cudaError_t launch_lm_stage_begin(cudaStream_t s, const DevWindow& w) {
    lm_stage_begin_kernel<<<1, 32, 0, s>>>(w);
    return cudaGetLastError();
}
cudaError_t launch_lm_iter_begin(cudaStream_t s, const DevWindow& w, int iteration) {
    lm_iter_begin_kernel<<<1, 256, 0, s>>>(w, iteration);
    return cudaGetLastError();
}
cudaError_t launch_lm_decide(cudaStream_t s, const DevWindow& w) {
    lm_decide_kernel<<<1, 32, 0, s>>>(w);
    return cudaGetLastError();
}
cudaError_t launch_cull(cudaStream_t s, const DevWindow& w, const LaunchCfg& cfg, int* n_culled) {
    flags_kernel<true><<<cfg.point_grid, kPointThreads, kf_smem_bytes(w.K), s>>>(w, nullptr, n_culled);
    return cudaGetLastError();
}
cudaError_t launch_final_flags(cudaStream_t s, const DevWindow& w, const LaunchCfg& cfg, uint8_t* outlier) {
    flags_kernel<false><<<cfg.point_grid, kPointThreads, kf_smem_bytes(w.K), s>>>(w, outlier, nullptr);
    return cudaGetLastError();
}

}  // namespace vilba
