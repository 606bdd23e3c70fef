// lba_kernels.cu -- per-point / per-edge kernels and the device-side LM controller of the VI local-BA
// path for sm_100a (FP64 CUDA-core work).
//
//   update_eval : SparseOptimizer::update + computeActiveErrors + activeRobustChi2
//                 (g2o/core/sparse_optimizer.cpp:422-435,61-76,100-114) and the landmark
//                 back-substitution of BlockSolver::solve (g2o/core/block_solver.hpp:459-485)
//   imu_prepare : information matrix of EdgeNavStatePVR (src/Optimizer.cpp:2510)
//   lm_*        : OptimizationAlgorithmLevenberg::solve bookkeeping (optimization_algorithm_levenberg.cpp:61-164)
//                 and the iteration loop of SparseOptimizer::optimize (sparse_optimizer.cpp:376-414), kept in
//                 device memory so that the host never synchronises inside an optimize() call.  The accept / reject
//                 decision runs in the last CTA of update_eval (lm_decide_warp), the start of an iteration in the last
//                 CTA of assemble_hpp; a sharded window keeps both as kernels behind its reductions
//   flags       : the cull / outlier loops of Optimizer::LocalBundleAdjustmentNavState (Optimizer.cpp:2659-2701)
// The accumulation kernels (linearize / assemble / Schur) live in lba_v2.cu, the reduced-system LDL^T in chol_la.cu / chol_big.cu.
//
// Work decomposition: eight lanes per map point (its mono edges sit on the lanes), one lane group per
// IMU edge pair; every CTA first stages the key-frame states it needs (camera rotation
// R_cw = R_cb R_wb^T, P_wb, full NavState) in shared memory.  Every kernel receives a pointer to the
// device-resident DevWindow and is launched with a fixed grid, so an LM slot is graph-capturable.
#include "lba_common.cuh"

namespace vilba {

// Eight lanes per map point (four points per warp): the typical point has ~8 observations, so a
// full warp per point would leave 3/4 of the lanes idle.  Points with more observations loop.
__device__ __forceinline__ double group8_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// The LM decision of one trial (OptimizationAlgorithmLevenberg::solve, :120-160, and the loop condition of optimize(),
// sparse_optimizer.cpp:376): one warp.  Runs at the tail of update_eval (the last CTA to arrive has chi2 and the landmark
// part of the gain scale in hand), or as its own kernel behind the allreduce when the window is sharded.
__device__ __forceinline__ void lm_decide_warp(const DevWindow& w, LmState* s, int lane) {
    const double lambda = s->lambda;
    double sc = 0.0;  // computeScale over the pose part (:182-189); landmarks arrive in scale_acc
    if (!w.sharded)   // (sharded: shard_scale_kernel added it to scale_acc before the reduction, b_p is a partial sum here)
        for (int j = lane; j < w.n; j += 32) sc += w.x[j] * (lambda * w.x[j] + w.bp[j]);
    sc = warp_sum(sc);
    if (lane == 0) {
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
            if (s->n_trace < VILBA_MAX_TRACE) {
                IterRec& r = s->trace[s->n_trace++];
                r.stage = s->stage, r.iteration = s->iter, r.trials = s->qmax, r.result = res;
                r.n_active = s->n_active, r.accepted = s->accepted;
                r.chi0 = s->ini_chi, r.chi1 = s->current_chi, r.lambda = s->lambda, r.lambda_first = s->lambda_first;
            }
            s->iter += 1;
            // optimize(): for (i < iterations && !terminate() && ok)   (sparse_optimizer.cpp:376)
            s->phase = (res != 0 || s->iter >= s->max_iters || s->stop) ? PH_DONE : PH_LINEARIZE;
        }
    }
}

template <bool APPLY>
__global__ void __launch_bounds__(kPointThreads, 2) update_eval_kernel(const DevWindow* __restrict__ wp) {
    const DevWindow w = wp[blockIdx.y];  // one window per grid row
    if (APPLY && w.lm->phase != PH_TRIAL) return;
    extern __shared__ double smem[];
    const KfSmem ks = kf_smem_carve(smem, w.K);
    const int cur = w.lm->cur;
    const double lambda = w.lm->lambda;
    kf_stage<APPLY>(w, ks, cur);
    // pose increments [dP, dPhi] of every key-frame (zeros for fixed ones) next to the states: the landmark
    // back-substitution reads them per edge
    double* xs6 = reinterpret_cast<double*>(reinterpret_cast<char*>(smem) + ((kf_smem_bytes(w.K) + 15) / 16) * 16);
    if (APPLY) {
        for (int i = threadIdx.x; i < 6 * w.K; i += blockDim.x) {
            const int k = i / 6, c = i - 6 * k;
            const int blk = w.kf_block[k];
            xs6[i] = blk >= 0 ? w.x[15 * (size_t)blk + (c < 3 ? c : c + 3)] : 0.0;
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int gl = lane & 7, grp = lane >> 3;
    const int warps_per_cta = blockDim.x >> 5;
    const int gwarp = blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * warps_per_cta;
    const double* pts_in = (cur ? w.pts[1] : w.pts[0]);
    double* pts_out = (cur ? w.pts[0] : w.pts[1]);
    const int total = w.P + (w.shard_owner ? w.NI : 0);  // the IMU edges are evaluated by one rank of a sharded window
    double chi = 0.0, scale = 0.0;

    for (int base = gwarp * 4; base < total; base += nwarps * 4) {
        const int item = base + grp;
        const bool is_pt = item < w.P;
        const bool is_imu = !is_pt && item < total;
        int e0i = 0, e1i = 0;
        V3 Pw = v3(0, 0, 0);
        if (is_pt) {
            e0i = w.pt_obs_begin[item], e1i = w.pt_obs_begin[item + 1];
            Pw = ld3(pts_in + 3 * (size_t)item);
        }
        if (APPLY) {
            // x_l = D^-1 (b_l - sum_i W_i^T x_p(i)),  D = H_ll + lambda I   (block_solver.hpp:459-481)
            V3 acc = v3(0, 0, 0);
            for (int e = e0i + gl; e < e1i; e += 8) {
                const int4 r = w.obs[e];
                const int kf = r.w & OBS_KF_MASK;
                if (!(r.w & OBS_CULLED) && ks.blk[kf] >= 0) {
                    const double* Wp = w.W + 18 * (size_t)e;
                    const double* xs = xs6 + 6 * kf;
#pragma unroll
                    for (int rr = 0; rr < 6; ++rr) {
                        acc.x += Wp[3 * rr + 0] * xs[rr];
                        acc.y += Wp[3 * rr + 1] * xs[rr];
                        acc.z += Wp[3 * rr + 2] * xs[rr];
                    }
                }
            }
            acc.x = group8_sum(acc.x), acc.y = group8_sum(acc.y), acc.z = group8_sum(acc.z);
            if (is_pt) {
                const double* H = w.Hll + 6 * (size_t)item;
                const V3 b = ld3(w.bl + 3 * (size_t)item);
                bool ok;
                const S3 Dinv = s3_inverse(S3{H[0] + lambda, H[1], H[2], H[3] + lambda, H[4], H[5] + lambda}, ok);
                const V3 xl = s3_mul(Dinv, b - acc);
                Pw = Pw + xl;  // VertexSBAPointXYZ::oplusImpl (types_sba.h:52-56)
                if (gl == 0) {
                    st3(pts_out + 3 * (size_t)item, Pw);
                    scale += xl.x * (lambda * xl.x + b.x) + xl.y * (lambda * xl.y + b.y) + xl.z * (lambda * xl.z + b.z);
                }
            }
        }
        for (int e = e0i + gl; e < e1i; e += 8) {
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
        if (is_imu && gl == 0) {
            // one IMU edge pair per group; its first lane evaluates both residuals
            const int e = item - w.P;
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
            const double c2 = rg.x * (wg * rg.x) + rg.y * (wg * rg.y) + rg.z * (wg * rg.z) + ra.x * (wa * ra.x) +
                              ra.y * (wa * ra.y) + ra.z * (wa * ra.z);
            huber(c2, w.huber_bias, rho0, rho1);
            chi += rho0;
        }
    }
    // CTA reduction, one partial per CTA; the last CTA to arrive adds the partials in a fixed order, so that chi2 and the
    // gain scale -- and with them every LM decision -- are bit-reproducible from run to run (no floating-point atomics)
    __shared__ double red[2][kPointThreads / 32];
    __shared__ bool is_last;
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
        w.chi_partial[2 * blockIdx.x] = c;
        w.chi_partial[2 * blockIdx.x + 1] = s;
        __threadfence();
        is_last = atomicAdd(w.chi_counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last && threadIdx.x < 32) {
        __threadfence();
        double c = 0.0, s = 0.0;
        for (int i = lane; i < (int)gridDim.x; i += 32) {
            c += __ldcg(w.chi_partial + 2 * i);
            s += __ldcg(w.chi_partial + 2 * i + 1);
        }
        c = warp_sum(c);
        s = warp_sum(s);
        if (lane == 0) {
            w.lm->chi_acc = c;
            if (APPLY) w.lm->scale_acc = s;
            *w.chi_counter = 0u;
        }
        if (APPLY && !w.sharded) {  // every other CTA of this launch is past its work: the phase may change now
            __syncwarp();
            lm_decide_warp(w, w.lm, lane);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// imu_prepare: information of EdgeNavStatePVR = inverse of cov_P_V_Phi (Optimizer.cpp:2510).
// Gauss-Jordan with partial pivoting, one warp per edge (lane = column of the augmented matrix).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) imu_prepare_kernel(const DevWindow* __restrict__ wp) {
    const DevWindow w = wp[blockIdx.y];  // one window per grid row
    __shared__ double A[9][18];
    const int lane = threadIdx.x;
    for (int e = blockIdx.x; e < w.NI; e += gridDim.x) {
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
    __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// device-side LM controller (single CTA each)
// ------------------------------------------------------------------------------------------------
// after the stage's initial computeActiveErrors: currentChi, iteration counter, phase
__global__ void lm_stage_begin_kernel(const DevWindow* __restrict__ wp, int stage, int max_iters) {
    const DevWindow w = wp[blockIdx.y];  // one window per grid row
    if (threadIdx.x == 0) {
        LmState* s = w.lm;
        s->current_chi = s->chi_acc;  // currentChi = activeRobustChi2() at the stage's first iteration
        s->chi_acc = 0.0;
        s->scale_acc = 0.0;
        s->maxdiag_bits = 0ull;
        s->stage = stage;
        s->iter = 0;
        s->max_iters = max_iters;
        s->n_active = w.E - s->n_culled + (w.shard_owner ? 2 * w.NI : 0);  // sums to the window's count over the ranks
        // for (i < iterations && !terminate() && ok)   (sparse_optimizer.cpp:376)
        s->phase = (max_iters > 0 && !s->stop) ? PH_LINEARIZE : PH_DONE;
    }
}

__global__ void __launch_bounds__(256) lm_iter_begin_kernel(const DevWindow* __restrict__ wp) {
    const DevWindow w = wp[blockIdx.y];  // one window per grid row
    LmState* s = w.lm;
    if (s->phase != PH_LINEARIZE || !w.sharded) return;  // (not sharded: done at the tail of assemble_hpp)
    lm_iter_begin_cta(w, s);
}

__global__ void __launch_bounds__(32) lm_decide_kernel(const DevWindow* __restrict__ wp) {
    const DevWindow w = wp[blockIdx.y];  // one window per grid row
    LmState* s = w.lm;
    if (s->phase != PH_TRIAL || !w.sharded) return;  // (not sharded: decided at the tail of update_eval)
    lm_decide_warp(w, s, threadIdx.x);
}

// ------------------------------------------------------------------------------------------------
// point-sharded window: the two small pieces that replace the allreduce of H_pp | b_p
// ------------------------------------------------------------------------------------------------
// after assemble_hpp: diag(H_pp) of this rank's partial sum and its max |diag H_ll| into the buffer RED_DIAG sums
__global__ void __launch_bounds__(256) shard_diag_kernel(const DevWindow* __restrict__ wp) {
    const DevWindow* w = wp + blockIdx.y;
    if (!w->sharded) return;
    const bool lin = w->lm->phase == PH_LINEARIZE;  // other phases: zeros (the in-place sum behind this kernel runs in every slot)
    for (int d = threadIdx.x; d < w->n + w->shard_world; d += blockDim.x)
        w->diag_red[d] = !lin ? 0.0 : d < w->n ? w->Hpp[(size_t)d * w->n + d]
                                  : (d - w->n == w->shard_rank ? __longlong_as_double((long long)w->lm->maxdiag_bits) : 0.0);
}
// after update_eval: the pose part of computeScale from this rank's partial b_p (lambda x^2 once per window)
__global__ void __launch_bounds__(32) shard_scale_kernel(const DevWindow* __restrict__ wp) {
    const DevWindow* w = wp + blockIdx.y;
    if (w->lm->phase != PH_TRIAL || !w->sharded) return;
    const double lambda = w->shard_owner ? w->lm->lambda : 0.0;
    double sc = 0.0;
    for (int j = threadIdx.x; j < w->n; j += 32) sc += w->x[j] * (lambda * w->x[j] + w->bp[j]);
    sc = warp_sum(sc);
    if (threadIdx.x == 0) w->lm->scale_acc += sc;
}

// ------------------------------------------------------------------------------------------------
// cull (after stage 1) and final outlier flags: chi2 from the cached errors, depth from the estimates
// ------------------------------------------------------------------------------------------------
template <bool CULL>
__global__ void __launch_bounds__(kPointThreads) flags_kernel(const DevWindow* __restrict__ wp, uint8_t* outlier) {
    const DevWindow w = wp[blockIdx.y];  // one window per grid row
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
        const V3 Pw = ld3((cur ? w.pts[1] : w.pts[0]) + 3 * (size_t)p);
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
                w.outlier[e] = bad ? 1 : 0;
            }
        }
    }
    if (CULL) {
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0 && cnt) atomicAdd(&w.lm->n_culled, cnt);
    }
}

// ------------------------------------------------------------------------------------------------
// reset (restart every window from its uploaded initial state) and export (pack the results of every
// window into one contiguous region => one D2H copy for the whole batch)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) reset_kernel(const DevWindow* __restrict__ wp) {
    const DevWindow w = wp[blockIdx.y];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (int i = tid; i < 22 * w.K; i += nt) w.kf_state[0][i] = w.kf_state[1][i] = w.kf_state0[i];
    for (int i = tid; i < 3 * w.P; i += nt) w.pts[0][i] = w.pts[1][i] = w.pts0[i];
    // the edge records {u, v, inv sigma^2 as f32 bits, key-frame index | flags} from the three uploaded arrays
    const int2* uv = reinterpret_cast<const int2*>(w.obs0);
    const int* is2 = reinterpret_cast<const int*>(w.obs0 + 8 * (size_t)w.E);
    const int* kf = reinterpret_cast<const int*>(w.obs0 + 12 * (size_t)w.E);
    for (int i = tid; i < w.E; i += nt) {
        const int2 p = uv[i];
        w.obs[i] = make_int4(p.x, p.y, is2[i], kf[i] | w.obs_flags0);
        w.obs_chi2[i] = 0.0;
    }
    int* lm = reinterpret_cast<int*>(w.lm);
    for (int i = tid; i < (int)(sizeof(LmState) / sizeof(int)); i += nt) lm[i] = 0;
    if (tid == 0) w.chi_counter[0] = w.chi_counter[1] = 0u;  // arrival counters of update_eval / assemble_hpp
}

__global__ void __launch_bounds__(256) export_kernel(const DevWindow* __restrict__ wp) {
    const DevWindow w = wp[blockIdx.y];
    const int cur = w.lm->cur;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (int i = tid; i < 22 * w.K; i += nt) w.out_kf_state[i] = (cur ? w.kf_state[1] : w.kf_state[0])[i];
    for (int i = tid; i < 3 * w.P; i += nt) w.out_pts[i] = (cur ? w.pts[1] : w.pts[0])[i];
    for (int i = tid; i < w.E; i += nt) {
        w.out_chi2[i] = w.obs_chi2[i];
        w.out_outlier[i] = w.outlier[i];
    }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
size_t point_smem_bytes(int K) { return kf_smem_bytes(K) + 16 + sizeof(double) * 6 * (size_t)K; }

cudaError_t launch_imu_prepare(cudaStream_t s, const DevWindow* wp, const LaunchDims& d) {
    imu_prepare_kernel<<<dim3(d.imu_grid, d.n_windows), 32, 0, s>>>(wp);
    return cudaGetLastError();
}
cudaError_t launch_eval_initial(cudaStream_t s, const DevWindow* wp, const LaunchDims& d) {
    update_eval_kernel<false><<<dim3(d.point_grid, d.n_windows), kPointThreads, d.smem_point, s>>>(wp);
    return cudaGetLastError();
}
cudaError_t launch_update_eval_apply(cudaStream_t s, const DevWindow* wp, const LaunchDims& d) {
    update_eval_kernel<true><<<dim3(d.point_grid, d.n_windows), kPointThreads, d.smem_point, s>>>(wp);
    return cudaGetLastError();
}
cudaError_t launch_stage_begin(cudaStream_t s, const DevWindow* wp, const LaunchDims& d, int stage, int max_iters) {
    lm_stage_begin_kernel<<<dim3(1, d.n_windows), 32, 0, s>>>(wp, stage, max_iters);
    return cudaGetLastError();
}
cudaError_t launch_lm_iter_begin(cudaStream_t s, const DevWindow* wp, const LaunchDims& d) {
    lm_iter_begin_kernel<<<dim3(1, d.n_windows), 256, 0, s>>>(wp);
    return cudaGetLastError();
}
cudaError_t launch_shard_diag(cudaStream_t s, const DevWindow* wp, const LaunchDims& d) {
    shard_diag_kernel<<<dim3(1, d.n_windows), 256, 0, s>>>(wp);
    return cudaGetLastError();
}
cudaError_t launch_shard_scale(cudaStream_t s, const DevWindow* wp, const LaunchDims& d) {
    shard_scale_kernel<<<dim3(1, d.n_windows), 32, 0, s>>>(wp);
    return cudaGetLastError();
}
cudaError_t launch_lm_decide(cudaStream_t s, const DevWindow* wp, const LaunchDims& d) {
    lm_decide_kernel<<<dim3(1, d.n_windows), 32, 0, s>>>(wp);
    return cudaGetLastError();
}
cudaError_t launch_cull(cudaStream_t s, const DevWindow* wp, const LaunchDims& d) {
    flags_kernel<true><<<dim3(d.point_grid, d.n_windows), kPointThreads, d.smem_point, s>>>(wp, nullptr);
    return cudaGetLastError();
}
cudaError_t launch_final_flags(cudaStream_t s, const DevWindow* wp, const LaunchDims& d) {
    flags_kernel<false><<<dim3(d.point_grid, d.n_windows), kPointThreads, d.smem_point, s>>>(wp, nullptr);
    return cudaGetLastError();
}
cudaError_t launch_reset(cudaStream_t s, const DevWindow* wp, const LaunchDims& d) {
    reset_kernel<<<dim3(d.point_grid, d.n_windows), 256, 0, s>>>(wp);
    return cudaGetLastError();
}
cudaError_t launch_export(cudaStream_t s, const DevWindow* wp, const LaunchDims& d) {
    export_kernel<<<dim3(d.point_grid, d.n_windows), 256, 0, s>>>(wp);
    return cudaGetLastError();
}
cudaError_t configure_point_kernels(const LaunchDims& d) {
    cudaError_t e = opt_in_max_smem(update_eval_kernel<true>);
    if (e != cudaSuccess) return e;
    e = opt_in_max_smem(update_eval_kernel<false>);
    if (e != cudaSuccess) return e;
    e = opt_in_max_smem(flags_kernel<true>);
    if (e != cudaSuccess) return e;
    return opt_in_max_smem(flags_kernel<false>);
}

}  // namespace vilba
