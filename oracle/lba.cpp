// oracle/lba.cpp -- TEST INFRASTRUCTURE (CPU oracle), not part of the product path.
//
// CPU restatement of the solve phases of Optimizer::LocalBundleAdjustmentNavState
// (src/Optimizer.cpp:2643-2701) on top of a restatement of the g2o machinery it drives:
//   SparseOptimizer::{initializeOptimization,optimize,computeActiveErrors,activeRobustChi2,update,
//                     push,pop,discardTop}      Thirdparty/g2o/g2o/core/sparse_optimizer.cpp:61-114,199-267,354-435
//   BlockSolver::{buildSystem,setLambda,solve,restoreDiagonal}   g2o/core/block_solver.hpp:354-604
//   BaseBinaryEdge/BaseMultiEdge::constructQuadraticForm         g2o/core/base_binary_edge.hpp:55-120,
//                                                                g2o/core/base_multi_edge.hpp:36-48,171-222
//   OptimizationAlgorithmLevenberg::solve                        g2o/core/optimization_algorithm_levenberg.cpp:61-189
//   LinearSolverEigen::solve (SimplicialLDLT)                    g2o/solvers/linear_solver_eigen.h:94-124
// Block-sparse containers are replaced by dense storage; edges are visited in g2o's deterministic
// order (IMU pairs in key-frame order, then mono edges point-major), so the FP64 sums are a fixed
// reference.  SimplicialLDLT (Eigen, un-vendored) is replaced by a dense LDL^T that, like it, fails
// only on an exactly-zero pivot.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <limits>
#include <vector>

#include "edges.h"
#include "vilba_oracle.h"

namespace oracle {

typedef Mat<6, 6> Mat6;

struct Problem {
    const vilba_window& w;
    vilba_params prm;
    int K, NI, P, E;
    int n_free, n;  // n = 15 * n_free (reduced camera system dimension)
    std::vector<NavState> ns, ns_backup;
    std::vector<Vec3> pts, pts_backup;
    std::vector<int> kf_block;  // block index among free key-frames, -1 if fixed
    Calib calib;
    Vec3 g;
    std::vector<Preintegrator> M;
    std::vector<Mat9> info_pvr;
    std::vector<Mat6> info_bias;
    // cached _error of every edge (only refreshed for ACTIVE edges, like g2o)
    std::vector<double> mono_err, pvr_err, bias_err;
    std::vector<uint8_t> mono_level, mono_robust, pt_active;
    // normal equations
    std::vector<double> Hpp, Hll, Hpl, b, x;
    std::vector<double> S, bs, Dinv, coeff, ldl;
    std::vector<double> diag_backup_p, diag_backup_l;
    volatile const uint8_t* stop;
    // LM state (OptimizationAlgorithmLevenberg members)
    double currentLambda, ni;
    int nBad, levenbergIterations;
    long edges_linearized;

    Problem(const vilba_window& win, const vilba_params& p, const volatile uint8_t* stop_flag)
        : w(win), prm(p), stop(stop_flag) {
        K = w.n_kf;
        NI = w.n_imu;
        P = w.n_pts;
        E = w.n_obs;
        ns.resize(K);
        kf_block.assign(K, -1);
        n_free = 0;
        for (int k = 0; k < K; ++k) {
            ns[k].load(w.kf_state + VILBA_NS_DOUBLES * k);
            if (!(w.kf_flags[k] & VILBA_KF_FIXED)) kf_block[k] = n_free++;
        }
        n = 15 * n_free;
        pts.resize(P);
        for (int p_ = 0; p_ < P; ++p_) pts[p_] = Vec3::from(w.pt_xyz + 3 * p_);
        calib.fx = w.fx;
        calib.fy = w.fy;
        calib.cx = w.cx;
        calib.cy = w.cy;
        calib.Rbc = Mat3::from(w.Rbc);
        calib.Pbc = Vec3::from(w.Pbc);
        g = Vec3::from(w.gravity);
        M.resize(NI);
        info_pvr.resize(NI);
        info_bias.resize(NI);
        for (int e = 0; e < NI; ++e) {
            M[e].load(w.imu_preint + VILBA_PREINT_DOUBLES * e);
            inverse_lu<9>(M[e].cov, info_pvr[e]);  // Optimizer.cpp:2510 getCovPVPhi().inverse()
            Mat6 I = Mat6::identity();             // Optimizer.cpp:2490-2492,2533
            for (int d = 0; d < 3; ++d) {
                I(d, d) = 1.0 / prm.gyr_bias_rw2;
                I(3 + d, 3 + d) = 1.0 / prm.acc_bias_rw2;
            }
            for (int i = 0; i < 36; ++i) info_bias[e].a[i] = I.a[i] / M[e].dt;
        }
        mono_err.assign(2 * (size_t)E, 0.0);
        pvr_err.assign(9 * (size_t)NI, 0.0);
        bias_err.assign(6 * (size_t)NI, 0.0);
        mono_level.assign(E, 0);
        mono_robust.assign(E, 1);
        pt_active.assign(P, 1);
        Hpp.assign((size_t)n * n, 0.0);
        Hll.assign(9 * (size_t)P, 0.0);
        Hpl.assign(18 * (size_t)E, 0.0);
        b.assign(n + 3 * (size_t)P, 0.0);
        x.assign(n + 3 * (size_t)P, 0.0);
        S.assign((size_t)n * n, 0.0);
        ldl.assign((size_t)n * n, 0.0);
        bs.assign(n, 0.0);
        coeff.assign(n, 0.0);
        Dinv.assign(9 * (size_t)P, 0.0);
        currentLambda = -1.;
        ni = 2.;
        nBad = 0;
        levenbergIterations = 0;
        edges_linearized = 0;
    }

    bool terminate() const { return stop && *stop; }  // sparse_optimizer.h:188

    int off_pvr(int kf) const { return 15 * kf_block[kf]; }
    int off_bias(int kf) const { return 15 * kf_block[kf] + 9; }
    bool kf_free(int kf) const { return kf_block[kf] >= 0; }

    // ------------------------------------------------------------------------------------------
    // SparseOptimizer::initializeOptimization(level=0): active edges / vertices
    // (sparse_optimizer.cpp:199-267).  IMU edges are always at level 0.
    // ------------------------------------------------------------------------------------------
    int initialize_optimization() {
        int n_active = 2 * NI;
        for (int p_ = 0; p_ < P; ++p_) {
            int cnt = 0;
            for (int e = w.pt_obs_begin[p_]; e < w.pt_obs_begin[p_ + 1]; ++e)
                if (mono_level[e] == 0) ++cnt;
            pt_active[p_] = cnt > 0;  // vertices without an active edge drop out (:241-243)
            n_active += cnt;
        }
        return n_active;
    }

    // SparseOptimizer::computeActiveErrors (sparse_optimizer.cpp:61-76)
    void compute_active_errors() {
        for (int e = 0; e < NI; ++e) {
            int i = w.imu_kf_i[e], j = w.imu_kf_j[e];
            Mat<9, 1> r = pvr_error(ns[i], ns[j], ns[i], M[e], g);
            r.store(&pvr_err[9 * (size_t)e]);
            Mat<6, 1> rb = bias_error(ns[i], ns[j]);
            rb.store(&bias_err[6 * (size_t)e]);
        }
        for (int p_ = 0; p_ < P; ++p_)
            for (int e = w.pt_obs_begin[p_]; e < w.pt_obs_begin[p_ + 1]; ++e) {
                if (mono_level[e] != 0) continue;
                mono_error(ns[w.obs_kf[e]], pts[p_], calib, (double)w.obs_uv[2 * e], (double)w.obs_uv[2 * e + 1],
                           &mono_err[2 * (size_t)e]);
            }
    }

    // BaseEdge::chi2 (base_edge.h:58-61)
    double chi2_mono(int e) const {
        double is2 = (double)w.obs_inv_sigma2[e];
        const double* r = &mono_err[2 * (size_t)e];
        // information()*_error with information = Identity*invSigma2 (dense 2x2 product incl. zeros)
        double o0 = is2 * r[0] + 0.0 * r[1];
        double o1 = 0.0 * r[0] + is2 * r[1];
        return r[0] * o0 + r[1] * o1;
    }
    double chi2_pvr(int e) const {
        Mat<9, 1> r = Mat<9, 1>::from(&pvr_err[9 * (size_t)e]);
        return dot(r, info_pvr[e] * r);
    }
    double chi2_bias(int e) const {
        Mat<6, 1> r = Mat<6, 1>::from(&bias_err[6 * (size_t)e]);
        return dot(r, info_bias[e] * r);
    }

    // SparseOptimizer::activeRobustChi2 (sparse_optimizer.cpp:100-114)
    double active_robust_chi2() const {
        double rho[3];
        double chi = 0.0;
        for (int e = 0; e < NI; ++e) {
            huber(chi2_pvr(e), prm.huber_pvr, rho);
            chi += rho[0];
            huber(chi2_bias(e), prm.huber_bias, rho);
            chi += rho[0];
        }
        for (int e = 0; e < E; ++e) {
            if (mono_level[e] != 0) continue;
            if (mono_robust[e]) {
                huber(chi2_mono(e), prm.huber_mono, rho);
                chi += rho[0];
            } else
                chi += chi2_mono(e);
        }
        return chi;
    }

    // add a dense block into the upper triangle of Hpp, transposing if it lies below the diagonal
    // (BlockSolver::buildStructure maps lower blocks onto their transposed upper twin, :217-229)
    template <int R, int C>
    void add_hpp(int r0, int c0, const Mat<R, C>& Bk) {
        if (r0 <= c0) {
            for (int r = 0; r < R; ++r)
                for (int c = 0; c < C; ++c) Hpp[(size_t)(r0 + r) * n + c0 + c] += Bk(r, c);
        } else {
            for (int r = 0; r < R; ++r)
                for (int c = 0; c < C; ++c) Hpp[(size_t)(c0 + c) * n + r0 + r] += Bk(r, c);
        }
    }

    // ------------------------------------------------------------------------------------------
    // BlockSolver::buildSystem (block_solver.hpp:502-560): linearizeOplus + constructQuadraticForm
    // ------------------------------------------------------------------------------------------
    void build_system() {
        std::fill(Hpp.begin(), Hpp.end(), 0.0);
        std::fill(Hll.begin(), Hll.end(), 0.0);
        std::fill(Hpl.begin(), Hpl.end(), 0.0);
        std::fill(b.begin(), b.end(), 0.0);
        double rho[3];
        for (int e = 0; e < NI; ++e) {
            int i = w.imu_kf_i[e], j = w.imu_kf_j[e];
            ++edges_linearized;
            // ---- EdgeNavStatePVR: BaseMultiEdge<9> over (PVR_i, PVR_j, Bias_i) ----
            {
                Mat<9, 1> err = Mat<9, 1>::from(&pvr_err[9 * (size_t)e]);
                Mat9 Ji, Jj;
                Mat<9, 6> Jb;
                pvr_linearize(ns[i], ns[j], ns[i], M[e], g, err, Ji, Jj, Jb);
                huber(chi2_pvr(e), prm.huber_pvr, rho);  // base_multi_edge.hpp:36-48 (robust branch)
                Mat<9, 1> omega_r = -(info_pvr[e] * err);
                omega_r = omega_r * rho[1];
                Mat9 omega = rho[1] * info_pvr[e];  // base_edge.h:96-102
                // computeQuadraticForm (base_multi_edge.hpp:171-222)
                Mat9 AtOi = transpose(Ji) * omega;
                Mat9 AtOj = transpose(Jj) * omega;
                Mat<6, 9> AtOb = transpose(Jb) * omega;
                bool fi = kf_free(i), fj = kf_free(j);
                if (fi) {
                    add_hpp(off_pvr(i), off_pvr(i), AtOi * Ji);
                    Mat<9, 1> bi = transpose(Ji) * omega_r;
                    for (int d = 0; d < 9; ++d) b[off_pvr(i) + d] += bi[d];
                    if (fj) add_hpp(off_pvr(i), off_pvr(j), AtOi * Jj);
                    add_hpp(off_pvr(i), off_bias(i), AtOi * Jb);
                }
                if (fj) {
                    add_hpp(off_pvr(j), off_pvr(j), AtOj * Jj);
                    Mat<9, 1> bj = transpose(Jj) * omega_r;
                    for (int d = 0; d < 9; ++d) b[off_pvr(j) + d] += bj[d];
                    if (fi) add_hpp(off_pvr(j), off_bias(i), AtOj * Jb);  // lands transposed (Bias_i < PVR_j)
                }
                if (fi) {
                    add_hpp(off_bias(i), off_bias(i), AtOb * Jb);
                    Mat<6, 1> bb = transpose(Jb) * omega_r;
                    for (int d = 0; d < 6; ++d) b[off_bias(i) + d] += bb[d];
                }
            }
            ++edges_linearized;
            // ---- EdgeNavStateBias: BaseBinaryEdge<6> (Bias_i, Bias_j), A=-I, B=+I, robust ----
            {
                Mat<6, 1> err = Mat<6, 1>::from(&bias_err[6 * (size_t)e]);
                Mat6 A = -Mat6::identity(), B = Mat6::identity();
                huber(chi2_bias(e), prm.huber_bias, rho);
                Mat<6, 1> omega_r = -(info_bias[e] * err);
                Mat6 wOmega = rho[1] * info_bias[e];
                omega_r = omega_r * rho[1];
                bool fi = kf_free(i), fj = kf_free(j);
                if (fi) {
                    Mat<6, 1> bi = transpose(A) * omega_r;
                    for (int d = 0; d < 6; ++d) b[off_bias(i) + d] += bi[d];
                    add_hpp(off_bias(i), off_bias(i), (transpose(A) * wOmega) * A);
                    if (fj) add_hpp(off_bias(i), off_bias(j), (transpose(A) * wOmega) * B);
                }
                if (fj) {
                    Mat<6, 1> bj = transpose(B) * omega_r;
                    for (int d = 0; d < 6; ++d) b[off_bias(j) + d] += bj[d];
                    add_hpp(off_bias(j), off_bias(j), (transpose(B) * wOmega) * B);
                }
            }
        }
        // ---- EdgeNavStatePVRPointXYZ: BaseBinaryEdge<2>(point, PVR) ----
        for (int p_ = 0; p_ < P; ++p_) {
            double* Hl = &Hll[9 * (size_t)p_];
            double* bl = &b[n + 3 * (size_t)p_];
            for (int e = w.pt_obs_begin[p_]; e < w.pt_obs_begin[p_ + 1]; ++e) {
                if (mono_level[e] != 0) continue;
                ++edges_linearized;
                int kf = w.obs_kf[e];
                Mat<2, 3> A;
                Mat<2, 9> B;
                mono_linearize(ns[kf], pts[p_], calib, A, B);
                Mat<2, 2> omega = Mat<2, 2>::identity() * (double)w.obs_inv_sigma2[e];
                Mat<2, 1> err = Mat<2, 1>::from(&mono_err[2 * (size_t)e]);
                Mat<2, 1> omega_r = -(omega * err);
                bool to_free = kf_free(kf);
                Mat3 HllAdd;
                Mat<3, 1> blAdd;
                Mat<9, 3> HplT;  // _hessianTransposed: B^T * wOmega * A  (9x3)
                Mat9 HppAdd;
                Mat<9, 1> bpAdd;
                if (!mono_robust[e]) {  // base_binary_edge.hpp:77-93
                    Mat<3, 2> AtO = transpose(A) * omega;
                    blAdd = transpose(A) * omega_r;
                    HllAdd = AtO * A;
                    if (to_free) {
                        HplT = transpose(B) * transpose(AtO);
                        bpAdd = transpose(B) * omega_r;
                        HppAdd = (transpose(B) * omega) * B;
                    }
                } else {  // :94-116
                    huber(chi2_mono(e), prm.huber_mono, rho);
                    Mat<2, 2> wOmega = rho[1] * omega;
                    omega_r = omega_r * rho[1];
                    blAdd = transpose(A) * omega_r;
                    HllAdd = (transpose(A) * wOmega) * A;
                    if (to_free) {
                        HplT = (transpose(B) * wOmega) * A;
                        bpAdd = transpose(B) * omega_r;
                        HppAdd = (transpose(B) * wOmega) * B;
                    }
                }
                for (int d = 0; d < 9; ++d) Hl[d] += HllAdd.a[d];
                for (int d = 0; d < 3; ++d) bl[d] += blAdd[d];
                if (to_free) {
                    // rows of V (3..5) are exact zeros: keep the 6 nonzero rows [P, Phi]
                    double* Hp = &Hpl[18 * (size_t)e];
                    for (int r = 0; r < 3; ++r)
                        for (int c = 0; c < 3; ++c) {
                            Hp[r * 3 + c] += HplT(r, c);
                            Hp[(3 + r) * 3 + c] += HplT(6 + r, c);
                        }
                    add_hpp(off_pvr(kf), off_pvr(kf), HppAdd);
                    for (int d = 0; d < 9; ++d) b[off_pvr(kf) + d] += bpAdd[d];
                }
            }
        }
        // the diagonal blocks were accumulated as full blocks: nothing to mirror.
    }

    // OptimizationAlgorithmLevenberg::computeLambdaInit (optimization_algorithm_levenberg.cpp:166-180)
    double compute_lambda_init() const {
        double maxDiagonal = 0.;
        for (int d = 0; d < n; ++d) maxDiagonal = std::max(std::fabs(Hpp[(size_t)d * n + d]), maxDiagonal);
        for (int p_ = 0; p_ < P; ++p_) {
            if (!pt_active[p_]) continue;
            for (int d = 0; d < 3; ++d) maxDiagonal = std::max(std::fabs(Hll[9 * (size_t)p_ + 4 * d]), maxDiagonal);
        }
        return prm.lm_tau * maxDiagonal;
    }

    // BlockSolver::setLambda / restoreDiagonal (block_solver.hpp:564-604)
    void set_lambda(double lambda) {
        diag_backup_p.resize(n);
        diag_backup_l.resize(3 * (size_t)P);
        for (int d = 0; d < n; ++d) {
            diag_backup_p[d] = Hpp[(size_t)d * n + d];
            Hpp[(size_t)d * n + d] += lambda;
        }
        for (int p_ = 0; p_ < P; ++p_)
            for (int d = 0; d < 3; ++d) {
                diag_backup_l[3 * (size_t)p_ + d] = Hll[9 * (size_t)p_ + 4 * d];
                Hll[9 * (size_t)p_ + 4 * d] += lambda;
            }
    }
    void restore_diagonal() {
        for (int d = 0; d < n; ++d) Hpp[(size_t)d * n + d] = diag_backup_p[d];
        for (int p_ = 0; p_ < P; ++p_)
            for (int d = 0; d < 3; ++d) Hll[9 * (size_t)p_ + 4 * d] = diag_backup_l[3 * (size_t)p_ + d];
    }

    // Dense stand-in for LinearSolverEigen::solve (linear_solver_eigen.h:94-124): LDL^T of the upper
    // triangle; like Eigen's SimplicialLDLT it reports failure only for an exactly-zero pivot.
    bool ldlt_solve(const std::vector<double>& A, const std::vector<double>& rhs, double* sol) {
        // factor A = L D L^T (row-oriented up-looking form), L unit lower in `ldl`, D in `d`
        std::vector<double> d(n), v(n), y(n);
        for (int j = 0; j < n; ++j) {
            const double* Lj = &ldl[(size_t)j * n];
            double dj = A[(size_t)j * n + j];
            for (int k = 0; k < j; ++k) {
                v[k] = Lj[k] * d[k];
                dj -= Lj[k] * v[k];
            }
            d[j] = dj;
            if (dj == 0.0) return false;
            for (int i = j + 1; i < n; ++i) {
                const double* Li = &ldl[(size_t)i * n];
                double s = A[(size_t)j * n + i];  // upper triangle: A(j,i) == A(i,j)
                for (int k = 0; k < j; ++k) s -= Li[k] * v[k];
                ldl[(size_t)i * n + j] = s / dj;
            }
        }
        for (int i = 0; i < n; ++i) {
            double s = rhs[i];
            const double* Li = &ldl[(size_t)i * n];
            for (int k = 0; k < i; ++k) s -= Li[k] * y[k];
            y[i] = s;
        }
        for (int i = 0; i < n; ++i) y[i] /= d[i];
        for (int i = n - 1; i >= 0; --i) {
            double s = y[i];
            for (int k = i + 1; k < n; ++k) s -= ldl[(size_t)k * n + i] * sol[k];
            sol[i] = s;
        }
        return true;
    }

    // ------------------------------------------------------------------------------------------
    // BlockSolver::solve, Schur branch (block_solver.hpp:369-486)
    // ------------------------------------------------------------------------------------------
    bool solve_schur() {
        S = Hpp;  // _Hschur = _Hpp (upper blocks)
        std::fill(coeff.begin(), coeff.end(), 0.0);
        for (int p_ = 0; p_ < P; ++p_) {
            if (!pt_active[p_]) continue;
            Mat3 D = Mat3::from(&Hll[9 * (size_t)p_]);
            Mat3 Di;
            inverse_lu<3>(D, Di);  // D->inverse() on a dynamic-size 3x3 (:389)
            Di.store(&Dinv[9 * (size_t)p_]);
            Vec3 db = Di * Vec3::from(&b[n + 3 * (size_t)p_]);
            const int e0 = w.pt_obs_begin[p_], e1 = w.pt_obs_begin[p_ + 1];
            for (int ei = e0; ei < e1; ++ei) {
                if (mono_level[ei] != 0 || !kf_free(w.obs_kf[ei])) continue;
                int o1 = off_pvr(w.obs_kf[ei]);
                Mat<6, 3> Bi = Mat<6, 3>::from(&Hpl[18 * (size_t)ei]);
                Mat<6, 3> BDinv = Bi * Di;
                Mat<6, 1> Bb = Bi * db;
                for (int r = 0; r < 3; ++r) {
                    coeff[o1 + r] += Bb[r];
                    coeff[o1 + 6 + r] += Bb[3 + r];
                }
                for (int ej = ei; ej < e1; ++ej) {  // i2 >= i1: observations are ordered by key-frame
                    if (mono_level[ej] != 0 || !kf_free(w.obs_kf[ej])) continue;
                    int o2 = off_pvr(w.obs_kf[ej]);
                    Mat<6, 3> Bj = Mat<6, 3>::from(&Hpl[18 * (size_t)ej]);
                    Mat6 upd = BDinv * transpose(Bj);
                    // scatter the [P,Phi] rows/cols into the 9x9 PVR block (V rows/cols stay untouched)
                    for (int r = 0; r < 6; ++r)
                        for (int c = 0; c < 6; ++c) {
                            int rr = o1 + (r < 3 ? r : r + 3), cc = o2 + (c < 3 ? c : c + 3);
                            if (o1 <= o2)
                                S[(size_t)rr * n + cc] -= upd(r, c);
                            else
                                S[(size_t)cc * n + rr] -= upd(r, c);
                        }
                }
            }
        }
        for (int i = 0; i < n; ++i) bs[i] = b[i] - coeff[i];
        bool ok = ldlt_solve(S, bs, x.data());
        if (!ok) return false;
        // landmarks: xl = Dinv * (bl - Hpl^T xp)   (:459-481)
        for (int p_ = 0; p_ < P; ++p_) {
            double* xl = &x[n + 3 * (size_t)p_];
            if (!pt_active[p_]) {
                xl[0] = xl[1] = xl[2] = 0.0;
                continue;
            }
            Vec3 cl = Vec3::from(&b[n + 3 * (size_t)p_]);
            for (int e = w.pt_obs_begin[p_]; e < w.pt_obs_begin[p_ + 1]; ++e) {
                if (mono_level[e] != 0 || !kf_free(w.obs_kf[e])) continue;
                int o1 = off_pvr(w.obs_kf[e]);
                const double* Bp = &Hpl[18 * (size_t)e];
                for (int c = 0; c < 3; ++c) {
                    double s = 0.0;
                    for (int r = 0; r < 6; ++r) s += Bp[r * 3 + c] * (-x[o1 + (r < 3 ? r : r + 3)]);
                    cl[c] += s;
                }
            }
            Vec3 r = Mat3::from(&Dinv[9 * (size_t)p_]) * cl;
            xl[0] = r[0];
            xl[1] = r[1];
            xl[2] = r[2];
        }
        return true;
    }

    // SparseOptimizer::update (sparse_optimizer.cpp:422-435)
    void apply_update() {
        for (int k = 0; k < K; ++k) {
            if (!kf_free(k)) continue;
            ns[k].IncSmallPVR(&x[off_pvr(k)]);    // VertexNavStatePVR::oplusImpl  (g2otypes.h:505-510)
            ns[k].IncSmallBias(&x[off_bias(k)]);  // VertexNavStateBias::oplusImpl (g2otypes.h:541-546)
        }
        for (int p_ = 0; p_ < P; ++p_) {
            if (!pt_active[p_]) continue;
            for (int d = 0; d < 3; ++d) pts[p_][d] += x[n + 3 * (size_t)p_ + d];  // types_sba.h:52-56
        }
    }

    // OptimizationAlgorithmLevenberg::computeScale (:182-189)
    double compute_scale() const {
        double scale = 0.;
        for (size_t j = 0; j < x.size(); ++j) scale += x[j] * (currentLambda * x[j] + b[j]);
        return scale;
    }

    // OptimizationAlgorithmLevenberg::solve (:61-164).  Returns 0 OK, 1 Terminate.
    int lm_solve(int iteration, vilba_iter_record& rec) {
        compute_active_errors();
        double currentChi = active_robust_chi2();
        double tempChi = currentChi;
        double iniChi = currentChi;
        build_system();
        if (iteration == 0) {
            currentLambda = compute_lambda_init();
            ni = 2;
            nBad = 0;
        }
        rec.chi2_initial = iniChi;
        rec.lambda_first_trial = currentLambda;
        double rho = 0;
        int& qmax = levenbergIterations;
        qmax = 0;
        int accepted = 0;
        do {
            ns_backup = ns;  // _optimizer->push()
            pts_backup = pts;
            set_lambda(currentLambda);
            bool ok2 = solve_schur();
            apply_update();
            restore_diagonal();
            compute_active_errors();
            tempChi = active_robust_chi2();
            if (!ok2) tempChi = std::numeric_limits<double>::max();
            rho = (currentChi - tempChi);
            double scale = compute_scale();
            scale += 1e-3;
            rho /= scale;
            if (rho > 0 && std::isfinite(tempChi)) {
                double alpha = 1. - std::pow((2 * rho - 1), 3);
                alpha = (std::min)(alpha, prm.lm_good_hi);
                double scaleFactor = (std::max)(prm.lm_good_lo, alpha);
                currentLambda *= scaleFactor;
                ni = 2;
                currentChi = tempChi;
                accepted = 1;  // discardTop
            } else {
                currentLambda *= ni;
                ni *= 2;
                ns = ns_backup;  // pop: estimates restored, edge errors stay stale
                pts = pts_backup;
                accepted = 0;
            }
            qmax++;
        } while (rho < 0 && qmax < prm.max_trials && !terminate());
        rec.trials = qmax;
        rec.accepted = accepted;
        rec.chi2_final = currentChi;
        rec.lambda = currentLambda;
        if (qmax == prm.max_trials || rho == 0) return 1;
        if ((iniChi - currentChi) * 1e3 < iniChi)
            nBad++;
        else
            nBad = 0;
        if (nBad >= 3) return 1;
        return 0;
    }

    // SparseOptimizer::optimize (sparse_optimizer.cpp:354-419)
    void optimize(int iterations, int stage, int n_active, vilba_result* out) {
        bool ok = true;
        for (int i = 0; i < iterations && !terminate() && ok; ++i) {
            vilba_iter_record rec;
            std::memset(&rec, 0, sizeof(rec));
            rec.stage = stage;
            rec.iteration = i;
            rec.n_active_edges = n_active;
            int result = lm_solve(i, rec);
            rec.result = result;
            ok = (result == 0);
            if (out->n_trace < VILBA_MAX_TRACE) out->trace[out->n_trace++] = rec;
        }
    }
};

}  // namespace oracle

using namespace oracle;

extern "C" {

void vilba_default_params_oracle(vilba_params* p) {
    std::memset(p, 0, sizeof(*p));
    p->iters_stage1 = 5;
    p->iters_stage2 = 10;
    p->max_trials = 10;
    p->huber_mono = (double)(float)std::sqrt(5.991);            // const float thHuberMono = sqrt(5.991)
    p->huber_pvr = (double)(float)std::sqrt(100 * 21.666);      // const float thHuberNavStatePVR
    p->huber_bias = (double)(float)std::sqrt(100 * 16.812);     // const float thHuberNavStateBias
    p->chi2_gate = 5.991;
    p->lm_tau = 1e-5;
    p->lm_good_lo = 1. / 3.;
    p->lm_good_hi = 2. / 3.;
    p->gyr_bias_rw2 = 2.0e-5 * 2.0e-5;
    p->acc_bias_rw2 = 5.0e-3 * 5.0e-3;
    p->gyr_meas_cov = 1.7e-4 * 1.7e-4 / 0.005;
    p->acc_meas_cov = 2.0e-3 * 2.0e-3 / 0.005 * 100;
}

static int check_window(const vilba_window* w) {
    if (!w || w->n_kf <= 0 || w->n_imu < 0 || w->n_pts < 0 || w->n_obs < 0) return VILBA_ERR_ARG;
    if (!w->kf_state || !w->kf_flags) return VILBA_ERR_ARG;
    if (w->n_imu && (!w->imu_kf_i || !w->imu_kf_j || !w->imu_preint)) return VILBA_ERR_ARG;
    if (w->n_pts && (!w->pt_xyz || !w->pt_obs_begin)) return VILBA_ERR_ARG;
    if (w->n_obs && (!w->obs_kf || !w->obs_uv || !w->obs_inv_sigma2)) return VILBA_ERR_ARG;
    for (int k = 0; k < w->n_kf; ++k)
        if (!(w->kf_flags[k] & VILBA_KF_FIXED) && !(w->kf_flags[k] & VILBA_KF_HAS_BIAS)) return VILBA_ERR_ARG;
    for (int e = 0; e < w->n_imu; ++e) {
        int i = w->imu_kf_i[e], j = w->imu_kf_j[e];
        if (i < 0 || j < 0 || i >= w->n_kf || j >= w->n_kf) return VILBA_ERR_ARG;
        if (!(w->kf_flags[i] & VILBA_KF_HAS_BIAS) || !(w->kf_flags[j] & VILBA_KF_HAS_BIAS)) return VILBA_ERR_ARG;
    }
    if (w->n_pts && (w->pt_obs_begin[0] != 0 || w->pt_obs_begin[w->n_pts] != w->n_obs)) return VILBA_ERR_ARG;
    for (int e = 0; e < w->n_obs; ++e)
        if (w->obs_kf[e] < 0 || w->obs_kf[e] >= w->n_kf) return VILBA_ERR_ARG;
    return VILBA_OK;
}

int oracle_local_ba(const vilba_window* win, const vilba_params* params, vilba_result* out,
                    const volatile uint8_t* stop_flag) {
    if (!out) return VILBA_ERR_ARG;
    int st = check_window(win);
    out->status = st;
    out->n_trace = 0;
    out->stage2_ran = 0;
    out->n_outliers_stage1 = 0;
    out->solve_ms = 0.0;
    if (st != VILBA_OK) return st;
    vilba_params prm;
    if (params)
        prm = *params;
    else
        vilba_default_params_oracle(&prm);
    // GlobalBundleAdjustmentNavState (Optimizer.cpp:1392-1668) is the same graph with one optimize(nIterations):
    // no early return on the stop flag (g2o then runs zero iterations), no cull, no second stage
    const bool single_stage = (prm.mode & VILBA_MODE_SINGLE_STAGE) != 0;
    if (!single_stage && stop_flag && *stop_flag) {  // Optimizer.cpp:2643-2645: return before optimising
        out->status = VILBA_ABORTED;
        return VILBA_ABORTED;
    }
    Problem pb(*win, prm, stop_flag);
    if (prm.mode & VILBA_MODE_MONO_NOT_ROBUST)  // bRobust == false: no kernel on the mono edges (:1590-1595)
        std::fill(pb.mono_robust.begin(), pb.mono_robust.end(), 0);
    auto t0 = std::chrono::steady_clock::now();
    // stage 1: optimizer.initializeOptimization(); optimizer.optimize(5);   (Optimizer.cpp:2647-2648, :1621-1624)
    int n_active = pb.initialize_optimization();
    pb.optimize(prm.iters_stage1, 1, n_active, out);
    bool do_more = !single_stage && !(stop_flag && *stop_flag);  // :2650-2654
    if (do_more) {
        // :2659-2673  cull + drop the robust kernel of every mono edge
        for (int p_ = 0; p_ < pb.P; ++p_)
            for (int e = win->pt_obs_begin[p_]; e < win->pt_obs_begin[p_ + 1]; ++e) {
                if (pb.chi2_mono(e) > prm.chi2_gate ||
                    !mono_depth_positive(pb.ns[win->obs_kf[e]], pb.pts[p_], pb.calib)) {
                    pb.mono_level[e] = 1;
                    out->n_outliers_stage1++;
                }
                pb.mono_robust[e] = 0;
            }
        n_active = pb.initialize_optimization();
        pb.optimize(prm.iters_stage2, 2, n_active, out);
        out->stage2_ran = 1;
    }
    // :2680-2701 final outlier flags from the cached (possibly stale) errors and the final estimates
    for (int p_ = 0; p_ < pb.P; ++p_)
        for (int e = win->pt_obs_begin[p_]; e < win->pt_obs_begin[p_ + 1]; ++e) {
            double c2 = pb.chi2_mono(e);
            bool bad = c2 > prm.chi2_gate || !mono_depth_positive(pb.ns[win->obs_kf[e]], pb.pts[p_], pb.calib);
            if (out->obs_outlier) out->obs_outlier[e] = bad ? 1 : 0;
            if (out->obs_chi2) out->obs_chi2[e] = c2;
        }
    auto t1 = std::chrono::steady_clock::now();
    out->solve_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    if (out->kf_state)
        for (int k = 0; k < pb.K; ++k) pb.ns[k].store(out->kf_state + VILBA_NS_DOUBLES * k);
    if (out->pt_xyz)
        for (int p_ = 0; p_ < pb.P; ++p_) pb.pts[p_].store(out->pt_xyz + 3 * p_);
    out->status = VILBA_OK;
    return VILBA_OK;
}

int oracle_preintegrate_batch(const vilba_params* params, int32_t n_pairs, const int32_t* sample_begin,
                              const double* gyro, const double* acc, const double* dt, const double* bg,
                              const double* ba, double* out) {
    if (n_pairs < 0 || (n_pairs && (!sample_begin || !gyro || !acc || !dt || !bg || !ba || !out)))
        return VILBA_ERR_ARG;
    vilba_params prm;
    if (params)
        prm = *params;
    else
        vilba_default_params_oracle(&prm);
    NoiseModel nm = {prm.gyr_meas_cov, prm.acc_meas_cov};
    for (int p = 0; p < n_pairs; ++p) {
        Preintegrator pi;  // reset()
        Vec3 bgp = Vec3::from(bg + 3 * p), bap = Vec3::from(ba + 3 * p);
        for (int s = sample_begin[p]; s < sample_begin[p + 1]; ++s)
            pi.update(Vec3::from(gyro + 3 * s) - bgp, Vec3::from(acc + 3 * s) - bap, dt[s], nm);
        pi.store(out + (size_t)VILBA_PREINT_DOUBLES * p);
    }
    return VILBA_OK;
}

void oracle_so3_exp(const double w[3], double q[4]) {
    SO3 s = so3_exp(Vec3::from(w));
    q[0] = s.q.w;
    q[1] = s.q.x;
    q[2] = s.q.y;
    q[3] = s.q.z;
}
void oracle_so3_log(const double q[4], double w[3]) {
    SO3 s;
    s.q.w = q[0];
    s.q.x = q[1];
    s.q.y = q[2];
    s.q.z = q[3];
    so3_log(s).store(w);
}
void oracle_quat_to_matrix(const double q[4], double R[9]) {
    Quat qq = {q[0], q[1], q[2], q[3]};
    quat_to_matrix(qq).store(R);
}
void oracle_matrix_to_quat(const double R[9], double q[4]) {
    Quat qq = quat_from_matrix(Mat3::from(R));
    q[0] = qq.w;
    q[1] = qq.x;
    q[2] = qq.y;
    q[3] = qq.z;
}
void oracle_jacobian_r(const double w[3], double J[9]) { jacobian_r(Vec3::from(w)).store(J); }
void oracle_jacobian_r_inv(const double w[3], double J[9]) { jacobian_r_inv(Vec3::from(w)).store(J); }
void oracle_inverse9(const double A[81], double inv[81]) {
    Mat9 o;
    inverse_lu<9>(Mat9::from(A), o);
    o.store(inv);
}
void oracle_navstate_oplus_pvr(double ns[22], const double d[9]) {
    NavState s;
    s.load(ns);
    s.IncSmallPVR(d);
    s.store(ns);
}
void oracle_navstate_oplus_bias(double ns[22], const double d[6]) {
    NavState s;
    s.load(ns);
    s.IncSmallBias(d);
    s.store(ns);
}

static Calib calib_from(const double c[16]) {
    Calib k;
    k.fx = c[0];
    k.fy = c[1];
    k.cx = c[2];
    k.cy = c[3];
    k.Rbc = Mat3::from(c + 4);
    k.Pbc = Vec3::from(c + 13);
    return k;
}

void oracle_mono_edge(const double ns[22], const double pw[3], const double calib[16], const double uv[2],
                      double err[2], double Jpoint[6], double Jpvr[18], int* depth_positive) {
    NavState s;
    s.load(ns);
    Calib k = calib_from(calib);
    Vec3 p = Vec3::from(pw);
    if (err) mono_error(s, p, k, uv[0], uv[1], err);
    if (Jpoint || Jpvr) {
        Mat<2, 3> A;
        Mat<2, 9> B;
        mono_linearize(s, p, k, A, B);
        if (Jpoint) A.store(Jpoint);
        if (Jpvr) B.store(Jpvr);
    }
    if (depth_positive) *depth_positive = mono_depth_positive(s, p, k) ? 1 : 0;
}

void oracle_pvr_edge(const double ns_i[22], const double ns_j[22], const double ns_bias_i[22],
                     const double preint[142], const double g[3], double err[9], double Ji[81], double Jj[81],
                     double Jb[54]) {
    NavState si, sj, sb;
    si.load(ns_i);
    sj.load(ns_j);
    sb.load(ns_bias_i);
    Preintegrator M;
    M.load(preint);
    Vec3 gv = Vec3::from(g);
    Mat<9, 1> e = pvr_error(si, sj, sb, M, gv);
    if (err) e.store(err);
    if (Ji || Jj || Jb) {
        Mat9 A, B;
        Mat<9, 6> C;
        pvr_linearize(si, sj, sb, M, gv, e, A, B, C);
        if (Ji) A.store(Ji);
        if (Jj) B.store(Jj);
        if (Jb) C.store(Jb);
    }
}

void oracle_bias_edge(const double ns_i[22], const double ns_j[22], double err[6]) {
    NavState si, sj;
    si.load(ns_i);
    sj.load(ns_j);
    bias_error(si, sj).store(err);
}

void oracle_huber(double e2, double delta, double rho[3]) { huber(e2, delta, rho); }

int oracle_debug_system(const vilba_window* win, const vilba_params* params, int robust_mono, double lambda,
                        double* Hpp, double* bp, double* Hll, double* bl, double* Hpl, double* S, double* bs,
                        double* x, double* chi2, double* obs_chi2) {
    int st = check_window(win);
    if (st != VILBA_OK) return st;
    vilba_params prm;
    if (params)
        prm = *params;
    else
        vilba_default_params_oracle(&prm);
    Problem pb(*win, prm, nullptr);
    if (!robust_mono) std::fill(pb.mono_robust.begin(), pb.mono_robust.end(), 0);
    pb.initialize_optimization();
    pb.compute_active_errors();
    if (chi2) chi2[0] = pb.active_robust_chi2();
    if (obs_chi2)
        for (int e = 0; e < pb.E; ++e) obs_chi2[e] = pb.chi2_mono(e);
    pb.build_system();
    const int n = pb.n;
    if (Hpp)
        for (int r = 0; r < n; ++r)
            for (int c = 0; c < n; ++c) {
                // upper blocks were written; mirror to a full symmetric matrix for the caller
                int br = r / 15 * 2 + (r % 15 >= 9), bc = c / 15 * 2 + (c % 15 >= 9);
                double v = (br <= bc) ? pb.Hpp[(size_t)r * n + c] : pb.Hpp[(size_t)c * n + r];
                Hpp[(size_t)r * n + c] = v;
            }
    if (bp) std::memcpy(bp, pb.b.data(), sizeof(double) * n);
    if (Hll) std::memcpy(Hll, pb.Hll.data(), sizeof(double) * 9 * pb.P);
    if (bl) std::memcpy(bl, pb.b.data() + n, sizeof(double) * 3 * pb.P);
    if (Hpl) std::memcpy(Hpl, pb.Hpl.data(), sizeof(double) * 18 * pb.E);
    pb.currentLambda = lambda;
    pb.set_lambda(lambda);
    bool ok = pb.solve_schur();
    pb.restore_diagonal();
    if (S)
        for (int r = 0; r < n; ++r)
            for (int c = 0; c < n; ++c) {
                int br = r / 15 * 2 + (r % 15 >= 9), bc = c / 15 * 2 + (c % 15 >= 9);
                S[(size_t)r * n + c] = (br <= bc) ? pb.S[(size_t)r * n + c] : pb.S[(size_t)c * n + r];
            }
    if (bs) std::memcpy(bs, pb.bs.data(), sizeof(double) * n);
    if (x) std::memcpy(x, pb.x.data(), sizeof(double) * pb.x.size());
    return ok ? n : -100;
}

const char* oracle_build_info(void) {
    return "vilba CPU oracle (dependency-free restatement of the reference g2o path; SO3 / pre-integration / NavState / the three factors / the LM step / the Huber kernel pinned against the compiled reference; g2o's block solver restated) "
#ifdef __VERSION__
           "gcc " __VERSION__
#endif
        ;
}

}  // extern "C"
