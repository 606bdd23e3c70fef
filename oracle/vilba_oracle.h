/* oracle/vilba_oracle.h -- TEST INFRASTRUCTURE.
 *
 * C entry points of the CPU oracle: a dependency-free restatement of the reference's g2o path
 * for Optimizer::LocalBundleAdjustmentNavState and IMUPreintegrator::update.  It is the checker
 * for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never
 * the product.  PARITY PARTLY PINNED: the reference ships no tests / golden vectors for this path.
 * Pinned against the reference's OWN sources, compiled unmodified and EXECUTED (oracle/Makefile target `ref`,
 * oracle/_ref/; Eigen replaced by the minimal stand-in oracle/eigen_stub, the four g2o base-class headers
 * by oracle/g2o_stub): SO3, IMUPreintegrator, NavState, the IMU constants (tests/test_oracle_vs_ref.py,
 * tests/golden/ref_imu_v1.npz) and the three factors with their Jacobians plus the vertex updates of
 * src/IMU/g2otypes.cpp (tests/test_oracle_edges_vs_ref.py, tests/golden/ref_edges_v1.npz); g2o's Levenberg-
 * Marquardt step (optimization_algorithm_levenberg.cpp) and Huber kernel (robust_kernel_impl.cpp), compiled
 * unmodified with oracle/g2o_lm_stub around them and run on the oracle's linear algebra
 * (oracle/ref_harness_lm.cpp, tests/test_oracle_lm_vs_ref.py).
 * RESTATED AND NOT PINNED BY EXECUTION: g2o's BlockSolver (buildSystem, Schur complement, linear solve:
 * needs Eigen proper), the loop of SparseOptimizer::optimize and the driver in src/Optimizer.cpp (needs
 * g2o + OpenCV + CHOLMOD); see DESIGN.md section 2.
 */
#ifndef VILBA_ORACLE_H
#define VILBA_ORACLE_H
#include "../include/vilba.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Same contract as vilba_local_ba (phases C..E of src/Optimizer.cpp:2643-2701), on the CPU. */
int oracle_local_ba(const vilba_window* win, const vilba_params* params, vilba_result* out,
                    const volatile uint8_t* stop_flag);

/* Same contract as vilba_preintegrate_batch (KeyFrame::ComputePreInt loop), on the CPU. */
int oracle_preintegrate_batch(const vilba_params* params, int32_t n_pairs, const int32_t* sample_begin,
                              const double* gyro, const double* acc, const double* dt,
                              const double* bg, const double* ba, double* out);

/* --- unit-level probes used by the tests ------------------------------------------------------ */
void oracle_so3_exp(const double w[3], double q_wxyz[4]);
void oracle_so3_log(const double q_wxyz[4], double w[3]);
void oracle_quat_to_matrix(const double q_wxyz[4], double R[9]);
void oracle_matrix_to_quat(const double R[9], double q_wxyz[4]);
void oracle_jacobian_r(const double w[3], double J[9]);
void oracle_jacobian_r_inv(const double w[3], double J[9]);
void oracle_inverse9(const double A[81], double inv[81]);
void oracle_navstate_oplus_pvr(double ns[22], const double d[9]);
void oracle_navstate_oplus_bias(double ns[22], const double d[6]);

/* EdgeNavStatePVRPointXYZ: calib = fx,fy,cx,cy,Rbc[9],Pbc[3] (16 doubles) */
void oracle_mono_edge(const double ns[22], const double pw[3], const double calib[16], const double uv[2],
                      double err[2], double Jpoint[6], double Jpvr[18], int* depth_positive);
/* EdgeNavStatePVR: vertices (ns_i, ns_j, bias from ns_bias_i) */
void oracle_pvr_edge(const double ns_i[22], const double ns_j[22], const double ns_bias_i[22],
                     const double preint[142], const double g[3], double err[9], double Ji[81],
                     double Jj[81], double Jb[54]);
void oracle_bias_edge(const double ns_i[22], const double ns_j[22], double err[6]);
/* RobustKernelHuber::robustify with setDelta(delta): rho, rho', rho'' */
void oracle_huber(double e2, double delta, double rho[3]);

/* Builds the normal equations at the window's initial state exactly like BlockSolver::buildSystem
 * and performs ONE Schur solve with the given lambda (block_solver.hpp:354-486).  Output sizes:
 * n = 15 * (#free key-frames); Hpp n*n (symmetric, both triangles filled); bp n; Hll n_pts*9; bl n_pts*3;
 * Hpl n_obs*18 (6x3 rows [P,Phi], zero for observations of fixed key-frames); S n*n; bs n;
 * x n + 3*n_pts; chi2[0] = activeRobustChi2; obs_chi2 n_obs.  Any output pointer may be NULL.
 * robust_mono != 0 keeps the Huber kernel on the mono edges.  Returns n, or < 0 on error. */
int oracle_debug_system(const vilba_window* win, const vilba_params* params, int robust_mono, double lambda,
                        double* Hpp, double* bp, double* Hll, double* bl, double* Hpl, double* S, double* bs,
                        double* x, double* chi2, double* obs_chi2);

const char* oracle_build_info(void);

#ifdef __cplusplus
}
#endif
#endif
