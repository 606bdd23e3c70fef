/*
 * vilba_diag.h -- diagnostic entry points of libvilba.so.  NOT part of the reference boundary (include/vilba.h):
 * they exist so that single kernels of the path can be checked against numpy / timed in isolation by tests/ and
 * tools/ without building a whole window around them.
 */
#ifndef VILBA_DIAG_H
#define VILBA_DIAG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Solve n_windows copies of one dense symmetric system S x = b (S: n x n, full symmetric storage) with a
 * reduced-system kernel of the library (the kernels that replace LinearSolverEigen::solve,
 * g2o/solvers/linear_solver_eigen.h:94-124):
 *   variant 0: (the round-1 cluster kernel, removed: VILBA_ERR_ARG)
 *   variant 1: look-ahead cluster kernel, trailing matrix in shared memory (chol_la.cu)
 *   variant 2: multi-kernel blocked factorisation              (chol_big.cu)
 * `cluster` = CTAs per system (variant 1).  The kernel is launched `reps` times (system restored in
 * between); avg_us receives the mean device time of one launch (CUDA events around the launch only).
 * x_out: n (solution of window 0), fail_out: the kernel's failure flag.  Returns a VILBA_* status. */
int vilba_diag_dense_solve(int32_t device, int32_t n, const double* S, const double* b, int32_t variant, int32_t cluster,
                           int32_t n_windows, int32_t reps, double* x_out, int32_t* fail_out, double* avg_us);

/* 1 if variant / cluster can handle a system of dimension n (shared-memory capacity), else 0 */
int vilba_diag_dense_supported(int32_t n, int32_t variant, int32_t cluster);

/* The normal equations of the FIRST Levenberg-Marquardt trial of a window, as the kernels leave them in device memory:
 * upload, initial evaluation, linearise + accumulate (what BlockSolver::buildSystem produces: H_pp, b_p, H_ll, b_l, the
 * H_pl blocks) and the Schur step with the initial lambda (block_solver.hpp:381-439: S = H_pp + lambda I - sum W D^-1 W^T,
 * b_s); the reduced solve is not run.  For tests: the same quantities come out of the oracle for the same lambda.
 *   Hpp n*n (row-major, upper triangle filled), bp n, Hll n_pts*6 (xx xy xz yy yz zz), bl n_pts*3,
 *   W n_obs*18 (6x3 blocks, rows [P, Phi]; zero for observations of fixed key-frames), S n*n (upper triangle), bs n;
 *   n = 15 * (free key-frames).  Any output pointer may be NULL.  `ctx` is a vilba_ctx from vilba_create. */
struct vilba_ctx;
struct vilba_window;
int vilba_diag_first_trial(struct vilba_ctx* ctx, const struct vilba_window* win, double* lambda_out, double* Hpp, double* bp,
                           double* Hll, double* bl, double* W, double* S, double* bs);

#ifdef __cplusplus
}
#endif
#endif /* VILBA_DIAG_H */
