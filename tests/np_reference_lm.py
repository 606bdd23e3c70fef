"""A second, independently written restatement of the reference's two-stage Levenberg-Marquardt driver, in numpy,
for SMALL windows (test infrastructure, CPU only).

oracle/lba.cpp restates g2o's machinery the way g2o structures it (block-sparse H, Schur complement, dense LDL^T).
This file restates the same published control flow with none of that machinery: one dense normal-equation matrix
over all poses AND landmarks, numpy.linalg.solve, Python lists for the bookkeeping.  The two only share the
per-edge residual / Jacobian / oplus functions, which tests/test_oracle_edges.py and test_oracle_math.py pin against
closed forms and central differences.  Agreement of the LM traces (trial counts, accept / reject, lambda, chi2, cull,
outlier flags, final states) therefore checks the oracle's assembly, Schur step, LM controller and stale-error
semantics against an implementation that has none of them in common.

Followed: Optimizer.cpp:2643-2701; optimization_algorithm_levenberg.cpp:61-189; sparse_optimizer.cpp:61-114,354-435;
robust_kernel_impl.cpp:78-91; block_solver.hpp:564-604.
"""
import numpy as np

from mc_slam_b200 import capi


def _huber(e, delta):
    d2 = float(np.float32(delta * delta))  # `float dsqr;` in the reference's kernel (robust_kernel_impl.h:84)
    if e <= d2:
        return e, 1.0
    s = np.sqrt(e)
    return 2 * s * delta - d2, delta / s


class DenseLM:
    def __init__(self, oracle, w, prm=None):
        self.o, self.w = oracle, w
        self.prm = prm or capi.default_params()
        self.kf = w.kf_state.copy()
        self.pts = w.pt_xyz.copy()
        self.free = [int(k) for k in np.nonzero((w.kf_flags & capi.KF_FIXED) == 0)[0]]
        self.blk = {k: i for i, k in enumerate(self.free)}
        self.n = 15 * len(self.free)
        self.N = self.n + 3 * w.n_pts
        self.calib = oracle.calib_vec(w)
        self.edge_pt = np.repeat(np.arange(w.n_pts), np.diff(w.pt_obs_begin))
        self.level = np.zeros(w.n_obs, int)      # e->setLevel(1) on culled mono edges
        self.robust = np.ones(w.n_obs, bool)     # mono edges start with their Huber kernel
        self.mono_chi2 = np.zeros(w.n_obs)       # e->chi2() as cached by the last computeActiveErrors
        self.trace = []

    # ---- residuals of the active set, cached like g2o caches _error -------------------------------------
    def compute_active_errors(self):
        w = self.w
        for e in range(w.n_obs):
            if self.level[e]:
                continue  # not in the active set: the cached value stays
            err, _, _, _ = self.o.mono_edge(self.kf[w.obs_kf[e]], self.pts[self.edge_pt[e]], self.calib, w.obs_uv[e].astype(np.float64))
            self.mono_chi2[e] = float(w.obs_inv_sigma2[e]) * float(err @ err)

    def imu_terms(self):
        w, prm = self.w, self.prm
        for e in range(w.n_imu):
            i, j = int(w.imu_kf_i[e]), int(w.imu_kf_j[e])
            M = w.imu_preint[e]
            err, Ji, Jj, Jb = self.o.pvr_edge(self.kf[i], self.kf[j], self.kf[i], M, w.gravity)
            info = np.linalg.inv(M[60:141].reshape(9, 9))
            yield "pvr", e, i, j, err, (Ji, Jj, Jb), info, prm.huber_pvr
            eb = self.o.bias_edge(self.kf[i], self.kf[j])
            infob = np.diag([1 / prm.gyr_bias_rw2] * 3 + [1 / prm.acc_bias_rw2] * 3) / M[141]
            yield "bias", e, i, j, eb, None, infob, prm.huber_bias

    def active_robust_chi2(self):
        chi = 0.0
        for kind, e, i, j, err, J, info, delta in self.imu_terms():
            chi += _huber(float(err @ info @ err), delta)[0]
        for e in range(self.w.n_obs):
            if self.level[e]:
                continue
            c2 = self.mono_chi2[e]
            chi += _huber(c2, self.prm.huber_mono)[0] if self.robust[e] else c2
        return chi

    # ---- dense normal equations over poses and landmarks ------------------------------------------------
    def build_system(self):
        w, n, N = self.w, self.n, self.N
        H, b = np.zeros((N, N)), np.zeros(N)
        for kind, e, i, j, err, J3, info, delta in self.imu_terms():
            rho1 = _huber(float(err @ info @ err), delta)[1]
            J = np.zeros((err.size, N))
            if kind == "pvr":
                Ji, Jj, Jb = J3
                if i in self.blk:
                    J[:, 15 * self.blk[i]:15 * self.blk[i] + 9] = Ji
                    J[:, 15 * self.blk[i] + 9:15 * self.blk[i] + 15] = Jb
                if j in self.blk:
                    J[:, 15 * self.blk[j]:15 * self.blk[j] + 9] = Jj
            else:
                if i in self.blk:
                    J[:, 15 * self.blk[i] + 9:15 * self.blk[i] + 15] = -np.eye(6)
                if j in self.blk:
                    J[:, 15 * self.blk[j] + 9:15 * self.blk[j] + 15] = np.eye(6)
            H += rho1 * J.T @ info @ J
            b += -rho1 * J.T @ info @ err
        for e in range(w.n_obs):
            if self.level[e]:
                continue
            k, p = int(w.obs_kf[e]), int(self.edge_pt[e])
            err, Jp, Jn, _ = self.o.mono_edge(self.kf[k], self.pts[p], self.calib, w.obs_uv[e].astype(np.float64))
            is2 = float(w.obs_inv_sigma2[e])
            rho1 = _huber(is2 * float(err @ err), self.prm.huber_mono)[1] if self.robust[e] else 1.0
            J = np.zeros((2, N))
            J[:, n + 3 * p:n + 3 * p + 3] = Jp
            if k in self.blk:
                J[:, 15 * self.blk[k]:15 * self.blk[k] + 9] = Jn
            H += rho1 * is2 * J.T @ J
            b += -rho1 * is2 * J.T @ err
        return H, b

    def apply(self, x):
        for k, i in self.blk.items():
            self.kf[k] = self.o.oplus_pvr(self.kf[k], x[15 * i:15 * i + 9])
            self.kf[k] = self.o.oplus_bias(self.kf[k], x[15 * i + 9:15 * i + 15])
        self.pts += x[self.n:].reshape(-1, 3)

    # ---- OptimizationAlgorithmLevenberg::solve ---------------------------------------------------------------
    def lm_solve(self, iteration, st):
        prm = self.prm
        self.compute_active_errors()
        current = self.active_robust_chi2()
        ini = current
        H, b = self.build_system()
        if iteration == 0:
            st["lam"] = prm.lm_tau * np.abs(np.diag(H)).max()
            st["ni"] = 2.0
            st["nbad"] = 0
        rec = dict(chi2_initial=ini, lambda_first_trial=st["lam"])
        q = 0
        while True:
            backup = (self.kf.copy(), self.pts.copy())
            lam = st["lam"]
            try:
                x = np.linalg.solve(H + lam * np.eye(self.N), b)
                ok = bool(np.all(np.isfinite(x)))
            except np.linalg.LinAlgError:
                x, ok = np.zeros(self.N), False
            self.apply(x)
            self.compute_active_errors()
            temp = self.active_robust_chi2() if ok else np.finfo(float).max
            rho = (current - temp) / (float(x @ (lam * x + b)) + 1e-3)
            if rho > 0 and np.isfinite(temp):
                alpha = min(1.0 - (2 * rho - 1) ** 3, prm.lm_good_hi)
                st["lam"] = lam * max(prm.lm_good_lo, alpha)
                st["ni"] = 2.0
                current = temp
                accepted = 1
            else:
                st["lam"] = lam * st["ni"]
                st["ni"] *= 2
                self.kf, self.pts = backup  # pop(): estimates restored, the cached edge errors are NOT
                accepted = 0
            q += 1
            if not (rho < 0 and q < prm.max_trials):
                break
        rec.update(trials=q, accepted=accepted, chi2_final=current, lambda_=st["lam"])
        if q == prm.max_trials or rho == 0:
            return 1, rec
        st["nbad"] = st["nbad"] + 1 if (ini - current) * 1e3 < ini else 0
        return (1 if st["nbad"] >= 3 else 0), rec

    def optimize(self, iterations, stage):
        st = {}
        n_active = int((self.level == 0).sum()) + 2 * self.w.n_imu
        for i in range(iterations):
            res, rec = self.lm_solve(i, st)
            rec.update(stage=stage, iteration=i, result=res, n_active_edges=n_active)
            self.trace.append(rec)
            if res != 0:
                break

    def depth_positive(self, e):
        return self.o.mono_edge(self.kf[self.w.obs_kf[e]], self.pts[self.edge_pt[e]], self.calib, self.w.obs_uv[e].astype(np.float64))[3]

    def run(self):
        prm, w = self.prm, self.w
        self.optimize(prm.iters_stage1, 1)
        culled = 0
        for e in range(w.n_obs):  # Optimizer.cpp:2659-2673
            if self.mono_chi2[e] > prm.chi2_gate or not self.depth_positive(e):
                self.level[e] = 1
                culled += 1
            self.robust[e] = False
        self.optimize(prm.iters_stage2, 2)
        outlier = np.array([1 if (self.mono_chi2[e] > prm.chi2_gate or not self.depth_positive(e)) else 0 for e in range(w.n_obs)], np.uint8)
        return dict(trace=self.trace, kf_state=self.kf, pt_xyz=self.pts, obs_outlier=outlier, obs_chi2=self.mono_chi2.copy(),
                    n_outliers_stage1=culled)
