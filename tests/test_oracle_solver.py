"""T3: self-consistency of the oracle's normal equations, Schur complement and LM control
(block_solver.hpp:354-560, optimization_algorithm_levenberg.cpp:61-164)."""
import numpy as np
import pytest

from mc_slam_b200 import capi, synth


def _huber(e, delta):
    d2 = float(np.float32(delta * delta))  # `float dsqr;` in the reference's kernel (robust_kernel_impl.h:84)
    if e <= d2:
        return e, 1.0
    s = np.sqrt(e)
    return 2 * s * delta - d2, delta / s


def _assemble_full(oracle, w, robust_mono=True):
    """Independent numpy assembly of the full (poses + landmarks) Gauss-Newton system from the
    per-edge residuals/Jacobians."""
    prm = capi.default_params()
    free = np.nonzero((w.kf_flags & capi.KF_FIXED) == 0)[0]
    blk = {int(k): i for i, k in enumerate(free)}
    n = 15 * len(free)
    N = n + 3 * w.n_pts
    H = np.zeros((N, N))
    b = np.zeros(N)
    chi = 0.0
    calib = oracle.calib_vec(w)
    for e in range(w.n_imu):
        i, j = int(w.imu_kf_i[e]), int(w.imu_kf_j[e])
        M = w.imu_preint[e]
        err, Ji, Jj, Jb = oracle.pvr_edge(w.kf_state[i], w.kf_state[j], w.kf_state[i], M, w.gravity)
        info = np.linalg.inv(M[60:141].reshape(9, 9))
        c2 = err @ info @ err
        rho0, rho1 = _huber(c2, prm.huber_pvr)
        chi += rho0
        cols, Js = [], []
        if i in blk:
            cols += [np.arange(15 * blk[i], 15 * blk[i] + 9), np.arange(15 * blk[i] + 9, 15 * blk[i] + 15)]
            Js += [Ji, Jb]
        if j in blk:
            cols += [np.arange(15 * blk[j], 15 * blk[j] + 9)]
            Js += [Jj]
        J = np.zeros((9, N))
        for c, Jk in zip(cols, Js):
            J[:, c] = Jk
        H += rho1 * J.T @ info @ J
        b += -rho1 * J.T @ info @ err
        eb = oracle.bias_edge(w.kf_state[i], w.kf_state[j])
        infob = np.diag([1 / prm.gyr_bias_rw2] * 3 + [1 / prm.acc_bias_rw2] * 3) / M[141]
        c2 = eb @ infob @ eb
        rho0, rho1 = _huber(c2, prm.huber_bias)
        chi += rho0
        J = np.zeros((6, N))
        if i in blk:
            J[:, 15 * blk[i] + 9:15 * blk[i] + 15] = -np.eye(6)
        if j in blk:
            J[:, 15 * blk[j] + 9:15 * blk[j] + 15] = np.eye(6)
        H += rho1 * J.T @ infob @ J
        b += -rho1 * J.T @ infob @ eb
    for p in range(w.n_pts):
        for e in range(w.pt_obs_begin[p], w.pt_obs_begin[p + 1]):
            k = int(w.obs_kf[e])
            err, Jp, Jn, _ = oracle.mono_edge(w.kf_state[k], w.pt_xyz[p], calib, w.obs_uv[e].astype(np.float64))
            is2 = float(w.obs_inv_sigma2[e])
            c2 = is2 * err @ err
            rho0, rho1 = _huber(c2, prm.huber_mono) if robust_mono else (c2, 1.0)
            chi += rho0
            J = np.zeros((2, N))
            J[:, n + 3 * p:n + 3 * p + 3] = Jp
            if k in blk:
                J[:, 15 * blk[k]:15 * blk[k] + 9] = Jn
            H += rho1 * is2 * J.T @ J
            b += -rho1 * is2 * J.T @ err
    return H, b, chi, n


@pytest.mark.parametrize("robust", [True, False])
def test_normal_equations_match_independent_assembly(oracle, robust):
    w = synth.make_config("tiny")
    H, b, chi, n = _assemble_full(oracle, w, robust)
    d = oracle.debug_system(w, lam=0.0 + 1.0, robust_mono=robust)
    assert d["n"] == n
    assert np.isclose(d["chi2"][0], chi, rtol=1e-12)
    scale = np.abs(H[:n, :n]).max()
    assert np.allclose(d["Hpp"], H[:n, :n], rtol=1e-9, atol=1e-12 * scale)
    assert np.allclose(d["bp"], b[:n], rtol=1e-9, atol=1e-9 * np.abs(b[:n]).max())
    for p in range(w.n_pts):
        assert np.allclose(d["Hll"][p], H[n + 3 * p:n + 3 * p + 3, n + 3 * p:n + 3 * p + 3], rtol=1e-10)
    assert np.allclose(d["bl"].reshape(-1), b[n:], rtol=1e-9, atol=1e-9)
    free = np.nonzero((w.kf_flags & capi.KF_FIXED) == 0)[0]
    blk = {int(k): i for i, k in enumerate(free)}
    for p in range(w.n_pts):
        for e in range(w.pt_obs_begin[p], w.pt_obs_begin[p + 1]):
            k = int(w.obs_kf[e])
            if k not in blk:
                assert not d["Hpl"][e].any()
                continue
            rows = np.r_[15 * blk[k]:15 * blk[k] + 3, 15 * blk[k] + 6:15 * blk[k] + 9]
            assert np.allclose(d["Hpl"][e], H[rows][:, n + 3 * p:n + 3 * p + 3], rtol=1e-9, atol=1e-9)
            # the V rows of the reference's 9x3 block are exact zeros
            assert not H[15 * blk[k] + 3:15 * blk[k] + 6, n + 3 * p:n + 3 * p + 3].any()


@pytest.mark.parametrize("lam", [1e-3, 10.0, 2.5e5])
def test_schur_solution_equals_full_dense_solve(oracle, lam):
    w = synth.make_config("tiny")
    H, b, _, n = _assemble_full(oracle, w, True)
    x_full = np.linalg.solve(H + lam * np.eye(H.shape[0]), b)
    d = oracle.debug_system(w, lam=lam, robust_mono=True)
    assert np.allclose(d["x"], x_full, rtol=1e-6, atol=1e-9 * np.abs(x_full).max())
    # reduced system itself
    Hpp, Hpl, Hll = H[:n, :n], H[:n, n:], H[n:, n:]
    Dl = Hll + lam * np.eye(Hll.shape[0])
    S = Hpp + lam * np.eye(n) - Hpl @ np.linalg.solve(Dl, Hpl.T)
    bs = b[:n] - Hpl @ np.linalg.solve(Dl, b[n:])
    assert np.allclose(d["S"], S, rtol=1e-8, atol=1e-10 * np.abs(S).max())
    assert np.allclose(d["bs"], bs, rtol=1e-8, atol=1e-9 * np.abs(bs).max())


def test_lm_trace_properties(oracle):
    w = synth.make_config("small")
    r = oracle.local_ba(w)
    assert r.status == 0 and r.stage2_ran == 1
    s1 = [t for t in r.trace if t["stage"] == 1]
    s2 = [t for t in r.trace if t["stage"] == 2]
    assert len(s1) <= 5 and len(s2) <= 10 and [t["iteration"] for t in s1] == list(range(len(s1)))
    for tr in r.trace:
        assert tr["chi2_final"] <= tr["chi2_initial"]  # chi2 never increases over accepted steps
        assert 1 <= tr["trials"] <= 10
        if tr["accepted"] and tr["trials"] == 1:
            # lambda shrinks by a factor in [1/3, 2/3] on a good step (levenberg.cpp:134-139)
            ratio = tr["lambda_"] / tr["lambda_first_trial"]
            assert 1 / 3 - 1e-12 <= ratio <= 2 / 3 + 1e-12
    for a, b_ in zip(s1[:-1], s1[1:]):
        assert a["chi2_final"] == b_["chi2_initial"]
    # lambda is re-initialised at iteration 0 of each optimize() call: tau * max diagonal
    d = oracle.debug_system(w, lam=1.0, robust_mono=True)
    maxdiag = max(np.abs(np.diag(d["Hpp"])).max(), np.abs(np.einsum("pii->pi", d["Hll"])).max())
    assert np.isclose(s1[0]["lambda_first_trial"], 1e-5 * maxdiag, rtol=1e-12)
    assert s2[0]["lambda_first_trial"] > s1[-1]["lambda_"]
    # the culled edges are out of stage 2's active set
    assert s2[0]["n_active_edges"] == s1[0]["n_active_edges"] - r.n_outliers_stage1


def test_outlier_flags_follow_gate(oracle):
    w = synth.make_config("small")
    r = oracle.local_ba(w)
    gate = capi.default_params().chi2_gate
    calib = oracle.calib_vec(w)
    for e in range(w.n_obs):
        p = int(np.searchsorted(w.pt_obs_begin, e, side="right") - 1)
        depth_pos = oracle.mono_edge(r.kf_state[w.obs_kf[e]], r.pt_xyz[p], calib, w.obs_uv[e].astype(np.float64))[3]
        assert bool(r.obs_outlier[e]) == (r.obs_chi2[e] > gate or not depth_pos)
    assert 0 < r.obs_outlier.sum() < w.n_obs // 4


def test_stop_flag_semantics(oracle):
    w = synth.make_config("tiny")
    flag = np.ones(1, np.uint8)
    res = capi.Result.alloc(w)
    r = oracle.local_ba(w, stop_flag=flag)
    assert r.status == capi.ABORTED and not r.trace and not r.kf_state.any()  # nothing written (Optimizer.cpp:2643-2645)
    r0 = oracle.local_ba(w, stop_flag=np.zeros(1, np.uint8))
    r1 = oracle.local_ba(w)
    assert np.array_equal(r0.kf_state, r1.kf_state)


def test_fixed_frames_and_inactive_points_untouched(oracle):
    w = synth.make_config("small", n_fixed_extra=2)
    r = oracle.local_ba(w)
    fixed = (w.kf_flags & capi.KF_FIXED) != 0
    assert fixed.sum() == 3
    assert np.array_equal(r.kf_state[fixed], w.kf_state[fixed])
    assert not np.array_equal(r.kf_state[~fixed], w.kf_state[~fixed])
    # base biases never change inside BA (Optimizer.cpp:2744-2750)
    assert np.array_equal(r.kf_state[:, 10:16], w.kf_state[:, 10:16])


def test_deterministic(oracle):
    w = synth.make_config("small")
    a, b = oracle.local_ba(w), oracle.local_ba(w)
    assert np.array_equal(a.kf_state, b.kf_state) and np.array_equal(a.pt_xyz, b.pt_xyz)
    assert [t["chi2_final"] for t in a.trace] == [t["chi2_final"] for t in b.trace]
