"""T1: analytic Jacobians of the three hot edges equal central differences taken through the
vertices' oplus (the scheme of g2o's numeric fallback, base_binary_edge.hpp:131-205,
base_multi_edge.hpp:63-126), and the edge residuals match independent numpy formulas."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from mc_slam_b200 import synth

H = 1e-6


def _window():
    return synth.make_config("small")


def _R(ns):
    return Rotation.from_quat([ns[7], ns[8], ns[9], ns[6]]).as_matrix()


def test_mono_residual_formula(oracle):
    w = _window()
    calib = oracle.calib_vec(w)
    for e in range(0, w.n_obs, 37):
        p = int(np.searchsorted(w.pt_obs_begin, e, side="right") - 1)
        ns = w.kf_state[w.obs_kf[e]]
        err, _, _, dp = oracle.mono_edge(ns, w.pt_xyz[p], calib, w.obs_uv[e].astype(np.float64))
        Rcb = w.Rbc.T
        Pc = Rcb @ _R(ns).T @ (w.pt_xyz[p] - ns[0:3]) - Rcb @ w.Pbc  # g2otypes.h:650-665
        proj = np.array([w.fx * Pc[0] / Pc[2] + w.cx, w.fy * Pc[1] / Pc[2] + w.cy])
        assert np.allclose(err, w.obs_uv[e] - proj, atol=1e-10)
        assert dp == (Pc[2] > 0)


def test_mono_jacobians_numeric(oracle):
    w = _window()
    calib = oracle.calib_vec(w)
    worst = 0.0
    for e in range(0, w.n_obs, 23):
        p = int(np.searchsorted(w.pt_obs_begin, e, side="right") - 1)
        ns = w.kf_state[w.obs_kf[e]].copy()
        pw = w.pt_xyz[p].copy()
        uv = w.obs_uv[e].astype(np.float64)
        _, Jp, Jn, _ = oracle.mono_edge(ns, pw, calib, uv)
        num_p = np.zeros((2, 3))
        for d in range(3):
            dv = np.zeros(3)
            dv[d] = H
            ep = oracle.mono_edge(ns, pw + dv, calib, uv)[0]
            em = oracle.mono_edge(ns, pw - dv, calib, uv)[0]
            num_p[:, d] = (ep - em) / (2 * H)
        num_n = np.zeros((2, 9))
        for d in range(9):
            dv = np.zeros(9)
            dv[d] = H
            ep = oracle.mono_edge(oracle.oplus_pvr(ns, dv), pw, calib, uv)[0]
            em = oracle.mono_edge(oracle.oplus_pvr(ns, -dv), pw, calib, uv)[0]
            num_n[:, d] = (ep - em) / (2 * H)
        scale = max(1.0, np.abs(num_n).max())
        worst = max(worst, np.abs(Jp - num_p).max() / scale, np.abs(Jn - num_n).max() / scale)
        assert np.array_equal(Jn[:, 3:6], np.zeros((2, 3)))  # J_V = 0 (g2otypes.cpp:777)
    assert worst < 1e-6, worst


def _pvr_setup(w, e, rng):
    i, j = int(w.imu_kf_i[e]), int(w.imu_kf_j[e])
    nsi, nsj = w.kf_state[i].copy(), w.kf_state[j].copy()
    nsi[16:22] = 1e-3 * rng.normal(size=6)  # non-zero delta biases exercise the J_*_bias terms
    return nsi, nsj, w.imu_preint[e], w.gravity


def test_pvr_residual_formula(oracle):
    w = _window()
    rng = np.random.default_rng(1)
    for e in range(w.n_imu):
        nsi, nsj, M, g = _pvr_setup(w, e, rng)
        err = oracle.pvr_edge(nsi, nsj, nsi, M, g)[0]
        T = M[141]
        Ri, Rj = _R(nsi), _R(nsj)
        dbg, dba = nsi[16:19], nsi[19:22]
        dP, dV, dR = M[0:3], M[3:6], M[6:15].reshape(3, 3)
        JPg, JPa, JVg, JVa, JRg = (M[15 + 9 * k:24 + 9 * k].reshape(3, 3) for k in range(5))
        rP = Ri.T @ (nsj[0:3] - nsi[0:3] - nsi[3:6] * T - 0.5 * g * T * T) - (dP + JPg @ dbg + JPa @ dba)
        rV = Ri.T @ (nsj[3:6] - nsi[3:6] - g * T) - (dV + JVg @ dbg + JVa @ dba)
        rR = Rotation.from_matrix((dR @ Rotation.from_rotvec(JRg @ dbg).as_matrix()).T @ Ri.T @ Rj).as_rotvec()
        assert np.allclose(err, np.concatenate([rP, rV, rR]), atol=1e-10)  # g2otypes.cpp:564-581


def test_pvr_jacobians_numeric(oracle):
    w = _window()
    rng = np.random.default_rng(2)
    for e in range(w.n_imu):
        nsi, nsj, M, g = _pvr_setup(w, e, rng)
        _, Ji, Jj, Jb = oracle.pvr_edge(nsi, nsj, nsi, M, g)

        def err(di=None, dj=None, db=None):
            a = oracle.oplus_pvr(nsi, di) if di is not None else nsi
            b = oracle.oplus_pvr(nsj, dj) if dj is not None else nsj
            c = oracle.oplus_bias(nsi, db) if db is not None else nsi
            return oracle.pvr_edge(a, b, c, M, g)[0]

        for J, key, dim in ((Ji, "di", 9), (Jj, "dj", 9), (Jb, "db", 6)):
            num = np.zeros((9, dim))
            for d in range(dim):
                dv = np.zeros(dim)
                dv[d] = H
                num[:, d] = (err(**{key: dv}) - err(**{key: -dv})) / (2 * H)
            assert np.allclose(J, num, atol=2e-6 * max(1.0, np.abs(num).max())), (key, np.abs(J - num).max())


def test_bias_residual(oracle):
    rng = np.random.default_rng(3)
    a, b = rng.normal(size=22), rng.normal(size=22)
    e = oracle.bias_edge(a, b)
    assert np.allclose(e[0:3], (b[10:13] + b[16:19]) - (a[10:13] + a[16:19]))
    assert np.allclose(e[3:6], (b[13:16] + b[19:22]) - (a[13:16] + a[19:22]))
