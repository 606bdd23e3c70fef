"""Pins for the CPU oracle's math core (T1 helpers, T2 closed forms).

The reference ships no tests or golden vectors for this path (parity unpinned), so the oracle is
pinned against independent mathematics: scipy/numpy closed forms and finite differences.
"""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from mc_slam_b200 import synth

RNG = np.random.default_rng(7)


def _q_to_scipy(q):  # (w,x,y,z) -> scipy (x,y,z,w)
    return Rotation.from_quat([q[1], q[2], q[3], q[0]])


@pytest.mark.parametrize("scale", [1e-12, 1e-6, 1e-3, 0.3, 2.0, 3.0])
def test_so3_exp_matches_rodrigues(oracle, scale):
    for _ in range(20):
        w = RNG.normal(size=3)
        w *= scale / np.linalg.norm(w)
        q = oracle.so3_exp(w)
        assert abs(np.linalg.norm(q) - 1) < 1e-15
        R = oracle.quat_to_matrix(q)
        assert np.allclose(R, Rotation.from_rotvec(w).as_matrix(), atol=1e-14)


@pytest.mark.parametrize("scale", [1e-11, 1e-6, 1e-2, 1.0, 2.5])
def test_so3_log_inverts_exp(oracle, scale):
    for _ in range(20):
        w = RNG.normal(size=3)
        w *= scale / np.linalg.norm(w)
        w2 = oracle.so3_log(oracle.so3_exp(w))
        assert np.allclose(w, w2, rtol=1e-12, atol=1e-15)


def test_so3_log_uses_atan_n_over_w(oracle):
    # so3.cpp:222 always evaluates 2*atan(n/w)/n (the +-pi branch is dead code): for w<0 the
    # returned vector is the short rotation with flipped sign convention, not 2*atan2(n,w).
    q = np.array([-0.2, 0.5, 0.3, 0.1])
    q /= np.linalg.norm(q)
    n = np.linalg.norm(q[1:])
    expect = 2 * np.arctan(n / q[0]) / n * q[1:]
    assert np.allclose(oracle.so3_log(q), expect, rtol=1e-14)


def test_matrix_quaternion_roundtrip_all_branches(oracle):
    # trace > 0 and the three largest-diagonal branches of Eigen's Quaternion(Matrix3)
    for rv in ([0.1, 0.2, 0.3], [3.0, 0.1, 0.0], [0.1, 3.0, 0.0], [0.0, 0.1, 3.0], [2.2, 2.2, 0.1]):
        R = Rotation.from_rotvec(rv).as_matrix()
        q = oracle.matrix_to_quat(R)
        assert abs(np.linalg.norm(q) - 1) < 1e-14
        assert np.allclose(oracle.quat_to_matrix(q), R, atol=1e-14)


def test_right_jacobian_and_inverse(oracle):
    for scale in (1e-7, 1e-3, 0.5, 2.0):
        w = RNG.normal(size=3)
        w *= scale / np.linalg.norm(w)
        J = oracle.jacobian_r(w)
        Ji = oracle.jacobian_r_inv(w)
        if scale < 1e-5:  # IMUPreintegrator.h:106-109 returns identity below 1e-5
            assert np.array_equal(J, np.eye(3)) and np.array_equal(Ji, np.eye(3))
            continue
        assert np.allclose(J @ Ji, np.eye(3), atol=1e-10)
        # definition: Exp(w + d) ~= Exp(w) Exp(Jr d)
        d = 1e-6 * RNG.normal(size=3)
        lhs = Rotation.from_rotvec(w + d).as_matrix()
        rhs = Rotation.from_rotvec(w).as_matrix() @ Rotation.from_rotvec(J @ d).as_matrix()
        assert np.allclose(lhs, rhs, atol=1e-11)


def test_inverse9_partial_pivot(oracle):
    A = RNG.normal(size=(9, 9))
    A = A @ A.T + 1e-3 * np.eye(9)
    assert np.allclose(oracle.inverse9(A), np.linalg.inv(A), rtol=1e-9, atol=1e-12)
    # needs pivoting: zero leading entry
    B = RNG.normal(size=(9, 9))
    B[0, 0] = 0.0
    assert np.allclose(oracle.inverse9(B) @ B, np.eye(9), atol=1e-10)


def test_navstate_oplus(oracle):
    ns = np.zeros(22)
    ns[0:3] = [1, 2, 3]
    ns[3:6] = [0.1, 0.2, 0.3]
    q = Rotation.from_rotvec([0.3, -0.2, 0.5])
    ns[6:10] = [q.as_quat()[3], *q.as_quat()[:3]]
    ns[10:16] = RNG.normal(size=6)
    d = np.array([0.01, -0.02, 0.03, 0.1, 0.2, -0.1, 0.02, 0.01, -0.03])
    out = oracle.oplus_pvr(ns, d)
    assert np.allclose(out[0:3], ns[0:3] + d[0:3])
    assert np.allclose(out[3:6], ns[3:6] + d[3:6])
    Rn = (q * Rotation.from_rotvec(d[6:9])).as_matrix()  # right-multiplicative update (NavState.cpp:93-95)
    assert np.allclose(oracle.quat_to_matrix(out[6:10]), Rn, atol=1e-14)
    assert np.array_equal(out[10:22], ns[10:22])
    out2 = oracle.oplus_bias(ns, np.arange(6) * 1e-3)
    assert np.allclose(out2[16:22], np.arange(6) * 1e-3)
    assert np.array_equal(out2[0:16], ns[0:16])  # base biases never change (NavState.cpp:100-109)


# ---------------------------------------------------------------------------------------------
# T2: pre-integration closed forms
# ---------------------------------------------------------------------------------------------
def _integrate(oracle, gyro, acc, dt, bg=None, ba=None):
    S = len(dt)
    bg = np.zeros(3) if bg is None else bg
    ba = np.zeros(3) if ba is None else ba
    return oracle.preintegrate_batch(np.array([0, S], np.int32), np.asarray(gyro), np.asarray(acc), np.asarray(dt),
                                     bg, ba)[0]


def test_preint_reset_state(oracle):
    out = oracle.preintegrate_batch(np.array([0, 0], np.int32), np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0),
                                    np.zeros(3), np.zeros(3))[0]
    expect = np.zeros(142)
    expect[6:15] = np.eye(3).reshape(-1)
    assert np.array_equal(out, expect)  # IMUPreintegrator.cpp:39-56


def test_preint_constant_acceleration(oracle):
    S, h = 40, 0.005
    a = np.array([0.3, -1.2, 9.0])
    out = _integrate(oracle, np.zeros((S, 3)), np.tile(a, (S, 1)), np.full(S, h))
    T = S * h
    assert np.allclose(out[3:6], a * T, rtol=1e-13)
    assert np.allclose(out[0:3], 0.5 * a * T * T, rtol=1e-13)
    assert np.allclose(out[6:15].reshape(3, 3), np.eye(3), atol=1e-15)
    assert abs(out[141] - T) < 1e-15
    # J_V_ba = -T I, J_P_ba = -T^2/2 I for zero rotation
    assert np.allclose(out[42:51].reshape(3, 3), -T * np.eye(3), rtol=1e-13)
    assert np.allclose(out[24:33].reshape(3, 3), -0.5 * T * T * np.eye(3), rtol=1e-12)


def test_preint_constant_rate(oracle):
    S, h = 40, 0.005
    w = np.array([0.4, -0.3, 0.9])
    out = _integrate(oracle, np.tile(w, (S, 1)), np.zeros((S, 3)), np.full(S, h))
    assert np.allclose(out[6:15].reshape(3, 3), Rotation.from_rotvec(w * S * h).as_matrix(), atol=1e-13)
    assert np.allclose(out[0:6], 0.0)


def test_preint_covariance_symmetric_psd_and_growing(oracle):
    b = synth.make_imu_batch(n_pairs=8, n_samples=40, seed=5)
    prev = None
    for S in (10, 20, 40):
        out = oracle.preintegrate_batch(np.array([0, S], np.int32), b.gyro[:40], b.acc[:40], b.dt[:40], b.bg[0], b.ba[0])[0]
        cov = out[60:141].reshape(9, 9)
        assert np.allclose(cov, cov.T, rtol=1e-12, atol=1e-20)
        ev = np.linalg.eigvalsh(0.5 * (cov + cov.T))
        assert ev.min() > 0
        if prev is not None:
            assert np.all(np.diag(cov) > np.diag(prev))
        prev = cov


def test_preint_bias_jacobians_predict_reintegration(oracle):
    b = synth.make_imu_batch(n_pairs=4, n_samples=40, seed=9)
    sb = b.sample_begin
    for p in range(4):
        sl = slice(sb[p], sb[p + 1])
        base = _integrate(oracle, b.gyro[sl], b.acc[sl], b.dt[sl], b.bg[p], b.ba[p])
        dbg = 1e-4 * RNG.normal(size=3)
        dba = 1e-4 * RNG.normal(size=3)
        pert = _integrate(oracle, b.gyro[sl], b.acc[sl], b.dt[sl], b.bg[p] + dbg, b.ba[p] + dba)
        JPg, JPa = base[15:24].reshape(3, 3), base[24:33].reshape(3, 3)
        JVg, JVa = base[33:42].reshape(3, 3), base[42:51].reshape(3, 3)
        JRg = base[51:60].reshape(3, 3)
        assert np.allclose(pert[0:3], base[0:3] + JPg @ dbg + JPa @ dba, atol=5e-10)
        assert np.allclose(pert[3:6], base[3:6] + JVg @ dbg + JVa @ dba, atol=5e-9)
        Rp = base[6:15].reshape(3, 3) @ Rotation.from_rotvec(JRg @ dbg).as_matrix()
        assert np.allclose(pert[6:15].reshape(3, 3), Rp, atol=5e-10)


def test_preint_numpy_generator_agrees_with_oracle(oracle):
    """The generator's numpy restatement and the C++ oracle are independent implementations."""
    b = synth.make_imu_batch(n_pairs=16, n_samples=40, seed=3)
    ref = oracle.preintegrate_batch(b.sample_begin, b.gyro, b.acc, b.dt, b.bg, b.ba)
    mine = synth.preintegrate_numpy(b.gyro.reshape(16, 40, 3), b.acc.reshape(16, 40, 3), b.dt.reshape(16, 40), b.bg, b.ba)
    assert np.allclose(mine[:, :60], ref[:, :60], rtol=1e-10, atol=1e-13)
    assert np.allclose(mine[:, 60:141], ref[:, 60:141], rtol=1e-9, atol=1e-22)
    assert np.allclose(mine[:, 141], ref[:, 141], rtol=1e-14)


def test_preint_ragged_and_leading_partial(oracle):
    b = synth.make_imu_batch(n_pairs=12, seed=4, ragged=True, leading_partial=True)
    out = oracle.preintegrate_batch(b.sample_begin, b.gyro, b.acc, b.dt, b.bg, b.ba)
    counts = np.diff(b.sample_begin)
    assert counts.min() >= 11 and counts.max() <= 101
    for p in range(12):  # delta_time is the plain sum of the dt's, leading partial interval included
        assert abs(out[p, 141] - b.dt[b.sample_begin[p]:b.sample_begin[p + 1]].sum()) < 1e-13
        one = _integrate(oracle, b.gyro[b.sample_begin[p]:b.sample_begin[p + 1]], b.acc[b.sample_begin[p]:b.sample_begin[p + 1]],
                         b.dt[b.sample_begin[p]:b.sample_begin[p + 1]], b.bg[p], b.ba[p])
        assert np.array_equal(one, out[p])
