"""The parity pin: the oracle's restatement of SO3 / IMUPreintegrator / NavState against the reference's OWN sources,
executed.  `make -C oracle ref` compiles /root/reference/src/IMU/{so3,IMUPreintegrator,NavState,imudata}.cpp UNMODIFIED
against the Eigen stand-in of oracle/eigen_stub into oracle/_ref/libref_imu.so (oracle/ref_harness.cpp only calls the
reference's public methods).  A line of oracle/preint.h or oracle/so3.h that drifts from IMUPreintegrator.cpp:63-112 /
so3.cpp:87-300 fails here.  Where neither the compiled reference nor the reference tree exists (the GPU box), the same
comparisons run against tests/golden/ref_imu_v1.npz, which was produced by that library (tests/golden/make_ref_golden.py).

The three factors of src/IMU/g2otypes.cpp are pinned the same way in tests/test_oracle_edges_vs_ref.py.  NOT pinned by
execution: g2o's optimiser and src/Optimizer.cpp (they need Eigen proper + OpenCV + CHOLMOD) stay restated, pinned by the
independent dense LM driver and the properties in tests/test_oracle_solver.py; see DESIGN.md section 2."""
import os

import numpy as np
import pytest

from mc_slam_b200 import capi, synth
from oracle import pyref

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_imu_v1.npz"))
needs_ref = pytest.mark.skipif(not pyref.available(), reason="neither oracle/_ref nor the reference tree is present")


def cmp_preint(got, ref):
    """Tolerances of the GPU parity test (tests/test_gpu_parity.py), tightened where the two CPU paths should agree
    to round-off: the only differences are summation order inside the 9x9 / 3x3 products."""
    assert np.allclose(got[:, 0:15], ref[:, 0:15], rtol=0, atol=1e-13)  # dP, dV, dR
    assert np.allclose(got[:, 15:60], ref[:, 15:60], rtol=1e-11, atol=1e-14)  # bias Jacobians
    scale = np.abs(ref[:, 60:141]).max(axis=1, keepdims=True)
    assert np.all(np.abs(got[:, 60:141] - ref[:, 60:141]) <= 1e-11 * scale + 1e-300)  # covariance
    assert np.allclose(got[:, 141], ref[:, 141], rtol=1e-15)


# ---- against the committed reference-executed vectors (run everywhere) -------------------------------------------
@pytest.mark.parametrize("tag", ["c2", "rg"])
def test_oracle_preintegration_matches_reference_executed_vectors(oracle, tag):
    out = oracle.preintegrate_batch(G[tag + "_sample_begin"], G[tag + "_gyro"], G[tag + "_acc"], G[tag + "_dt"],
                                    G[tag + "_bg"], G[tag + "_ba"])
    cmp_preint(out, G[tag + "_out"])


def test_oracle_so3_matches_reference_executed_vectors(oracle):
    for i, w in enumerate(G["so3_w"]):
        q = oracle.so3_exp(w)
        assert np.allclose(q, G["so3_exp"][i], rtol=0, atol=2e-16), (w, q, G["so3_exp"][i])
        assert np.allclose(oracle.so3_log(q), G["so3_log"][i], rtol=1e-14, atol=1e-18)
        assert np.allclose(oracle.quat_to_matrix(q), G["so3_matrix"][i], rtol=0, atol=5e-16)
        assert np.allclose(oracle.jacobian_r(w), G["jr"][i], rtol=0, atol=1e-15)
        assert np.allclose(oracle.jacobian_r_inv(w), G["jr_inv"][i], rtol=1e-13, atol=1e-15)


def test_default_parameters_are_the_reference_statics():
    p = capi.default_params()
    gyr_cov, acc_cov, gyr_rw2, acc_rw2 = G["imu_constants"]
    assert p.gyr_meas_cov == gyr_cov and p.acc_meas_cov == acc_cov  # imudata.cpp:28-31
    assert p.gyr_bias_rw2 == gyr_rw2 and p.acc_bias_rw2 == acc_rw2  # imudata.cpp:25-26


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["c2", "rg"])
def test_cuda_preintegration_matches_reference_executed_vectors(vilba, tag):
    with vilba.Context(0) as ctx:
        out = ctx.preintegrate_batch(G[tag + "_sample_begin"], G[tag + "_gyro"], G[tag + "_acc"], G[tag + "_dt"],
                                     G[tag + "_bg"], G[tag + "_ba"])
    ref = G[tag + "_out"]
    assert np.allclose(out[:, 0:15], ref[:, 0:15], rtol=0, atol=1e-12)
    assert np.allclose(out[:, 15:60], ref[:, 15:60], rtol=1e-10, atol=1e-13)
    scale = np.abs(ref[:, 60:141]).max(axis=1, keepdims=True)
    assert np.all(np.abs(out[:, 60:141] - ref[:, 60:141]) <= 1e-10 * scale + 1e-300)
    assert np.allclose(out[:, 141], ref[:, 141], rtol=1e-14)


# ---- against the compiled reference itself (this container) -----------------------------------------------------
@needs_ref
def test_golden_vectors_are_what_the_compiled_reference_produces():
    assert "unmodified" in pyref.build_info()
    for tag in ("c2", "rg"):
        out = pyref.preintegrate_batch(G[tag + "_sample_begin"], G[tag + "_gyro"], G[tag + "_acc"], G[tag + "_dt"],
                                       G[tag + "_bg"], G[tag + "_ba"])
        assert np.array_equal(out, G[tag + "_out"])
    assert np.array_equal(pyref.imu_constants(), G["imu_constants"])


@needs_ref
def test_oracle_preintegration_on_the_full_config_2_batch(oracle):
    b = synth.make_imu_batch(n_pairs=4096, n_samples=40)
    cmp_preint(oracle.preintegrate_batch(b.sample_begin, b.gyro, b.acc, b.dt, b.bg, b.ba),
               pyref.preintegrate_batch(b.sample_begin, b.gyro, b.acc, b.dt, b.bg, b.ba))
    b = synth.make_imu_batch(n_pairs=300, ragged=True, leading_partial=True, seed=5)
    cmp_preint(oracle.preintegrate_batch(b.sample_begin, b.gyro, b.acc, b.dt, b.bg, b.ba),
               pyref.preintegrate_batch(b.sample_begin, b.gyro, b.acc, b.dt, b.bg, b.ba))


@needs_ref
def test_oracle_so3_and_navstate_against_the_compiled_reference(oracle):
    rng = np.random.default_rng(3)
    for scale in (1.0, 1e-3, 3e-6, 1e-11, 3.0):
        for _ in range(25):
            w = rng.normal(0, 1, 3) * scale
            q = pyref.so3_exp(w)
            assert np.allclose(oracle.so3_exp(w), q, rtol=0, atol=2e-16)
            assert np.allclose(oracle.so3_log(q), pyref.so3_log(q), rtol=1e-14, atol=1e-18)
            R = pyref.so3_matrix(q)
            assert np.allclose(oracle.quat_to_matrix(q), R, rtol=0, atol=1e-15)  # <= 4 ulp: SO3 copies renormalise
            q2 = pyref.so3_from_matrix(R)  # Quaterniond(Matrix3d) + normalise
            o2 = oracle.matrix_to_quat(R)
            o2 = o2 / np.linalg.norm(o2)
            assert np.allclose(o2, q2, rtol=0, atol=5e-16) or np.allclose(o2, -q2, rtol=0, atol=5e-16)
            assert np.allclose(oracle.jacobian_r(w), pyref.jacobian_r(w), rtol=0, atol=1e-15)
            assert np.array_equal(pyref.jacobian_r(w), pyref.so3_jacobian_r(w))  # the two copies in the reference agree
            if scale < 3.0:
                assert np.allclose(oracle.jacobian_r_inv(w), pyref.jacobian_r_inv(w), rtol=1e-13, atol=1e-15)
    # matrices whose trace is negative: every branch of Eigen's matrix -> quaternion
    for axis in np.eye(3):
        for ang in (3.0, 3.1, np.pi - 1e-6):
            R = pyref.so3_matrix(pyref.so3_exp(axis * ang))
            q2, o2 = pyref.so3_from_matrix(R), oracle.matrix_to_quat(R)
            o2 = o2 / np.linalg.norm(o2)
            assert np.allclose(o2, q2, atol=1e-15) or np.allclose(o2, -q2, atol=1e-15)
    # NavState::IncSmallPVR / IncSmallBias (NavState.cpp:81-109) vs the oracle's oplus
    for _ in range(50):
        ns = np.concatenate([rng.normal(0, 2, 6), pyref.so3_exp(rng.normal(0, 1, 3)), rng.normal(0, 0.01, 12)])
        d9, d6 = rng.normal(0, 0.05, 9), rng.normal(0, 1e-3, 6)
        assert np.allclose(oracle.oplus_pvr(ns, d9), pyref.inc_pvr(ns, d9), rtol=0, atol=5e-16)
        ob, rb = oracle.oplus_bias(ns, d6), pyref.inc_bias(ns, d6)
        assert np.array_equal(ob[10:], rb[10:])  # the increments themselves: exact
        assert np.allclose(ob[:10], rb[:10], rtol=0, atol=5e-16)  # (the harness round-trips q through SO3: renormalised)
