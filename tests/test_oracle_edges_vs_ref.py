"""The parity pin of the three factors: the oracle's restatement (oracle/edges.h) against the reference's OWN
EdgeNavStatePVR / EdgeNavStateBias / EdgeNavStatePVRPointXYZ and the vertices' oplusImpl, EXECUTED -- src/IMU/g2otypes.cpp
compiled unmodified against oracle/eigen_stub + a stand-in for the four g2o base-class headers it derives from
(oracle/g2o_stub; oracle/ref_harness_edges.cpp only constructs the reference objects and calls setVertex / setMeasurement /
SetParams / computeError / linearizeOplus).  A line of oracle/edges.h that drifts from g2otypes.cpp:500-788 or
g2otypes.h:616-706 fails here.  Where the compiled reference is absent (the GPU box) the same comparisons run against
tests/golden/ref_edges_v1.npz, produced by that library (tests/golden/make_ref_edges_golden.py).

Still restated and not executed: g2o's optimiser itself (block solver, Levenberg-Marquardt, robust kernel, Schur
complement: Thirdparty/g2o) and the driver in src/Optimizer.cpp -- see DESIGN.md section 2."""
import os

import numpy as np
import pytest

from oracle import pyref

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_edges_v1.npz"))
needs_ref = pytest.mark.skipif(not pyref.available(), reason="neither oracle/_ref nor the reference tree is present")


def close(a, b, rtol, atol):
    return np.allclose(np.asarray(a, float), np.asarray(b, float), rtol=rtol, atol=atol)


def test_pvr_edge_error_and_jacobians(oracle):
    g = G["gravity"]
    for k in range(G["pvr_err"].shape[0]):
        ns_i, ns_j, pre = G["pvr_ns_i"][k], G["pvr_ns_j"][k], G["pvr_preint"][k]
        err, Ji, Jj, Jb = oracle.pvr_edge(ns_i, ns_j, ns_i, pre, g)
        assert close(err, G["pvr_err"][k], 1e-12, 1e-13)
        # the Jacobians contain products of up to four 3x3 factors: summation order is all that may differ
        for got, ref in ((Ji, G["pvr_Ji"][k]), (Jj, G["pvr_Jj"][k]), (Jb, G["pvr_Jb"][k])):
            assert np.abs(got - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
        assert np.abs(G["pvr_Ji"][k]).max() > 0.5 and np.abs(G["pvr_Jb"][k]).max() > 1e-3  # (not vacuous)


def test_bias_edge(oracle):
    for k in range(G["bias_err"].shape[0]):
        assert close(oracle.bias_edge(G["bias_ns_i"][k], G["bias_ns_j"][k]), G["bias_err"][k], 0, 1e-16)
        # oracle/lba.cpp linearises this edge with A = -I, B = +I (g2otypes.cpp:728-734)
        assert np.array_equal(G["bias_Ji"][k], -np.eye(6)) and np.array_equal(G["bias_Jj"][k], np.eye(6))


def test_mono_edge_error_jacobians_and_depth_test(oracle):
    calib = G["calib"]
    seen_behind = 0
    for k in range(G["mono_err"].shape[0]):
        err, Jp, Jn, dpos = oracle.mono_edge(G["mono_ns"][k], G["mono_pw"][k], calib, G["mono_uv"][k])
        ref = G["mono_err"][k]
        assert np.abs(err - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
        for got, r in ((Jp, G["mono_Jp"][k]), (Jn, G["mono_Jn"][k])):
            assert np.abs(got - r).max() <= 1e-12 * max(1.0, np.abs(r).max())
        assert dpos == bool(G["mono_depth"][k])
        seen_behind += not dpos
    assert seen_behind >= 5  # the set contains points behind the camera


def test_vertex_updates(oracle):
    for k in range(G["oplus_ns"].shape[0]):
        ns = G["oplus_ns"][k]
        assert close(oracle.oplus_pvr(ns, G["oplus_d9"][k]), G["oplus_pvr"][k], 0, 5e-16)
        ob = oracle.oplus_bias(ns, G["oplus_d6"][k])
        assert np.array_equal(ob[10:], G["oplus_bias"][k][10:]) and close(ob[:10], G["oplus_bias"][k][:10], 0, 5e-16)


@needs_ref
def test_golden_edge_vectors_are_what_the_compiled_reference_produces():
    """Re-executes the reference on the stored inputs: the committed vectors are its output, not an edited copy."""
    calib = G["calib"]
    for k in (0, 7, 31):
        err, Jp, Jn, dpos = pyref.edge_mono(G["mono_ns"][k], G["mono_pw"][k], calib, G["mono_uv"][k])
        assert np.array_equal(err, G["mono_err"][k]) and np.array_equal(Jn, G["mono_Jn"][k]) and dpos == bool(G["mono_depth"][k])
        err, Ji, Jj = pyref.edge_bias(G["bias_ns_i"][k], G["bias_ns_j"][k])
        assert np.array_equal(err, G["bias_err"][k])
        assert np.array_equal(pyref.vertex_pvr_oplus(G["oplus_ns"][k], G["oplus_d9"][k]), G["oplus_pvr"][k])


@needs_ref
def test_fresh_random_factors_against_the_compiled_reference(oracle):
    """Beyond the committed set: new random states every run of this container, large rotations and tiny ones."""
    from mc_slam_b200 import synth
    rng = np.random.default_rng(77)
    b = synth.make_imu_batch(n_pairs=16, n_samples=25, seed=5)
    g = np.array([0.0, 0.0, -9.81])
    for p in range(16):
        s0, s1 = int(b.sample_begin[p]), int(b.sample_begin[p + 1])
        scale = (1.0, 1e-4, 2.5)[p % 3]
        ns_i = np.concatenate([rng.normal(0, 3, 6), pyref.so3_exp(rng.normal(0, 1, 3) * scale), b.bg[p], b.ba[p], rng.normal(0, 1e-3, 6)])
        ns_j = np.concatenate([rng.normal(0, 3, 6), pyref.so3_exp(rng.normal(0, 1, 3) * scale), b.bg[p], b.ba[p], rng.normal(0, 1e-3, 6)])
        r_err, r_Ji, r_Jj, r_Jb = pyref.edge_pvr(b.gyro[s0:s1], b.acc[s0:s1], b.dt[s0:s1], b.bg[p], b.ba[p], ns_i, ns_j, ns_i, g)
        pre = pyref.preintegrate_batch(np.array([0, s1 - s0]), b.gyro[s0:s1], b.acc[s0:s1], b.dt[s0:s1], b.bg[p:p + 1], b.ba[p:p + 1])[0]
        err, Ji, Jj, Jb = oracle.pvr_edge(ns_i, ns_j, ns_i, pre, g)
        assert np.abs(err - r_err).max() <= 1e-12 * max(1.0, np.abs(r_err).max())
        for got, ref in ((Ji, r_Ji), (Jj, r_Jj), (Jb, r_Jb)):
            assert np.abs(got - ref).max() <= 1e-11 * max(1.0, np.abs(ref).max())
