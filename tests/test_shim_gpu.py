"""T5: the reference-shaped C++ interface (mc_slam_b200/shim/vilba_shim.h) round trip.

A C++ driver builds KeyFrame / MapPoint objects, calls Optimizer::LocalBundleAdjustmentNavState and
KeyFrame::ComputePreInt exactly like LocalMapping would, and we compare what it wrote back with the C-ABI
called directly (and with the float32 quantisation rules of Converter::toCvMat / UpdatePoseFromNS)."""
import os
import subprocess

import numpy as np
import pytest

from mc_slam_b200 import capi, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    exe = tmp_path_factory.mktemp("shim") / "shim_driver"
    subprocess.check_call([
        "g++", "-std=c++14", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "shim_driver.cpp"),
        "-o", str(exe), "-L", os.path.join(ROOT, "mc_slam_b200"), "-lvilba", "-Wl,-rpath," + os.path.join(ROOT, "mc_slam_b200")])
    return str(exe)


def _write_window(path, w):
    with open(path, "wb") as f:
        np.array([w.n_kf, w.n_imu, w.n_pts, w.n_obs], np.int32).tofile(f)
        for a in (w.kf_state, w.kf_flags, w.kf_id, w.imu_kf_i, w.imu_kf_j, w.imu_preint, w.pt_xyz, w.pt_obs_begin, w.obs_kf,
                  w.obs_uv, w.obs_inv_sigma2):
            a.tofile(f)
        np.array([w.fx, w.fy, w.cx, w.cy], np.float64).tofile(f)
        w.Rbc.tofile(f)
        w.Pbc.tofile(f)
        w.gravity.tofile(f)


def _read_result(path, w):
    K, P, E = w.n_kf, w.n_pts, w.n_obs
    with open(path, "rb") as f:
        meta = np.fromfile(f, np.int32, 4)
        st = np.fromfile(f, np.float64, 22 * K).reshape(K, 22)
        tcw = np.fromfile(f, np.float32, 16 * K).reshape(K, 4, 4)
        pw = np.fromfile(f, np.float32, 3 * P).reshape(P, 3)
        erased = np.fromfile(f, np.uint8, E)
        chi = np.fromfile(f, np.float64, 64)
    return meta, st, tcw, pw, erased, chi


def test_optimizer_entry_point_round_trip(driver, vilba, tmp_path):
    w = synth.make_config("small", n_fixed_extra=2)
    _write_window(tmp_path / "in.bin", w)
    subprocess.check_call([driver, "lba", str(tmp_path / "in.bin"), str(tmp_path / "out.bin")])
    meta, st, tcw, pw, erased, chi = _read_result(tmp_path / "out.bin", w)
    with vilba.Context(0) as ctx:
        r = ctx.local_ba(w)
    assert meta[0] == 1  # pLM->SetMapUpdateFlagInTracking(true)
    assert meta[1] == len(r.trace) and meta[2] == 1
    assert meta[3] == 1  # UpdateNormalAndDepth called once per local point
    # the shim visits the points in lLocalMapPoints order (first seen by the local key-frames), the direct call
    # in window order: only the summation order differs
    for i, t in enumerate(r.trace):
        assert abs(chi[i] - t["chi2_final"]) <= 1e-9 * abs(t["chi2_final"])
    free = (w.kf_flags & capi.KF_FIXED) == 0
    assert np.abs(st[free][:, :10] - r.kf_state[free][:, :10]).max() < 1e-9
    assert np.abs(st[free][:, 16:] - r.kf_state[free][:, 16:]).max() < 1e-9
    assert np.array_equal(st[~free], w.kf_state[~free])  # fixed key-frames are not written
    assert np.array_equal(st[:, 10:16], w.kf_state[:, 10:16])  # base biases never change
    # points come back through Converter::toCvMat: float32
    assert np.abs(pw - r.pt_xyz.astype(np.float32)).max() <= 2e-6
    # outlier observations are erased both ways (Optimizer.cpp:2704-2715)
    assert np.array_equal(erased, r.obs_outlier)
    # camera pose of the local key-frames follows UpdatePoseFromNS in float32
    from scipy.spatial.transform import Rotation
    for k in np.nonzero(free)[0]:
        q = r.kf_state[k, 6:10]
        Rwb = Rotation.from_quat([q[1], q[2], q[3], q[0]]).as_matrix()
        Rwc = Rwb @ w.Rbc
        Pwc = Rwb @ w.Pbc + r.kf_state[k, 0:3]
        assert np.abs(tcw[k][:3, :3] - Rwc.T).max() < 1e-5
        assert np.abs(tcw[k][:3, 3] + Rwc.T @ Pwc).max() < 1e-4
        assert tcw[k][3, 3] == 1.0


def test_stop_flag_set_before_returns_without_writing(driver, tmp_path):
    w = synth.make_config("tiny")
    _write_window(tmp_path / "in.bin", w)
    subprocess.check_call([driver, "lba", str(tmp_path / "in.bin"), str(tmp_path / "out.bin"), "stop"])
    meta, st, tcw, pw, erased, chi = _read_result(tmp_path / "out.bin", w)
    assert meta[0] == 0 and not erased.any()
    assert np.array_equal(st, w.kf_state)
    assert np.array_equal(pw, w.pt_xyz.astype(np.float32))
    assert not tcw.any()


def test_keyframe_compute_preint(driver, vilba, oracle, tmp_path):
    """KeyFrame::ComputePreInt feeds update() once for the leading partial interval and once per sample
    (src/KeyFrame.cpp:214-241); the shim records the samples and integrates them on the GPU."""
    N, rng = 6, np.random.default_rng(5)
    counts = rng.integers(8, 60, N)
    begin = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    S = int(begin[-1])
    g = rng.normal(0, 0.3, (S, 3))
    a = rng.normal(0, 1.0, (S, 3)) + [0, 0, 9.8]
    bias = rng.normal(0, 0.01, (N, 6))
    tprev = rng.uniform(0, 10, N)
    t = np.zeros(S)
    tcur = np.zeros(N)
    for p in range(N):
        lead = rng.uniform(0.0005, 0.005)
        t[begin[p]:begin[p + 1]] = tprev[p] + lead + 0.005 * np.arange(counts[p])
        tcur[p] = t[begin[p + 1] - 1] + rng.uniform(0.001, 0.005)
    with open(tmp_path / "pin.bin", "wb") as f:
        np.array([N], np.int32).tofile(f)
        for arr in (begin, bias, tprev, tcur, g, a, t):
            np.ascontiguousarray(arr).tofile(f)
    subprocess.check_call([driver, "preint", str(tmp_path / "pin.bin"), str(tmp_path / "pout.bin")])
    out = np.fromfile(tmp_path / "pout.bin", np.float64).reshape(2, N, 142)
    # expected: the same update() sequence through the oracle
    G, A, D, sb = [], [], [], [0]
    for p in range(N):
        sl = slice(begin[p], begin[p + 1])
        tt = t[sl]
        dts = np.concatenate([[tt[0] - tprev[p]], np.diff(tt), [tcur[p] - tt[-1]]])
        G.append(np.concatenate([g[sl][:1], g[sl]]))
        A.append(np.concatenate([a[sl][:1], a[sl]]))
        D.append(dts)
        sb.append(sb[-1] + len(dts))
    ref = oracle.preintegrate_batch(np.array(sb, np.int32), np.concatenate(G), np.concatenate(A), np.concatenate(D),
                                    bias[:, :3], bias[:, 3:])
    for got in out:  # [0] one launch per key-frame, [1] ComputePreIntBatch
        assert np.allclose(got[:, :60], ref[:, :60], rtol=1e-9, atol=1e-12)
        scale = np.abs(ref[:, 60:141]).max(axis=1, keepdims=True)
        assert np.all(np.abs(got[:, 60:141] - ref[:, 60:141]) <= 1e-9 * scale)
        assert np.allclose(got[:, 141], ref[:, 141], rtol=1e-12)


def _read_gba(path, w):
    K, P, E = w.n_kf, w.n_pts, w.n_obs
    with open(path, "rb") as f:
        f.seek(4 * 4 + 8 * 22 * K + 4 * 16 * K + 4 * 3 * P + E + 8 * 64)
        stg = np.fromfile(f, np.float64, 22 * K).reshape(K, 22)
        tcwg = np.fromfile(f, np.float32, 16 * K).reshape(K, 4, 4)
        pg = np.fromfile(f, np.float32, 3 * P).reshape(P, 3)
        tag = np.fromfile(f, np.int64, K + P)
    return stg, tcwg, pg, tag


def _global_map():
    w = synth.make_window(n_kf=8, n_pts=400, mean_run=5.0, seed=synth.SEED_BASE + 41)
    w.kf_id = np.arange(w.n_kf, dtype=np.int64)  # the first key-frame of the map has mnId 0 and is the fixed one
    return w


@pytest.mark.parametrize("robust", [0, 1])
def test_global_ba_entry_point_writes_the_live_state(driver, vilba, tmp_path, robust):
    """Optimizer::GlobalBundleAdjustmentNavState with nLoopKF == 0 (src/Optimizer.cpp:1637-1641,1657-1660)."""
    w = _global_map()
    _write_window(tmp_path / "in.bin", w)
    subprocess.check_call([driver, "gba", str(tmp_path / "in.bin"), str(tmp_path / "out.bin"), "0", str(robust)])
    meta, st, tcw, pw, erased, chi = _read_result(tmp_path / "out.bin", w)
    stg, tcwg, pg, tag = _read_gba(tmp_path / "out.bin", w)
    with vilba.Context(0) as ctx:
        r = ctx.global_ba(w, n_iterations=10, robust=bool(robust))
    assert meta[1] == len(r.trace) and meta[2] == 0 and meta[3] == 1
    for i, t in enumerate(r.trace):
        assert abs(chi[i] - t["chi2_final"]) <= 1e-9 * abs(t["chi2_final"])
    assert np.abs(st[:, :10] - r.kf_state[:, :10]).max() < 1e-9 and np.abs(st[:, 16:] - r.kf_state[:, 16:]).max() < 1e-9
    assert np.array_equal(st[0], w.kf_state[0])  # mnId 0 is fixed
    assert np.array_equal(st[:, 10:16], w.kf_state[:, 10:16])
    assert np.abs(pw - r.pt_xyz.astype(np.float32)).max() <= 2e-6
    assert not erased.any()  # the global BA erases nothing
    assert tcw[:, 3, 3].tolist() == [1.0] * w.n_kf  # UpdatePoseFromNS on every key-frame
    assert not tag.any() and not tcwg.any() and not pg.any()


def test_global_ba_entry_point_beside_the_mapping_thread(driver, vilba, tmp_path):
    """nLoopKF != 0: results go to mNavStateGBA / mTcwGBA / mPosGBA, the live map is untouched (:1643-1665)."""
    w = _global_map()
    _write_window(tmp_path / "in.bin", w)
    subprocess.check_call([driver, "gba", str(tmp_path / "in.bin"), str(tmp_path / "out.bin"), "7", "0"])
    meta, st, tcw, pw, erased, chi = _read_result(tmp_path / "out.bin", w)
    stg, tcwg, pg, tag = _read_gba(tmp_path / "out.bin", w)
    with vilba.Context(0) as ctx:
        r = ctx.global_ba(w, n_iterations=10, robust=False)
    assert np.array_equal(st, w.kf_state) and np.array_equal(pw, w.pt_xyz.astype(np.float32)) and not tcw.any()
    assert meta[3] == 0  # no UpdateNormalAndDepth
    assert (tag == 7).all()
    assert np.abs(stg[:, :10] - r.kf_state[:, :10]).max() < 1e-9 and np.abs(stg[:, 16:] - r.kf_state[:, 16:]).max() < 1e-9
    assert np.array_equal(stg[:, 10:16], w.kf_state[:, 10:16])
    assert np.abs(pg - r.pt_xyz.astype(np.float32)).max() <= 2e-6
    from scipy.spatial.transform import Rotation
    for k in range(w.n_kf):
        q = r.kf_state[k, 6:10]
        Rwc = Rotation.from_quat([q[1], q[2], q[3], q[0]]).as_matrix() @ w.Rbc
        Pwc = Rotation.from_quat([q[1], q[2], q[3], q[0]]).as_matrix() @ w.Pbc + r.kf_state[k, 0:3]
        assert np.abs(tcwg[k][:3, :3] - Rwc.T).max() < 1e-5 and np.abs(tcwg[k][:3, 3] + Rwc.T @ Pwc).max() < 1e-4


def test_global_ba_entry_point_with_the_stop_flag_up(driver, tmp_path):
    """g2o runs zero iterations and the function still writes the (unchanged) estimates back."""
    w = _global_map()
    _write_window(tmp_path / "in.bin", w)
    subprocess.check_call([driver, "gba", str(tmp_path / "in.bin"), str(tmp_path / "out.bin"), "0", "0", "stop"])
    meta, st, tcw, pw, erased, chi = _read_result(tmp_path / "out.bin", w)
    assert meta[1] == 0 and meta[3] == 1
    assert np.abs(st - w.kf_state).max() < 1e-15  # quaternion -> matrix -> quaternion of Set_Rot
    assert np.array_equal(pw, w.pt_xyz.astype(np.float32))
    assert tcw[:, 3, 3].tolist() == [1.0] * w.n_kf
