"""T4/T5: the CUDA path, called through the C ABI, against the CPU oracle on identical windows.

Tolerances (BASELINE.json north_star): per-iteration chi2 relative error <= 1e-6, final pose/point
deltas <= 1e-5, identical accept/reject sequence and identical outlier set."""
import numpy as np
import pytest

from mc_slam_b200 import capi, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(vilba):
    c = vilba.Context(0)
    yield c
    c.close()


from parity_util import compare as _compare  # noqa: E402


@pytest.mark.parametrize("name", ["tiny", "small", "c1"])
def test_local_ba_matches_oracle(ctx, oracle, name):
    w = synth.make_config(name)
    _compare(ctx.local_ba(w), oracle.local_ba(w), w)


def test_local_ba_c3_headline_window(ctx, oracle):
    w = synth.make_config("c3")
    r, o = ctx.local_ba(w), oracle.local_ba(w)
    _compare(r, o, w)
    assert len(r.trace) == 15


def test_extra_fixed_keyframes(ctx, oracle):
    w = synth.make_config("small", n_fixed_extra=3)
    r, o = ctx.local_ba(w), oracle.local_ba(w)
    _compare(r, o, w)
    fixed = (w.kf_flags & capi.KF_FIXED) != 0
    assert np.array_equal(r.kf_state[fixed], w.kf_state[fixed])


def test_other_seeds_and_ragged_windows(ctx, oracle):
    for wi in range(1, 4):
        w = synth.make_config("small", window_index=wi, n_fixed_extra=wi % 2, outlier_frac=0.05 * wi)
        _compare(ctx.local_ba(w), oracle.local_ba(w), w)


def test_resident_solve_is_repeatable(ctx, oracle):
    w = synth.make_config("small")
    ctx.upload(w)
    a = ctx.solve_resident()
    da = ctx.download()
    b = ctx.solve_resident()  # restarts from the uploaded state
    db = ctx.download()
    # no floating-point atomics anywhere on the path: the solve is bit-reproducible
    assert a.trace == b.trace
    assert np.array_equal(da.kf_state, db.kf_state) and np.array_equal(da.pt_xyz, db.pt_xyz)
    assert np.array_equal(da.obs_chi2, db.obs_chi2)
    _compare(ctx.download().take_trace(b) if hasattr(capi.Result, "take_trace") else _merge(ctx.download(), b),
             oracle.local_ba(w), w)


def _merge(downloaded, solved):
    downloaded.trace = solved.trace
    downloaded.status = solved.status
    downloaded.stage2_ran = solved.stage2_ran
    downloaded.n_outliers_stage1 = solved.n_outliers_stage1
    return downloaded


def test_stop_flag(ctx, oracle):
    w = synth.make_config("tiny")
    r = ctx.local_ba(w, stop_flag=np.ones(1, np.uint8))
    assert r.status == capi.ABORTED and not r.trace and not r.kf_state.any()
    r = ctx.local_ba(w, stop_flag=np.zeros(1, np.uint8))
    assert r.status == 0 and len(r.trace) > 0


def test_batch_entry(ctx, oracle):
    wins = [synth.make_config("tiny", window_index=i) for i in range(3)]
    rs = ctx.local_ba_batch(wins)
    for w, r in zip(wins, rs):
        _compare(r, oracle.local_ba(w), w)


def test_batch_of_mixed_windows_in_one_launch(ctx, oracle):
    """Windows of different sizes / fixed-key-frame counts / outlier rates solved by ONE batched launch per
    kernel (grid row = window), each with its own device-side LM controller."""
    wins = [synth.make_config("small", window_index=0),
            synth.make_config("tiny", window_index=1),
            synth.make_config("small", window_index=2, n_fixed_extra=2, outlier_frac=0.1),
            synth.make_config("c1", window_index=3),
            synth.make_config("tiny", window_index=4, outlier_frac=0.2)]
    ctx.upload_batch(wins)
    solved = ctx.solve_batch_resident()
    again = ctx.solve_batch_resident()  # restarts from the uploaded state
    down = ctx.download_batch()
    for w, s, a, d in zip(wins, solved, again, down):
        assert [t["trials"] for t in s.trace] == [t["trials"] for t in a.trace]
        _compare(_merge(d, a), oracle.local_ba(w), w)
    # the host-buffer entry point takes the same path
    for w, r in zip(wins, ctx.local_ba_batch(wins)):
        _compare(r, oracle.local_ba(w), w)


def test_batch_larger_than_one_chunk(ctx, oracle, monkeypatch):
    """More windows than the lanes hold at once: several rounds of concurrent lanes."""
    monkeypatch.setenv("VILBA_MAX_BATCH", "3")
    from mc_slam_b200 import api
    c = api.Context(0)
    try:
        # max_batch 3 x 4 lanes = 12 windows per round: 14 windows take two rounds
        wins = [synth.make_config("tiny", window_index=i % 7, outlier_frac=0.03 * (i % 7)) for i in range(14)]
        for w, r in zip(wins, c.local_ba_batch(wins)):
            _compare(r, oracle.local_ba(w), w)
    finally:
        c.close()


def test_invalid_window_is_rejected(ctx, vilba):
    w = synth.make_config("tiny")
    w.obs_kf = w.obs_kf.copy()
    w.obs_kf[0] = 99
    with pytest.raises(vilba.VilbaError):
        ctx.local_ba(w)
    # observations of a point must be ordered by key-frame (MapPoint::GetObservations order)
    w = synth.make_config("tiny")
    w.obs_kf = w.obs_kf.copy()
    b = w.pt_obs_begin
    p = int(np.argmax(np.diff(b) >= 2))
    w.obs_kf[b[p]], w.obs_kf[b[p] + 1] = w.obs_kf[b[p] + 1], w.obs_kf[b[p]]
    with pytest.raises(vilba.VilbaError):
        ctx.local_ba(w)


# ---- entry 2 -------------------------------------------------------------------------------------
def _cmp_preint(got, ref):
    assert np.allclose(got[:, 0:6], ref[:, 0:6], rtol=0, atol=1e-12)  # dP, dV
    assert np.allclose(got[:, 6:15], ref[:, 6:15], rtol=0, atol=1e-12)  # dR
    assert np.allclose(got[:, 15:60], ref[:, 15:60], rtol=1e-10, atol=1e-13)  # bias Jacobians
    # covariance: 1e-10 relative to each pair's largest entry (off-diagonals ~1e-12 of it are cancellation noise)
    scale = np.abs(ref[:, 60:141]).max(axis=1, keepdims=True)
    assert np.all(np.abs(got[:, 60:141] - ref[:, 60:141]) <= 1e-10 * scale + 1e-300)
    assert np.allclose(got[:, 141], ref[:, 141], rtol=1e-14)


def test_preintegrate_batch_c2(ctx, oracle):
    b = synth.make_imu_batch(n_pairs=4096, n_samples=40)
    got = ctx.preintegrate_batch(b.sample_begin, b.gyro, b.acc, b.dt, b.bg, b.ba)
    ref = oracle.preintegrate_batch(b.sample_begin, b.gyro, b.acc, b.dt, b.bg, b.ba)
    _cmp_preint(got, ref)


def test_preintegrate_ragged_leading_partial(ctx, oracle):
    b = synth.make_imu_batch(n_pairs=333, seed=77, ragged=True, leading_partial=True)
    got = ctx.preintegrate_batch(b.sample_begin, b.gyro, b.acc, b.dt, b.bg, b.ba)
    ref = oracle.preintegrate_batch(b.sample_begin, b.gyro, b.acc, b.dt, b.bg, b.ba)
    _cmp_preint(got, ref)


def test_preintegrate_empty_and_single(ctx, oracle):
    assert ctx.preintegrate_batch(np.zeros(1, np.int32), np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0),
                                  np.zeros((0, 3)), np.zeros((0, 3))).shape == (0, 142)
    # a pair with zero samples keeps the reset state (IMUPreintegrator.cpp:39-56)
    sb = np.array([0, 0, 3], np.int32)
    g = np.random.default_rng(0).normal(size=(3, 3))
    a = np.random.default_rng(1).normal(size=(3, 3))
    got = ctx.preintegrate_batch(sb, g, a, np.full(3, 0.005), np.zeros((2, 3)), np.zeros((2, 3)))
    ref = oracle.preintegrate_batch(sb, g, a, np.full(3, 0.005), np.zeros((2, 3)), np.zeros((2, 3)))
    expect0 = np.zeros(142)
    expect0[6:15] = np.eye(3).reshape(-1)
    assert np.array_equal(got[0], expect0)
    _cmp_preint(got, ref)


def test_preintegrate_closed_forms_on_device(ctx):
    S, h = 40, 0.005
    a = np.array([0.3, -1.2, 9.0])
    sb = np.array([0, S], np.int32)
    out = ctx.preintegrate_batch(sb, np.zeros((S, 3)), np.tile(a, (S, 1)), np.full(S, h), np.zeros((1, 3)), np.zeros((1, 3)))[0]
    T = S * h
    assert np.allclose(out[3:6], a * T, rtol=1e-13) and np.allclose(out[0:3], 0.5 * a * T * T, rtol=1e-13)
    assert np.allclose(out[6:15].reshape(3, 3), np.eye(3), atol=1e-15)


def test_local_ba_c4_large_window(ctx, oracle):
    """BASELINE config 4 shape: 100 KF / 50k points / ~600k edges, reduced system n = 1485
    (16-column Cholesky steps, unknowns of the back substitution in shared memory)."""
    w = synth.make_config("c4")
    assert w.n_obs > 550_000 and 15 * w.n_free == 1485
    r, o = ctx.local_ba(w), oracle.local_ba(w)
    _compare(r, o, w)


# ---- edge cases of the window shape ------------------------------------------------------------------
def _subset_points(w, keep):
    """Window with only the map points `keep` (boolean mask), observations re-packed."""
    import dataclasses
    keep = np.asarray(keep, bool)
    cnt = np.diff(w.pt_obs_begin)
    emask = np.repeat(keep, cnt)
    begin = np.zeros(int(keep.sum()) + 1, np.int32)
    np.cumsum(cnt[keep], out=begin[1:])
    return dataclasses.replace(w, pt_xyz=w.pt_xyz[keep].copy(), pt_obs_begin=begin, obs_kf=w.obs_kf[emask].copy(),
                               obs_uv=w.obs_uv[emask].copy(), obs_inv_sigma2=w.obs_inv_sigma2[emask].copy(), truth={})


def test_window_without_map_points_is_imu_only(ctx, oracle):
    """No mono edges at all: the reduced system is the IMU chain alone (no Schur step to speak of)."""
    w = synth.make_config("small")
    w0 = _subset_points(w, np.zeros(w.n_pts, bool))
    assert w0.n_pts == 0 and w0.n_obs == 0
    _compare(ctx.local_ba(w0), oracle.local_ba(w0), w0)


def test_few_points_and_ragged_batch(ctx, oracle):
    """A handful of points (most tiles / CTAs have nothing to do) next to ordinary windows in one batch."""
    w = synth.make_config("small")
    keep = np.zeros(w.n_pts, bool)
    keep[[0, 7, 8, 150, 299]] = True
    wins = [_subset_points(w, keep), synth.make_config("tiny", window_index=2), _subset_points(w, np.zeros(w.n_pts, bool)),
            synth.make_config("small", window_index=5)]
    for wi, r in zip(wins, ctx.local_ba_batch(wins)):
        _compare(r, oracle.local_ba(wi), wi)


@pytest.mark.parametrize("n_kf", [32, 33])
def test_window_at_the_tile_scan_limit(ctx, oracle, n_kf):
    """32 key-frames is the largest window the tile-scan Schur kernel takes (32-bit observer masks); 33 switches
    to the gather over pair lists.  Both against the oracle."""
    w = synth.make_window(n_kf=n_kf, n_pts=600, mean_run=6.0, seed=synth.SEED_BASE + 77)
    assert w.n_kf == n_kf
    _compare(ctx.local_ba(w), oracle.local_ba(w), w)


def test_lane_per_pair_schur_variant(oracle, monkeypatch):
    """The experimental lane-per-pair tile kernel (VILBA_SP_PAIR=1) solves the same windows."""
    monkeypatch.setenv("VILBA_SP_PAIR", "1")
    from mc_slam_b200 import api
    c = api.Context(0)
    try:
        for w in (synth.make_config("small", n_fixed_extra=1), synth.make_config("c1", window_index=2)):
            _compare(c.local_ba(w), oracle.local_ba(w), w)
    finally:
        c.close()


def test_fp64_mma_schur_variant(oracle, monkeypatch):
    """The tile kernel that runs every hit as one m8n8k4 FP64 MMA on Z = W G with D^-1 = G G^T (VILBA_SP_MMA=1):
    single windows, a batch in one launch, extra fixed key-frames, a culled stage 2 and rejected trials."""
    monkeypatch.setenv("VILBA_SP_MMA", "1")
    from mc_slam_b200 import api
    from parity_util import perturbed_window
    c = api.Context(0)
    try:
        for w in (synth.make_config("tiny"), synth.make_config("small", n_fixed_extra=2), synth.make_config("c1", window_index=1),
                  synth.make_config("c3"), perturbed_window("small", scale=1.0)):
            _compare(c.local_ba(w), oracle.local_ba(w), w)
        wins = [synth.make_config("small", window_index=i) for i in range(5)] + [synth.make_config("tiny")]
        for r, w in zip(c.local_ba_batch(wins), wins):
            _compare(r, oracle.local_ba(w), w)
    finally:
        c.close()


def test_rejected_trials_on_the_device(ctx, oracle):
    """Far-off initial estimates make LM reject trials: lambda growth, estimate restore and the stale per-edge chi2 that
    the cull reads must all follow the oracle (and, through tests/test_oracle_vs_dense_lm.py, the dense numpy driver)."""
    from parity_util import perturbed_window
    for name, wi, scale, seed in (("tiny", 0, 2.0, 0), ("tiny", 0, 3.0, 6), ("small", 1, 2.0, 39), ("small", 1, 3.0, 39)):
        w = perturbed_window(name, wi, scale, seed)
        o = oracle.local_ba(w)
        assert any(t["trials"] > 1 for t in o.trace)
        _compare(ctx.local_ba(w), o, w, chi_rtol=1e-6, state_atol=1e-5)


def test_contexts_of_different_window_sizes_do_not_disturb_each_other(vilba, oracle):
    """The opt-in shared-memory limit of a kernel is process-wide, not per context: a context that only ever saw tiny
    windows must not lower it under one that solves large windows (regression: 'eval: invalid argument')."""
    big, small = vilba.Context(0), vilba.Context(0)
    try:
        wb = synth.make_window(n_kf=40, n_pts=500, mean_run=6.0, seed=synth.SEED_BASE + 78)
        ws = synth.make_config("tiny", window_index=7)
        ob, os_ = oracle.local_ba(wb), oracle.local_ba(ws)
        _compare(big.local_ba(wb), ob, wb)
        _compare(small.local_ba(ws), os_, ws)  # configures its kernels for a 4-key-frame window
        _compare(big.local_ba(wb), ob, wb)
        _compare(small.local_ba(ws), os_, ws)
    finally:
        big.close()
        small.close()


def _subset_obs(w, keep):
    """Window with only the observations `keep` (boolean mask over the mono edges); points may end up with 0 or 1."""
    import dataclasses
    keep = np.asarray(keep, bool)
    pt = np.repeat(np.arange(w.n_pts), np.diff(w.pt_obs_begin))
    cnt = np.bincount(pt[keep], minlength=w.n_pts)
    begin = np.zeros(w.n_pts + 1, np.int32)
    np.cumsum(cnt, out=begin[1:])
    return dataclasses.replace(w, pt_obs_begin=begin, obs_kf=w.obs_kf[keep].copy(), obs_uv=w.obs_uv[keep].copy(),
                               obs_inv_sigma2=w.obs_inv_sigma2[keep].copy(), truth={})


def test_irregular_observation_patterns(ctx, oracle):
    """The generator only produces contiguous key-frame runs; real covisibility has gaps.  Random deletions give
    arbitrary observer sets (the popcount edge lookup of the Schur kernel), points with one or no observation."""
    rng = np.random.default_rng(11)
    for name, wi, drop in (("small", 0, 0.35), ("c1", 1, 0.5), ("small", 3, 0.8)):
        w = synth.make_config(name, window_index=wi)
        w2 = _subset_obs(w, rng.uniform(size=w.n_obs) > drop)
        m = np.diff(w2.pt_obs_begin)
        assert (m == 0).any() or drop < 0.5
        assert (m == 1).any()
        _compare(ctx.local_ba(w2), oracle.local_ba(w2), w2)


def test_vision_only_window_without_imu_edges(ctx, oracle):
    """No EdgeNavStatePVR / EdgeNavStateBias at all: the velocity and bias blocks of H_pp are empty, only lambda keeps
    the reduced system positive definite."""
    import dataclasses
    w = synth.make_config("small", window_index=2)
    w0 = dataclasses.replace(w, imu_kf_i=np.zeros(0, np.int32), imu_kf_j=np.zeros(0, np.int32),
                             imu_preint=np.zeros((0, 142)), truth={})
    _compare(ctx.local_ba(w0), oracle.local_ba(w0), w0)


def test_points_seen_only_by_fixed_keyframes_and_heavy_outliers(ctx, oracle):
    w = synth.make_config("small", window_index=6, n_fixed_extra=3, outlier_frac=0.5)
    fixed = (w.kf_flags & capi.KF_FIXED) != 0
    pt = np.repeat(np.arange(w.n_pts), np.diff(w.pt_obs_begin))
    only_fixed = np.zeros(w.n_obs, bool)
    for p in range(0, w.n_pts, 7):  # every 7th point keeps only its observations from fixed key-frames (maybe none)
        only_fixed |= (pt == p) & ~fixed[w.obs_kf]
    w2 = _subset_obs(w, ~only_fixed)
    _compare(ctx.local_ba(w2), oracle.local_ba(w2), w2)


@pytest.mark.parametrize("n", [17, 40])
def test_batches_that_split_over_lanes(ctx, oracle, n):
    """16 or more windows are split over concurrent lanes (2 lanes at 17, 4 at 40), resident and end to end."""
    wins = [synth.make_config("tiny" if i % 3 else "small", window_index=i % 5, outlier_frac=0.02 * (i % 4)) for i in range(n)]
    refs = {}
    ctx.upload_batch(wins)
    assert ctx.batch_groups() == (2 if n == 17 else 4)
    solved = ctx.solve_batch_resident()
    down = ctx.download_batch()
    e2e = ctx.local_ba_batch(wins)
    for i, w in enumerate(wins):
        key = (w.n_pts, i % 5, i % 4)
        if key not in refs:
            refs[key] = oracle.local_ba(w)
        _compare(_merge(down[i], solved[i]), refs[key], w)
        _compare(e2e[i], refs[key], w)


def test_minimal_and_degenerate_windows(ctx, oracle):
    # anchor + ONE free key-frame
    w = synth.make_window(n_kf=2, n_pts=30, mean_run=2.0, seed=synth.SEED_BASE + 90)
    assert w.n_free == 1
    _compare(ctx.local_ba(w), oracle.local_ba(w), w)
    # a free key-frame that observes nothing: only its IMU edges constrain it
    w = synth.make_config("small", window_index=8)
    blind = int(np.nonzero((w.kf_flags & capi.KF_FIXED) == 0)[0][2])
    w2 = _subset_obs(w, w.obs_kf != blind)
    _compare(ctx.local_ba(w2), oracle.local_ba(w2), w2)


def test_many_keyframes_few_points(ctx, oracle):
    """64 key-frames (gather Schur, reduced system n = 945 -> whole-GPU Cholesky) with a sparse map."""
    w = synth.make_window(n_kf=64, n_pts=400, mean_run=5.0, seed=synth.SEED_BASE + 91)
    assert 15 * w.n_free == 945
    _compare(ctx.local_ba(w), oracle.local_ba(w), w)


def test_context_on_the_second_device(vilba, oracle):
    """Lanes run on their own host threads: every entry point has to select the context's device itself."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    c = vilba.Context(1)
    try:
        wins = [synth.make_config("tiny", window_index=i % 4) for i in range(20)]
        refs = [oracle.local_ba(w) for w in wins[:4]]
        c.upload_batch(wins)
        solved, down = c.solve_batch_resident(), c.download_batch()
        for i, w in enumerate(wins):
            _compare(_merge(down[i], solved[i]), refs[i % 4], w)
        for i, r in enumerate(c.local_ba_batch(wins)):
            _compare(r, refs[i % 4], wins[i])
        _compare(c.local_ba(wins[1]), refs[1], wins[1])
    finally:
        c.close()


def test_full_c5_share_in_one_call(ctx, oracle, vilba):
    """BASELINE config 5 at full size for one GPU: 64 independent C3 windows through ONE vilba_local_ba_batch call (four
    concurrent lanes of 16).  Size-independent properties instead of 64 oracle solves: every window's result is what the
    same window gives when it is solved alone (same LM decisions, states to round-off), the order of the windows in the
    batch does not matter, results do not leak between windows (distinct windows -> distinct results); two windows
    are also checked against the oracle."""
    wins = [synth.make_config("c3", window_index=i) for i in range(64)]
    res = ctx.local_ba_batch(wins)
    assert all(r.status == 0 and len(r.trace) == 15 and r.stage2_ran == 1 for r in res)
    alone = vilba.Context(0)
    try:
        for i in (0, 17, 40, 63):
            a = alone.local_ba(wins[i])
            assert [(t["trials"], t["accepted"], t["n_active_edges"]) for t in a.trace] == \
                   [(t["trials"], t["accepted"], t["n_active_edges"]) for t in res[i].trace]
            assert np.abs(a.kf_state - res[i].kf_state).max() < 1e-9 and np.abs(a.pt_xyz - res[i].pt_xyz).max() < 1e-9
            assert np.array_equal(a.obs_outlier, res[i].obs_outlier)
    finally:
        alone.close()
    for i in (5, 58):
        _compare(res[i], oracle.local_ba(wins[i]), wins[i])
    perm = np.random.default_rng(3).permutation(64)
    res_p = ctx.local_ba_batch([wins[j] for j in perm])
    for k, j in enumerate(perm[:16]):
        assert np.abs(res_p[k].kf_state - res[j].kf_state).max() < 1e-9
        assert np.array_equal(res_p[k].obs_outlier, res[j].obs_outlier)
    finals = {round(r.trace[-1]["chi2_final"], 6) for r in res}
    assert len(finals) == 64


def test_solving_the_solution_again_changes_nothing_much(ctx):
    """Idempotence of the solve as a whole: started from its own result, the two-stage schedule ends at a cost that is not
    higher than where the first solve stopped its robust stage, and moves the key-frames by far less than the first solve."""
    w = synth.make_config("c1")
    r1 = ctx.local_ba(w)
    w2 = synth.make_config("c1")
    w2.kf_state = r1.kf_state.copy()
    w2.pt_xyz = r1.pt_xyz.astype(np.float32).astype(np.float64)  # the shim hands points back as float32
    r2 = ctx.local_ba(w2)
    free = (w.kf_flags & capi.KF_FIXED) == 0
    moved1 = np.abs(r1.kf_state[free, 0:3] - w.kf_state[free, 0:3]).max()
    moved2 = np.abs(r2.kf_state[free, 0:3] - w2.kf_state[free, 0:3]).max()
    assert moved2 < 0.1 * moved1 + 1e-4
    assert r2.trace[0]["chi2_initial"] <= 1.05 * r1.trace[4]["chi2_final"]
