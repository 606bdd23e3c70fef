"""Point-sharded large window (BASELINE config 4): the host-side partition on CPU, the collective
solve on the GPU (world 1 in-process; world 2 through torchrun when the box has two GPUs)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from mc_slam_b200 import sharding, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partition_covers_the_window_and_balances_edges(world):
    w = synth.make_config("c1")
    ranges = [sharding.shard_window(w, r, world) for r in range(world)]
    assert ranges[0][1] == 0 and ranges[-1][2] == w.n_pts
    for (sa, a0, a1, ea0, ea1), (sb, b0, b1, eb0, eb1) in zip(ranges, ranges[1:]):
        assert a1 == b0 and ea1 == eb0
    for sub, p0, p1, e0, e1 in ranges:
        assert sub.n_kf == w.n_kf and sub.n_imu == w.n_imu and sub.n_pts == p1 - p0 and sub.n_obs == e1 - e0
        assert sub.pt_obs_begin[0] == 0 and sub.pt_obs_begin[-1] == sub.n_obs
        assert np.array_equal(sub.obs_kf, w.obs_kf[e0:e1])
        assert abs((e1 - e0) - w.n_obs / world) <= 32  # balanced by edge count up to one point


def test_partition_of_more_ranks_than_points():
    w = synth.make_config("tiny")
    w2 = sharding.shard_window(w, 0, 1)[0]
    assert w2.n_pts == w.n_pts
    tot = 0
    for r in range(64):
        sub, p0, p1, e0, e1 = sharding.shard_window(w, r, 64)
        tot += sub.n_pts
    assert tot == w.n_pts


def test_merge_is_the_inverse_of_the_partition():
    from mc_slam_b200.capi import Result
    w = synth.make_config("small")
    rng = np.random.default_rng(0)
    full = Result.alloc(w)
    full.kf_state[:] = rng.normal(size=full.kf_state.shape)
    full.pt_xyz[:] = rng.normal(size=full.pt_xyz.shape)
    full.obs_chi2[:] = rng.normal(size=full.obs_chi2.shape)
    full.obs_outlier[:] = rng.integers(0, 2, size=full.obs_outlier.shape)
    parts = []
    for r in range(3):
        sub, p0, p1, e0, e1 = sharding.shard_window(w, r, 3)
        res = Result.alloc(sub)
        res.kf_state[:] = full.kf_state
        res.pt_xyz[:], res.obs_chi2[:], res.obs_outlier[:] = full.pt_xyz[p0:p1], full.obs_chi2[e0:e1], full.obs_outlier[e0:e1]
        res.trace = [dict(trials=1, n_active_edges=(e1 - e0) + (2 * w.n_imu if r == 0 else 0))]
        parts.append((res, p0, p1, e0, e1))
    m = sharding.merge_sharded(w, parts)
    assert np.array_equal(m.pt_xyz, full.pt_xyz) and np.array_equal(m.obs_chi2, full.obs_chi2)
    assert np.array_equal(m.obs_outlier, full.obs_outlier) and np.array_equal(m.kf_state, full.kf_state)
    assert m.trace[0]["n_active_edges"] == w.n_obs + 2 * w.n_imu


def _run_worker(nproc, config, port, extra=()):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "sharded_worker.py"),
           "--config", config, "--check", "--steps", "1", "--warmup", "0", *extra]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert "SHARDED" in p.stdout and '"parity": "ok"' in p.stdout and '"ranks_identical": true' in p.stdout, p.stdout[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("config", ["small", "c1"])
def test_sharded_solve_world_1(config):
    """The collective code path (send buffers, allreduces, NCCL loaded at run time) with a single rank."""
    _run_worker(1, config, 29611)


@pytest.mark.gpu
def test_sharded_context_solves_windows_of_different_shapes():
    """Two windows with different P, E and n_free on ONE communicator context: the captured slot holds the allreduces'
    addresses and counts, so it has to follow the window (a stale graph would reduce the wrong bytes)."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=1", "--master-addr", "127.0.0.1",
           "--master-port", "29614", os.path.join(ROOT, "tests", "sharded_worker.py"), "--config", "c1", "--check",
           "--steps", "1", "--warmup", "0", "--then", "small"]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert "SHARDED_THEN" in p.stdout and p.stdout.count('"parity": "ok"') == 2, p.stdout[-2000:]


@pytest.mark.gpu
def test_sharded_solve_world_2():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run_worker(2, "c1", 29612)


@pytest.mark.gpu
def test_sharded_solve_with_an_empty_shard():
    """One map point on two ranks: rank 1 owns no points and no IMU edges, its partial sums are zeros."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run_worker(2, "tiny", 29613, ("--keep-points", "1"))
