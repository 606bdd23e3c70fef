"""Optimizer::GlobalBundleAdjustmentNavState (src/Optimizer.cpp:1392-1668) through
vilba_global_ba against the CPU oracle run with the same settings (one optimize(n), mnId 0 fixed, Huber kernels on
every edge only when bRobust).  Same tolerances as the local BA (BASELINE.json north_star)."""
import ctypes as C

import numpy as np
import pytest

from mc_slam_b200 import capi, synth
from parity_util import compare, perturbed_window


def _map_window(n_kf, n_pts, seed, mean_run=6.0, **kw):
    # the whole map as one window: key-frame 0 is the fixed origin, every other key-frame is free
    return synth.make_window(n_kf=n_kf, n_pts=n_pts, mean_run=mean_run, seed=synth.SEED_BASE + seed, **kw)


def test_global_params_follow_the_reference_literals():
    p = capi.global_ba_params(20, True)
    assert p.mode == capi.MODE_SINGLE_STAGE and p.iters_stage1 == 20 and p.iters_stage2 == 0
    assert p.huber_pvr == float(np.float32(np.sqrt(21.666))) and p.huber_bias == float(np.float32(np.sqrt(16.812)))
    assert p.huber_mono == float(np.float32(np.sqrt(5.99)))
    q = capi.global_ba_params(10, False)
    assert q.mode == capi.MODE_SINGLE_STAGE | capi.MODE_MONO_NOT_ROBUST
    assert np.isinf(q.huber_pvr) and np.isinf(q.huber_bias)
    assert q.chi2_gate == capi.default_params().chi2_gate and q.max_trials == 10


@pytest.mark.parametrize("robust", [False, True])
def test_oracle_single_stage_schedule(oracle, robust):
    w = _map_window(6, 200, 21)
    o = oracle.local_ba(w, params=capi.global_ba_params(7, robust))
    assert o.status == 0 and o.stage2_ran == 0 and o.n_outliers_stage1 == 0
    assert 0 < len(o.trace) <= 7 and all(t["stage"] == 1 for t in o.trace)
    assert o.trace[-1]["chi2_final"] < o.trace[0]["chi2_initial"]
    # a non-robust solve sees the full quadratic cost of the 2 % gross outliers the generator plants
    if not robust:
        r = oracle.local_ba(w, params=capi.global_ba_params(7, True))
        assert o.trace[0]["chi2_initial"] > r.trace[0]["chi2_initial"]


def test_oracle_stop_flag_set_at_entry_runs_zero_iterations(oracle):
    w = _map_window(5, 100, 22)
    o = oracle.local_ba(w, params=capi.global_ba_params(10, False), stop_flag=np.ones(1, np.uint8))
    assert o.status == 0 and not o.trace
    assert np.array_equal(o.kf_state, w.kf_state) and np.array_equal(o.pt_xyz, w.pt_xyz)


def test_library_and_python_agree_on_the_parameters(vilba):
    lib = capi.load_library()
    for n_it, robust in ((10, 0), (20, 1)):
        got = capi.Params()
        lib.vilba_global_ba_params(None, n_it, robust, C.byref(got))
        want = capi.global_ba_params(n_it, bool(robust))
        assert bytes(got) == bytes(want)


@pytest.fixture(scope="module")
def ctx(vilba):
    c = vilba.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("robust", [False, True])
@pytest.mark.parametrize("shape", [(4, 60, 31), (12, 1500, 32), (40, 6000, 33)])
def test_global_ba_matches_oracle(ctx, oracle, shape, robust):
    w = _map_window(*shape)
    r = ctx.global_ba(w, n_iterations=10, robust=robust)
    o = oracle.local_ba(w, params=capi.global_ba_params(10, robust))
    compare(r, o, w)
    assert r.stage2_ran == 0 and 0 < len(r.trace) <= 10


@pytest.mark.gpu
def test_global_ba_twenty_iterations_with_rejected_trials(ctx, oracle):
    # the callers of the global BA family pass nIterations = 10, bRobust = false (src/LocalMapping.cpp:773,
    # src/LoopClosing.cpp:813, both on the ...PRV twin of this function); 20 here to reach rejected trials
    w = perturbed_window("small", scale=1.0)
    r = ctx.global_ba(w, n_iterations=20, robust=False)
    o = oracle.local_ba(w, params=capi.global_ba_params(20, False))
    compare(r, o, w)
    assert any(t["trials"] > 1 for t in r.trace)


@pytest.mark.gpu
def test_global_ba_large_map(ctx, oracle):
    # 120 free key-frames: a 1785 x 1785 reduced system, solved by the multi-kernel factorisation
    w = _map_window(121, 20000, 34, mean_run=10.0)
    r = ctx.global_ba(w, n_iterations=4, robust=True)
    o = oracle.local_ba(w, params=capi.global_ba_params(4, True))
    compare(r, o, w)


@pytest.mark.gpu
def test_global_ba_leaves_the_context_parameters_alone(ctx, oracle):
    w = synth.make_config("small")
    before = ctx.local_ba(w)
    ctx.global_ba(_map_window(5, 100, 35), n_iterations=3, robust=False)
    after = ctx.local_ba(w)
    assert len(after.trace) == len(before.trace) and after.stage2_ran == 1
    assert np.array_equal(after.kf_state, before.kf_state) and np.array_equal(after.obs_outlier, before.obs_outlier)
    compare(after, oracle.local_ba(w), w)


@pytest.mark.gpu
def test_global_ba_stop_flag(ctx, oracle):
    w = _map_window(5, 100, 22)
    r = ctx.global_ba(w, n_iterations=10, robust=False, stop_flag=np.ones(1, np.uint8))
    assert r.status == 0 and not r.trace
    assert np.array_equal(r.kf_state, w.kf_state) and np.array_equal(r.pt_xyz, w.pt_xyz)
    r = ctx.global_ba(w, n_iterations=10, robust=False, stop_flag=np.zeros(1, np.uint8))
    compare(r, oracle.local_ba(w, params=capi.global_ba_params(10, False)), w)


@pytest.mark.gpu
def test_global_ba_zero_iterations(ctx):
    w = _map_window(5, 100, 22)
    r = ctx.global_ba(w, n_iterations=0, robust=True)
    assert r.status == 0 and not r.trace and np.array_equal(r.kf_state, w.kf_state)
