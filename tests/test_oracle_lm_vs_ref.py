"""The parity pin of the Levenberg-Marquardt control logic and of the robust kernel: g2o's OWN
optimization_algorithm_levenberg.cpp, robust_kernel.cpp and robust_kernel_impl.cpp, compiled unmodified (the headers
they include besides their own class declarations are shadowed by oracle/g2o_lm_stub) and EXECUTED with the oracle's
linear algebra behind the abstract g2o::Solver / g2o::SparseOptimizer faces (oracle/ref_harness_lm.cpp).  The same
windows are solved twice -- once with the oracle's restatement of solve() / computeLambdaInit() / computeScale()
(oracle/lba.cpp: lm_solve), once with the reference's code in its place -- and must give the same trace: trials per
iteration, accept / reject, stop result, lambda, chi2, and the same final estimates and outlier flags.  Both runs share
build_system / solve_schur / the edges, so what is compared is exactly the control flow of
optimization_algorithm_levenberg.cpp:61-189.  Needs the reference tree (this container); skipped elsewhere."""
import numpy as np
import pytest

from mc_slam_b200 import capi, synth
from oracle import pyref
from parity_util import perturbed_window

pytestmark = pytest.mark.skipif(not pyref.available(), reason="neither oracle/_ref nor the reference tree is present")


def same(o, r, w):
    assert o.status == r.status == 0 and o.stage2_ran == r.stage2_ran and o.n_outliers_stage1 == r.n_outliers_stage1
    assert len(o.trace) == len(r.trace)
    for a, b in zip(o.trace, r.trace):
        assert (a["stage"], a["iteration"], a["trials"], a["accepted"], a["result"], a["n_active_edges"]) == (
            b["stage"], b["iteration"], b["trials"], b["accepted"], b["result"], b["n_active_edges"])
        # identical arithmetic on both sides: the values agree exactly, not to a tolerance
        assert a["lambda_"] == b["lambda_"] and a["chi2_initial"] == b["chi2_initial"] and a["chi2_final"] == b["chi2_final"]
    assert np.array_equal(o.kf_state, r.kf_state) and np.array_equal(o.pt_xyz, r.pt_xyz)
    assert np.array_equal(o.obs_outlier, r.obs_outlier) and np.array_equal(o.obs_chi2, r.obs_chi2)


@pytest.mark.parametrize("name", ["tiny", "small", "c1"])
def test_two_stage_local_ba_with_the_reference_lm(oracle, name):
    w = synth.make_config(name)
    same(oracle.local_ba(w), pyref.lm_local_ba(w), w)


def test_rejected_trials_lambda_growth_and_pop(oracle):
    """Far-off initial estimates: LM rejects trials, lambda grows by ni = 2, 4, 8, ..., pop() keeps stale errors."""
    seen = 0
    for name, wi, scale, seed in (("tiny", 0, 2.0, 0), ("tiny", 0, 3.0, 6), ("small", 1, 2.0, 39), ("small", 1, 3.0, 39)):
        w = perturbed_window(name, wi, scale, seed)
        o, r = oracle.local_ba(w), pyref.lm_local_ba(w)
        same(o, r, w)
        seen += sum(t["trials"] > 1 for t in r.trace)
    assert seen >= 4


def test_trial_limit_and_stop_rules(oracle):
    """max_trials = 2 makes solve() return Terminate (qmax == maxTrialsAfterFailure) and optimize() stop early."""
    hit = 0
    for name, wi, scale, seed in (("tiny", 0, 3.0, 6), ("small", 1, 3.0, 39)):
        w = perturbed_window(name, wi, scale, seed)
        p = capi.default_params()
        p.max_trials = 1
        o, r = oracle.local_ba(w, params=p), pyref.lm_local_ba(w, params=p)
        same(o, r, w)
        hit += sum(t["result"] == 1 for t in r.trace)
    assert hit >= 1


def test_global_ba_schedule_and_stop_flag(oracle):
    w = synth.make_window(n_kf=8, n_pts=400, mean_run=5.0, seed=synth.SEED_BASE + 41)
    for robust in (False, True):
        p = capi.global_ba_params(12, robust)
        same(oracle.local_ba(w, params=p), pyref.lm_local_ba(w, params=p), w)
    flag = np.ones(1, np.uint8)
    o, r = oracle.local_ba(w, params=capi.global_ba_params(5, False), stop_flag=flag), \
        pyref.lm_local_ba(w, params=capi.global_ba_params(5, False), stop_flag=flag)
    assert not o.trace and not r.trace and np.array_equal(o.kf_state, r.kf_state)


def test_huber_kernel(oracle):
    """RobustKernelHuber::robustify (robust_kernel_impl.cpp:78-91) against the oracle's huber().  Executing the reference
    here is what showed that its kernel keeps delta^2 in a FLOAT member: rho(e) differs by 5e-9 relative from the
    double-precision closed form, and the inlier threshold sits at float(delta^2)."""
    rng = np.random.default_rng(1)
    for delta in (float(np.float32(np.sqrt(5.991))), float(np.float32(np.sqrt(100 * 21.666))), 1e-3, 50.0):
        for e2 in np.concatenate([rng.uniform(0, 4 * delta * delta, 40), [0.0, delta * delta, np.nextafter(delta * delta, np.inf)]]):
            rho = pyref.huber(e2, delta)
            d2 = float(np.float32(delta * delta))  # robust_kernel_impl.h:84: `float dsqr;` -- delta^2 is rounded to single
            if e2 <= d2:
                assert rho[0] == e2 and rho[1] == 1.0 and rho[2] == 0.0
            else:
                assert np.isclose(rho[0], 2 * np.sqrt(e2) * delta - d2, rtol=1e-14) and np.isclose(rho[1], delta / np.sqrt(e2), rtol=1e-14)
            assert np.array_equal(oracle.huber(e2, delta), rho)
