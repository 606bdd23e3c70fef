"""The oracle's whole two-stage LM solve against tests/np_reference_lm.py, an independently written dense numpy
restatement of the same reference control flow (no Schur complement, no block structure, numpy.linalg.solve)."""
import numpy as np
import pytest

from mc_slam_b200 import synth
from np_reference_lm import DenseLM


@pytest.mark.parametrize("kw", [dict(), dict(window_index=3, outlier_frac=0.15), dict(window_index=5, n_fixed_extra=1)])
def test_oracle_trace_matches_the_dense_numpy_driver(oracle, kw):
    w = synth.make_config("tiny", **kw)
    o = oracle.local_ba(w)
    d = DenseLM(oracle, w).run()
    assert len(o.trace) == len(d["trace"])
    for a, b in zip(o.trace, d["trace"]):
        assert (a["stage"], a["iteration"], a["trials"], a["accepted"], a["result"], a["n_active_edges"]) == (
            b["stage"], b["iteration"], b["trials"], b["accepted"], b["result"], b["n_active_edges"])
        assert abs(a["chi2_initial"] - b["chi2_initial"]) <= 1e-8 * abs(b["chi2_initial"])
        assert abs(a["chi2_final"] - b["chi2_final"]) <= 1e-8 * abs(b["chi2_final"])
        assert abs(a["lambda_"] - b["lambda_"]) <= 1e-7 * abs(b["lambda_"])
        assert abs(a["lambda_first_trial"] - b["lambda_first_trial"]) <= 1e-7 * abs(b["lambda_first_trial"])
    assert o.n_outliers_stage1 == d["n_outliers_stage1"]
    assert np.array_equal(o.obs_outlier, d["obs_outlier"])
    assert np.allclose(o.obs_chi2, d["obs_chi2"], rtol=1e-7, atol=1e-10)
    assert np.abs(o.kf_state - d["kf_state"]).max() <= 1e-8
    assert np.abs(o.pt_xyz - d["pt_xyz"]).max() <= 1e-8


def test_rejected_trials_follow_the_same_rules(oracle):
    """Far-off initial estimates: some trials are rejected (lambda *= ni, ni *= 2, pop() with stale edge errors)."""
    from parity_util import perturbed_window
    for scale, seed in ((2.0, 0), (3.0, 6)):  # one and two rejected trials in a row
        _check_rejections(oracle, perturbed_window("tiny", 0, scale, seed))


def _check_rejections(oracle, w):
    o = oracle.local_ba(w)
    assert any(t["trials"] > 1 for t in o.trace), "the fixture is meant to provoke rejected trials"
    d = DenseLM(oracle, w).run()
    assert [t["trials"] for t in o.trace] == [t["trials"] for t in d["trace"]]
    assert [t["accepted"] for t in o.trace] == [t["accepted"] for t in d["trace"]]
    for a, b in zip(o.trace, d["trace"]):
        assert abs(a["chi2_final"] - b["chi2_final"]) <= 1e-7 * abs(b["chi2_final"])
        assert abs(a["lambda_"] - b["lambda_"]) <= 1e-6 * abs(b["lambda_"])
    assert o.n_outliers_stage1 == d["n_outliers_stage1"] and np.array_equal(o.obs_outlier, d["obs_outlier"])
    assert np.abs(o.kf_state - d["kf_state"]).max() <= 1e-6 and np.abs(o.pt_xyz - d["pt_xyz"]).max() <= 1e-6
