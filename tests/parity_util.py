"""GPU-vs-oracle comparison shared by the parity tests and the sharded worker.

Tolerances (BASELINE.json north_star): per-iteration chi2 relative error <= 1e-6, final pose/point
deltas <= 1e-5, identical accept/reject sequence and identical outlier set."""
import numpy as np


def compare(r, o, w, chi_rtol=1e-6, state_atol=1e-5):
    assert r.status == o.status == 0
    assert r.stage2_ran == o.stage2_ran
    assert r.n_outliers_stage1 == o.n_outliers_stage1
    assert len(r.trace) == len(o.trace)
    for a, b in zip(r.trace, o.trace):
        assert (a["stage"], a["iteration"], a["trials"], a["accepted"], a["result"], a["n_active_edges"]) == (
            b["stage"], b["iteration"], b["trials"], b["accepted"], b["result"], b["n_active_edges"])
        assert abs(a["chi2_initial"] - b["chi2_initial"]) <= chi_rtol * abs(b["chi2_initial"])
        assert abs(a["chi2_final"] - b["chi2_final"]) <= chi_rtol * abs(b["chi2_final"])
        assert abs(a["lambda_"] - b["lambda_"]) <= 1e-6 * abs(b["lambda_"])
    assert np.abs(r.kf_state[:, 0:3] - o.kf_state[:, 0:3]).max() <= state_atol  # P
    assert np.abs(r.kf_state[:, 3:6] - o.kf_state[:, 3:6]).max() <= state_atol  # V
    assert np.abs(r.kf_state[:, 6:10] - o.kf_state[:, 6:10]).max() <= state_atol  # R (quaternion)
    assert np.abs(r.kf_state[:, 16:22] - o.kf_state[:, 16:22]).max() <= state_atol  # dbg, dba
    assert np.array_equal(r.kf_state[:, 10:16], w.kf_state[:, 10:16])  # base biases untouched
    assert np.abs(r.pt_xyz - o.pt_xyz).max(initial=0.0) <= state_atol
    assert np.array_equal(r.obs_outlier, o.obs_outlier)
    assert np.allclose(r.obs_chi2, o.obs_chi2, rtol=1e-6, atol=1e-9)


def perturbed_window(name="tiny", window_index=0, scale=2.0, seed=5, **kw):
    """A synthetic window whose initial estimates are far off (points +- scale m, free key-frame positions +- scale / 2):
    Levenberg-Marquardt then REJECTS trials, which exercises the lambda growth, pop() and stale-error rules."""
    from mc_slam_b200 import capi, synth
    rng = np.random.default_rng(seed)
    w = synth.make_config(name, window_index=window_index, **kw)
    free = (w.kf_flags & capi.KF_FIXED) == 0
    w.pt_xyz = (w.pt_xyz + rng.normal(0, scale, w.pt_xyz.shape)).astype(np.float32).astype(np.float64)
    w.kf_state = w.kf_state.copy()
    w.kf_state[free, 0:3] += rng.normal(0, scale * 0.5, (int(free.sum()), 3))
    return w
