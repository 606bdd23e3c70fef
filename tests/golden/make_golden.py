#!/usr/bin/env python
"""Generates tests/golden/*.npz from the CPU oracle (the reference ships no golden vectors for this path and
cannot be built here, so these fixtures pin the ORACLE's current behaviour: any later change of the oracle or
of the CUDA path shows up against them).  Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mc_slam_b200 import synth  # noqa: E402
from oracle import pyoracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def window_case(name, **kw):
    w = synth.make_config(name, **kw)
    r = pyoracle.local_ba(w)
    tr = np.array([[t["stage"], t["iteration"], t["trials"], t["accepted"], t["result"], t["n_active_edges"],
                    t["chi2_initial"], t["chi2_final"], t["lambda_"]] for t in r.trace])
    return dict(kf_state_in=w.kf_state, kf_flags=w.kf_flags, pt_in=w.pt_xyz, obs_kf=w.obs_kf, obs_uv=w.obs_uv,
                preint=w.imu_preint, kf_state_out=r.kf_state, pt_out=r.pt_xyz, outlier=r.obs_outlier, chi2=r.obs_chi2,
                trace=tr, n_outliers_stage1=np.array([r.n_outliers_stage1]))


def main():
    np.savez_compressed(os.path.join(HERE, "lba_tiny.npz"), **window_case("tiny"))
    np.savez_compressed(os.path.join(HERE, "lba_small_fixed2.npz"), **window_case("small", n_fixed_extra=2))
    b = synth.make_imu_batch(n_pairs=24, seed=123, ragged=True, leading_partial=True)
    out = pyoracle.preintegrate_batch(b.sample_begin, b.gyro, b.acc, b.dt, b.bg, b.ba)
    np.savez_compressed(os.path.join(HERE, "preint_ragged24.npz"), sample_begin=b.sample_begin, gyro=b.gyro, acc=b.acc,
                        dt=b.dt, bg=b.bg, ba=b.ba, out=out)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
