#!/usr/bin/env python
"""Golden vectors EXECUTED BY THE REFERENCE: runs the reference's own IMUPreintegrator / SO3 sources (compiled
unmodified by `make -C oracle ref`, see oracle/ref_harness.cpp) on seeded inputs and stores inputs + outputs in
tests/golden/ref_imu_v1.npz.  The reference tree only exists in the build container, so the vectors are committed; the
GPU box checks the CUDA pre-integration kernel and the oracle against them without the reference.
Run from the repository root:  python tests/golden/make_ref_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mc_slam_b200 import synth  # noqa: E402
from oracle import pyref  # noqa: E402

rng = np.random.default_rng(20261018)
# BASELINE config 2 shape (40 samples at 200 Hz per pair), 64 pairs, and a ragged batch with 1..600-sample pairs
# (S = 10..100 typical, ~600 worst case) incl. the leading partial interval of KeyFrame::ComputePreInt
b = synth.make_imu_batch(n_pairs=64, n_samples=40, seed=7)
lens = np.concatenate([[1, 2, 3, 600], rng.integers(10, 101, size=20)])
sb = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
ns = int(sb[-1])
t = np.arange(ns) * 0.005
gyro = np.stack([0.3 * np.sin(0.7 * t) + 0.01, 0.2 * np.cos(0.5 * t), 0.4 * np.sin(0.3 * t + 1)], 1) + rng.normal(0, 2.4e-3, (ns, 3))
acc = np.stack([0.5 * np.sin(t), 0.3 * np.cos(0.8 * t), 9.81 + 0.2 * np.sin(0.4 * t)], 1) + rng.normal(0, 2.8e-2, (ns, 3))
dt = np.full(ns, 0.005)
dt[sb[:-1]] = rng.uniform(5e-4, 5e-3, size=sb.size - 1)  # leading partial interval of every pair
bg = rng.normal(0, 0.01, (lens.size, 3))
ba = rng.normal(0, 0.05, (lens.size, 3))
# SO3 probes: generic, tiny (both sides of the 1e-10 and 1e-5 thresholds), near pi
ws = np.concatenate([rng.normal(0, 1.0, (24, 3)), rng.normal(0, 1, (4, 3)) * 1e-11, rng.normal(0, 1, (4, 3)) * 3e-6,
                     rng.normal(0, 1, (4, 3)) * 3e-5, np.array([[3.1, 0.2, -0.1], [0, 0, 3.14159], [2.2, -2.2, 0.1]])])
out = dict(
    c2_sample_begin=b.sample_begin, c2_gyro=b.gyro, c2_acc=b.acc, c2_dt=b.dt, c2_bg=b.bg, c2_ba=b.ba,
    c2_out=pyref.preintegrate_batch(b.sample_begin, b.gyro, b.acc, b.dt, b.bg, b.ba),
    rg_sample_begin=sb, rg_gyro=gyro, rg_acc=acc, rg_dt=dt, rg_bg=bg, rg_ba=ba,
    rg_out=pyref.preintegrate_batch(sb, gyro, acc, dt, bg, ba),
    so3_w=ws, so3_exp=np.stack([pyref.so3_exp(w) for w in ws]),
    so3_log=np.stack([pyref.so3_log(pyref.so3_exp(w)) for w in ws]),
    so3_matrix=np.stack([pyref.so3_matrix(pyref.so3_exp(w)) for w in ws]),
    jr=np.stack([pyref.jacobian_r(w) for w in ws]), jr_inv=np.stack([pyref.jacobian_r_inv(w) for w in ws]),
    imu_constants=pyref.imu_constants(),
)
path = os.path.join(ROOT, "tests", "golden", "ref_imu_v1.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes;", pyref.build_info())
