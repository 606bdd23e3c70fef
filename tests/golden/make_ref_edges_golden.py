#!/usr/bin/env python
"""Golden vectors EXECUTED BY THE REFERENCE for the three factors of the path: runs the reference's own
EdgeNavStatePVR / EdgeNavStateBias / EdgeNavStatePVRPointXYZ (src/IMU/g2otypes.cpp, compiled unmodified by
`make -C oracle ref` against oracle/eigen_stub + oracle/g2o_stub, see oracle/ref_harness_edges.cpp) and the vertices'
oplusImpl on seeded inputs and stores inputs + outputs in tests/golden/ref_edges_v1.npz.  The reference tree only exists
in the build container, so the vectors are committed; everywhere else the oracle is checked against them.
Run from the repository root:  python tests/golden/make_ref_edges_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mc_slam_b200 import synth  # noqa: E402
from oracle import pyref  # noqa: E402

rng = np.random.default_rng(20261019)
N = 48


def nav_state(pos_scale=2.0):
    return np.concatenate([rng.normal(0, pos_scale, 3), rng.normal(0, 1.0, 3), pyref.so3_exp(rng.normal(0, 1.0, 3)),
                           rng.normal(0, 0.01, 3), rng.normal(0, 0.05, 3), rng.normal(0, 1e-3, 3), rng.normal(0, 1e-2, 3)])


# ---- IMU edges: a key-frame pair 0.2 s apart (40 samples at 200 Hz), states near what the samples imply plus a perturbation
b = synth.make_imu_batch(n_pairs=N, n_samples=40, seed=11)
g = np.array([0.0, 0.0, -9.81])
pvr = dict(ns_i=[], ns_j=[], err=[], Ji=[], Jj=[], Jb=[], preint=[])
for p in range(N):
    s0, s1 = int(b.sample_begin[p]), int(b.sample_begin[p + 1])
    ns_i = nav_state()
    ns_i[10:13], ns_i[13:16] = b.bg[p], b.ba[p]
    ns_j = nav_state()
    ns_j[0:3] = ns_i[0:3] + 0.2 * ns_i[3:6] + rng.normal(0, 0.05, 3)
    ns_j[3:6] = ns_i[3:6] + rng.normal(0, 0.3, 3)
    ns_j[6:10] = pyref.so3_mul(ns_i[6:10], pyref.so3_exp(rng.normal(0, 0.15, 3)))
    ns_j[10:16] = ns_i[10:16]
    err, Ji, Jj, Jb = pyref.edge_pvr(b.gyro[s0:s1], b.acc[s0:s1], b.dt[s0:s1], b.bg[p], b.ba[p], ns_i, ns_j, ns_i, g)
    pre = pyref.preintegrate_batch(np.array([0, s1 - s0]), b.gyro[s0:s1], b.acc[s0:s1], b.dt[s0:s1], b.bg[p:p + 1], b.ba[p:p + 1])[0]
    for k, v in zip(("ns_i", "ns_j", "err", "Ji", "Jj", "Jb", "preint"), (ns_i, ns_j, err, Ji, Jj, Jb, pre)):
        pvr[k].append(v)
bias = dict(ns_i=[], ns_j=[], err=[], Ji=[], Jj=[])
for p in range(N):
    a, c = nav_state(), nav_state()
    err, Ji, Jj = pyref.edge_bias(a, c)
    for k, v in zip(("ns_i", "ns_j", "err", "Ji", "Jj"), (a, c, err, Ji, Jj)):
        bias[k].append(v)
# ---- mono edges: points in front of and (a few) behind the camera, EuRoC-like intrinsics and extrinsics
Rbc, Pbc = synth.calib_tbc()
calib = np.concatenate([[synth.FX, synth.FY, synth.CX, synth.CY], np.asarray(Rbc).reshape(-1), np.asarray(Pbc).reshape(-1)])
mono = dict(ns=[], pw=[], uv=[], err=[], Jp=[], Jn=[], depth=[])
for p in range(3 * N):
    ns = nav_state()
    Rwb = pyref.so3_matrix(ns[6:10])
    pc = np.array([rng.uniform(-2, 2), rng.uniform(-1.5, 1.5), rng.uniform(1.0, 9.0) * (-1 if p % 16 == 15 else 1)])
    pw = Rwb @ (np.asarray(Rbc) @ pc + np.asarray(Pbc).reshape(3)) + ns[0:3]
    pw = pw.astype(np.float32).astype(np.float64)  # map points are float-valued at the boundary
    uv = np.array([rng.uniform(0, 752), rng.uniform(0, 480)]).astype(np.float32).astype(np.float64)
    err, Jp, Jn, dpos = pyref.edge_mono(ns, pw, calib, uv)
    for k, v in zip(("ns", "pw", "uv", "err", "Jp", "Jn", "depth"), (ns, pw, uv, err, Jp, Jn, dpos)):
        mono[k].append(v)
# ---- vertex updates
opl = dict(ns=[], d9=[], d6=[], pvr=[], bias=[])
for p in range(N):
    ns, d9, d6 = nav_state(), rng.normal(0, 0.05, 9), rng.normal(0, 1e-3, 6)
    for k, v in zip(("ns", "d9", "d6", "pvr", "bias"), (ns, d9, d6, pyref.vertex_pvr_oplus(ns, d9), pyref.vertex_bias_oplus(ns, d6))):
        opl[k].append(v)
out = {"gravity": g, "calib": calib}
for tag, d in (("pvr", pvr), ("bias", bias), ("mono", mono), ("oplus", opl)):
    for k, v in d.items():
        out[f"{tag}_{k}"] = np.asarray(v)
path = os.path.join(ROOT, "tests", "golden", "ref_edges_v1.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
