#!/usr/bin/env python
"""BASELINE config 2: batched IMUPreintegrator::update, 4096 key-frame pairs x 40 samples (covariance + bias
Jacobians).  Device-resident kernel time (CUDA events via torch on the library's work: the call is
stream-ordered, so a synchronize brackets it), end-to-end through vilba_preintegrate_batch with host buffers,
and the CPU restatement on one host thread."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mc_slam_b200 import api, synth  # noqa: E402
from oracle import pyoracle  # noqa: E402

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
S = int(sys.argv[2]) if len(sys.argv) > 2 else 40
b = synth.make_imu_batch(n_pairs=n_pairs, n_samples=S)
ctx = api.Context(0)
dev = torch.device("cuda", 0)
sb = torch.from_numpy(np.ascontiguousarray(b.sample_begin, np.int32)).to(dev)
g = torch.from_numpy(np.ascontiguousarray(b.gyro, np.float64).reshape(-1)).to(dev)
a = torch.from_numpy(np.ascontiguousarray(b.acc, np.float64).reshape(-1)).to(dev)
t = torch.from_numpy(np.ascontiguousarray(b.dt, np.float64).reshape(-1)).to(dev)
bg = torch.from_numpy(np.ascontiguousarray(b.bg, np.float64).reshape(-1)).to(dev)
ba = torch.from_numpy(np.ascontiguousarray(b.ba, np.float64).reshape(-1)).to(dev)
out = torch.empty(n_pairs * 142, dtype=torch.float64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run():
    ctx.preintegrate_batch_dev(n_pairs, n_pairs * S, sb.data_ptr(), g.data_ptr(), a.data_ptr(), t.data_ptr(), bg.data_ptr(),
                               ba.data_ptr(), out.data_ptr())


for _ in range(3):
    run()
torch.cuda.synchronize()
times = []
for _ in range(20):
    flush.fill_(1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run()
    ctx.sync() if hasattr(ctx, "sync") else torch.cuda.synchronize()
    times.append(time.perf_counter() - t0)
dev_s = float(np.median(times))
ref = pyoracle.preintegrate_batch(b.sample_begin, b.gyro, b.acc, b.dt, b.bg, b.ba)
got = out.cpu().numpy().reshape(n_pairs, 142)
ok = bool(np.allclose(got[:, :60], ref[:, :60], rtol=1e-10, atol=1e-12))
e2e = []
for _ in range(5):
    t0 = time.perf_counter()
    ctx.preintegrate_batch(b.sample_begin, b.gyro, b.acc, b.dt, b.bg, b.ba)
    e2e.append(time.perf_counter() - t0)
t0 = time.perf_counter()
pyoracle.preintegrate_batch(b.sample_begin, b.gyro, b.acc, b.dt, b.bg, b.ba)
cpu_s = time.perf_counter() - t0
bytes_alg = n_pairs * S * 56 + n_pairs * (48 + 1136)
print("PREINT " + json.dumps({
    "pairs": n_pairs, "samples_per_pair": S, "device_ms_host_timed": 1e3 * dev_s, "pairs_per_sec_device": n_pairs / dev_s,
    "updates_per_sec_device": n_pairs * S / dev_s, "algorithmic_bytes": bytes_alg, "achieved_gbs": bytes_alg / dev_s / 1e9,
    "e2e_ms": 1e3 * float(np.median(e2e)), "pairs_per_sec_e2e": n_pairs / float(np.median(e2e)),
    "cpu_1thread_ms": 1e3 * cpu_s, "pairs_per_sec_cpu": n_pairs / cpu_s, "parity_ok": ok}))
