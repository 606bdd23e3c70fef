#!/usr/bin/env python
"""Throughput of independent windows through vilba_local_ba_batch (BASELINE config 5 shape, scaled down):
N copies of distinct C3-shaped windows, host buffers in, host buffers out."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mc_slam_b200 import api, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
name = sys.argv[2] if len(sys.argv) > 2 else "c3"
base = [synth.make_config(name, window_index=i) for i in range(min(n, 8))]
wins = [base[i % len(base)] for i in range(n)]
ctx = api.Context(0)
ctx.local_ba_batch(wins[: min(n, 8)])  # warm-up: lanes, graphs, arenas
for lanes_note in (os.environ.get("VILBA_BATCH_LANES", "8"),):
    t0 = time.perf_counter()
    rs = ctx.local_ba_batch(wins)
    dt = time.perf_counter() - t0
    iters = sum(len(r.trace) for r in rs)
    print(f"lanes={lanes_note} windows={n} {name}: {dt*1e3:.1f} ms -> {n/dt:.1f} windows/s, {iters/dt:.0f} LM iters/s (e2e, host buffers)")
