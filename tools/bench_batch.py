#!/usr/bin/env python
"""Throughput of independent windows solved as ONE batched launch per kernel (BASELINE config 5 shape:
64 C3 windows per GPU).  Prints the device time of the resident batch and the end-to-end time through
vilba_local_ba_batch (host buffers in, host buffers out)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mc_slam_b200 import api, synth  # noqa: E402

name = sys.argv[2] if len(sys.argv) > 2 else "c3"
sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "32").split(",")]
base = [synth.make_config(name, window_index=i) for i in range(min(max(sizes), 8))]
ctx = api.Context(0)
prof = os.environ.get("BATCH_PROFILE", "0") == "1"
for n in sizes:
    wins = [base[i % len(base)] for i in range(n)]
    if n <= min(api.max_batch(), int(os.environ.get("VILBA_MAX_BATCH", "64"))):
        ctx.upload_batch(wins)
        ctx.solve_batch_resident()
        rs = ctx.solve_batch_resident()
        iters = sum(len(r.trace) for r in rs)
        ms = rs[0].solve_ms
        print(f"resident batch n={n} ({ctx.batch_groups()} lanes) {name}: {ms:.2f} ms device -> {iters/ms*1e3:.0f} LM iters/s, {n/ms*1e3:.1f} windows/s", flush=True)
        if prof:
            ctx.reset_stats()
            ctx.set_profiling(True)
            ctx.solve_batch_resident()
            s = ctx.stats()
            ctx.set_profiling(False)
            print(f"   per launch: linearize {1e3*s.linearize_ms/max(1,s.linearize_launches):.1f} us x{s.linearize_launches}, "
                  f"schur {1e3*s.schur_ms/max(1,s.schur_launches):.1f} us, chol {1e3*s.solve_ms/max(1,s.solve_launches):.1f} us x{s.solve_launches}", flush=True)
    prep = ctx.prepare(wins)  # caller-owned host buffers, reused
    prep.run()  # warm-up: lanes, graphs, arenas
    prep.run()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        prep.run()
    dt = (time.perf_counter() - t0) / reps
    iters = sum(len(r.trace) for r in prep.collect())
    print(f"e2e batch n={n} {name}: {dt*1e3:.1f} ms -> {n/dt:.1f} windows/s, {iters/dt:.0f} LM iters/s (host buffers)", flush=True)
