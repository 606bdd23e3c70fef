#!/usr/bin/env python
"""bench.py -- headline benchmark of the VI local-BA hot path (BASELINE.json metric:
"local-BA LM iters/sec + edges linearized/sec (20KF/5k pts), % HBM roofline").

One "step" = one full pass of the path over one batch of synthetic EuRoC-shaped 20-KF / 5k-point windows
(BASELINE config 5's per-GPU share: 512 windows over 8 GPUs = 64 independent windows per GPU, solved by ONE
batched launch per kernel): phases C..E of Optimizer::LocalBundleAdjustmentNavState for every window
(5 robust + 10 non-robust LM iterations, cull, outlier flags).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one process per GPU)
    python bench.py --impl reference ...                     # the CPU restatement of the reference path
    python bench.py --windows 1                              # latency of ONE window (BASELINE config 3)

value   : LM outer iterations / s over all windows, windows resident in HBM when the timed region starts
          (device time from CUDA events on the library's stream, L2 flushed between steps, max over ranks)
e2e     : same metric through the C-ABI call vilba_local_ba_batch() with HOST buffers
          (flatten + H2D + solve + D2H inside the timed region)
roofline / roofline_all : the dominant kernel group of the TIMED configuration / all four groups (per-group device
          times from a pass with CUDA events between the kernels, lanes concurrent as in the timed pass)
single_window : the same two numbers for ONE 20-KF window (latency-bound; BASELINE config 3) and their ratio to ONE
          host thread of the CPU restatement (the north star's >= 50x target is written on this pair)
c1, c2_preint, c4 : BASELINE configs 1, 2 and 4 on one GPU (N = 1 only)
sharded_c4 : config 4 point-sharded over the N ranks (vilba_comm_init: NCCL allreduce of the partial normal equations
          per LM trial, reduced system solved redundantly) -- the strong-scaling curve of the driver's 1/2/4/8 run
Multi-GPU: windows are independent -> each rank owns its own batch, no data-path collective, "weak"
          scaling; torch.distributed only carries the barrier/max.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "local_ba_lm_iters_per_sec"
UNIT = "LM iters/s"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_desc(w, name, nw):
    one = (f"{name}: {w.n_kf} KF ({w.n_free} free) / {w.n_pts} pts / {w.n_obs} mono + {w.n_imu}+{w.n_imu} IMU edges, "
           f"5+10 LM iters")
    return one if nw == 1 else f"{nw} independent windows per GPU in one batched launch (C5 share), each ~ {one}"


def algorithmic_bytes_linearize(w):
    """B_lin = E*20 + P*24 + E'*144 + P*(72+24) + n^2*8/2 + n*8 (FP64 mode)."""
    from mc_slam_b200 import capi
    free = (w.kf_flags & capi.KF_FIXED) == 0
    e_free = int(np.count_nonzero(free[w.obs_kf]))
    n = 15 * w.n_free
    return w.n_obs * 20 + w.n_pts * 24 + e_free * 144 + w.n_pts * 96 + n * n * 4 + n * 8


def _edges_to_free(w):
    from mc_slam_b200 import capi
    free = (w.kf_flags & capi.KF_FIXED) == 0
    return int(np.count_nonzero(free[w.obs_kf]))


def algorithmic_bytes_schur(w):
    """B_schur = E'*144 + P*96 read + n^2*4 + n*8 written."""
    n = 15 * w.n_free
    return _edges_to_free(w) * 144 + w.n_pts * 96 + n * n * 4 + n * 8


def algorithmic_bytes_update(w):
    """B_bs + B_ev = (E'*144 + P*96 + n*8 + P*24) + (E*20 + P*48 + E*8)."""
    n = 15 * w.n_free
    return (_edges_to_free(w) * 144 + w.n_pts * 96 + n * 8 + w.n_pts * 24) + (w.n_obs * 20 + w.n_pts * 48 + w.n_obs * 8)


def solve_flops(w):
    n = 15 * w.n_free
    return n ** 3 / 3.0 + 2.0 * n ** 2


FP64_PEAK_TFLOPS = 37.2  # DFMA throughput measured on this part with tools/ubench_fp64.cu (MEASURED_PEAKS.json has no FP64 figure)


def ncu_traffic(kernel, nw, workload):
    """dram read+write bytes per launch of `kernel` from the committed ncu --set full capture
    (profiles/ncu_traffic.json: {"<workload>x<windows>": {"<kernel>": bytes}}), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f)[f"{workload}x{nw}"][kernel]
    except Exception:
        return None


def cpu_throughput(wins, threads, reps_per_thread, warm=True):
    """The CPU restatement of the reference path (oracle/, g++ -O3 -march=native) on `threads` host threads,
    one window per thread at a time (g2o itself is single-threaded per optimisation: Thirdparty/g2o/config.h:4).
    Returns (LM iters/s over the wall time of the sample, windows solved, iters, edges, wall seconds)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle
    if warm:
        pyoracle.local_ba(wins[0])  # warm-up (library load, page-in)
    jobs = [wins[i % len(wins)] for i in range(threads * reps_per_thread)]

    def one(w):
        r = pyoracle.local_ba(w)  # ctypes releases the GIL inside the C call
        return len(r.trace), sum(t["n_active_edges"] for t in r.trace), r.solve_ms

    t0 = time.perf_counter()
    if threads == 1:
        out = [one(w) for w in jobs]
    else:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            out = list(ex.map(one, jobs))
    wall = time.perf_counter() - t0
    iters = sum(o[0] for o in out)
    edges = sum(o[1] for o in out)
    solve_s = sum(o[2] for o in out) * 1e-3
    # single-thread figure: phases C..E only (CUDA-event equivalent); multi-thread: wall time of the sample
    val = iters / solve_s if threads == 1 else iters / wall
    return val, len(jobs), iters, edges, wall


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def config_of(w, name, nw):
    """`config` of the JSON line: the SAME dict in both arms (the driver compares it)."""
    return {"workload": workload_desc(w, name, nw)}


def run_reference(args, rank, world):
    """CPU arm: the dependency-free restatement of the reference's g2o path (the reference itself cannot be
    compiled here as a whole: Eigen/OpenCV/CHOLMOD are absent; its pre-integration, factors, LM step and Huber kernel are
    pinned against the compiled reference sources, tests/test_oracle_*vs_ref.py), on all host threads the workload can use: one window per thread (each
    optimisation is single-threaded like the reference's).  One step = one window solve on every thread: a bounded
    sample of the workload (a 20-KF window is ~0.3 s on one core)."""
    if rank != 0:
        return
    from mc_slam_b200 import synth
    nw = args.windows
    threads = 1 if nw == 1 else min(host_threads(), nw)
    wins = [synth.make_config(args.workload, window_index=i) for i in range(min(nw, max(threads, 1)))]
    W, K = max(0, args.warmup), max(1, args.steps)
    if W:
        cpu_throughput(wins, threads, min(W, 2))
    val, n_solved, iters, edges, wall = cpu_throughput(wins, threads, K)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 * wall / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_of(wins[0], args.workload, nw),
        "edges_linearized_per_sec": edges / wall if threads > 1 else None,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n_solved} window solves ({K} steps of one window per thread) of the same workload, "
                                   f"oracle/libvilba_oracle.so, {threads} threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def window_bytes(w):
    h2d = (w.kf_state.nbytes + w.pt_xyz.nbytes + w.imu_preint.nbytes + 16 * w.n_obs + w.pt_obs_begin.nbytes
           + 4 * w.n_kf + 8 * w.n_imu)
    d2h = w.kf_state.nbytes + w.pt_xyz.nbytes + w.n_obs  # states, points, outlier flags: what the reference's function returns
    return h2d, d2h


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c1", "c3", "c4", "small", "tiny"])
    ap.add_argument("--windows", type=int, default=64, help="independent windows per GPU solved as one batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-single", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the config 1 / 2 / 4 sections")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from mc_slam_b200 import api, sharding, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the VI local-BA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W = max(3, args.warmup)
    K = max(1, args.steps)
    nw = max(1, min(args.windows, api.max_batch()))
    # every rank owns its own slice of the global list of independent windows (C5: window w uses seed + 1000 w)
    w0, _ = sharding.shard_range(nw * world, rank, world)
    wins = [synth.make_config(args.workload, window_index=w0 + i) for i in range(nw)]
    ctx = api.Context(local_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def flush_l2():
        flush.fill_(1)
        torch.cuda.synchronize()

    def timed_resident(c, k):
        ms, iters, edges = 0.0, 0, 0
        for _ in range(k):
            flush_l2()
            rs = c.solve_batch_resident()
            ms += rs[0].solve_ms  # device time of the whole batch
            iters += sum(len(r.trace) for r in rs)
            edges += sum(t["n_active_edges"] for r in rs for t in r.trace)
        return ms, iters, edges

    # ---- HBM-resident timing --------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()  # nvidia-smi needs ~1 s to start: sample from the warm-up on (same load), through the timed region
    ctx.upload_batch(wins)
    groups = max(1, ctx.batch_groups())  # concurrent lanes: one launch covers nw / groups windows
    t_warm = time.perf_counter()
    while True:  # at least W warm-up steps and ~1.5 s of load so that the clock samples cover the timed region
        timed_resident(ctx, W)
        if time.perf_counter() - t_warm > 1.5:
            break
    ctx.reset_stats()
    barrier()
    wall0 = time.perf_counter()
    dev_ms, iters, edges = timed_resident(ctx, K)
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    st = ctx.stats()
    # ---- per-kernel-group device times: the same steps again with CUDA events around each group (this
    #      pass launches the kernels one by one instead of through the CUDA graph; it is not the timed one)
    ctx.reset_stats()
    ctx.set_profiling(True)
    timed_resident(ctx, min(K, 5))
    stp = ctx.stats()
    ctx.set_profiling(False)

    # ---- end-to-end through the C ABI with host buffers --------------------------------------------
    # (the caller owns its input and output buffers, like the C++ shim does: they are allocated once, outside
    #  the timed region; the timed call flattens, copies H2D, solves, copies D2H and scatters into them)
    prep = ctx.prepare(wins, chi2=False)
    for _ in range(2):
        prep.run()
    barrier()
    e2e_s, e2e_iters = 0.0, 0
    for _ in range(K):
        flush_l2()
        t0 = time.perf_counter()
        prep.run()
        e2e_s += time.perf_counter() - t0
        e2e_iters += sum(len(r.trace) for r in prep.collect())
    barrier()
    h2d = sum(window_bytes(w)[0] for w in wins)
    d2h = sum(window_bytes(w)[1] for w in wins)

    # ---- one window alone (BASELINE config 3: latency) ------------------------------------------------
    single = None
    if nw > 1 and not args.no_single:
        c1 = api.Context(local_rank)
        c1.upload_batch(wins[:1])
        timed_resident(c1, W)
        s_ms, s_iters, _ = timed_resident(c1, K)
        prep1 = c1.prepare(wins[:1], chi2=False)
        for _ in range(2):
            prep1.run()
        s_e2e, s_e2e_iters = 0.0, 0
        for _ in range(K):
            flush_l2()
            t0 = time.perf_counter()
            prep1.run()
            s_e2e += time.perf_counter() - t0
            s_e2e_iters += len(prep1.collect()[0].trace)
        single = {"value": s_iters / (s_ms * 1e-3), "unit": UNIT, "ms_per_window": s_ms / K,
                  "e2e": s_e2e_iters / s_e2e, "e2e_ms_per_window": 1e3 * s_e2e / K,
                  "note": "ONE 20-KF window per call (BASELINE config 3): L2-resident and latency-bound"}
        c1.close()

    # ---- BASELINE configs 1, 2, 4 on one GPU, and config 4 point-sharded over the ranks ---------------------
    def one_window(name, k):
        """resident + end-to-end LM iterations/s of ONE window of config `name` (non-sharded path)."""
        w = synth.make_config(name)
        c = api.Context(local_rank)
        c.upload_batch([w])
        timed_resident(c, 2)
        ms, it, ed = timed_resident(c, k)
        pr = c.prepare([w], chi2=False)
        pr.run()
        t_e2e, it_e2e = 0.0, 0
        for _ in range(k):
            flush_l2()
            t0 = time.perf_counter()
            pr.run()
            t_e2e += time.perf_counter() - t0
            it_e2e += len(pr.collect()[0].trace)
        c.reset_stats()
        c.set_profiling(True)
        timed_resident(c, 1)
        sp = c.stats()
        c.set_profiling(False)
        c.close()
        per = lambda ms_, n_: 1e3 * ms_ / max(1, n_)  # noqa: E731
        return w, {"workload": workload_desc(w, name, 1), "value": it / (ms * 1e-3), "unit": UNIT, "ms_per_solve": ms / k,
                   "ms_per_iter": ms / max(1, it), "edges_linearized_per_sec": ed / (ms * 1e-3),
                   "e2e": it_e2e / t_e2e, "e2e_ms_per_solve": 1e3 * t_e2e / k,
                   "kernels_us": {"linearize": per(sp.linearize_ms, sp.linearize_launches), "schur": per(sp.schur_ms, sp.schur_launches),
                                  "chol_solve": per(sp.solve_ms, sp.solve_launches), "update_eval": per(sp.update_ms, sp.update_launches)}}

    extra = {}
    if world == 1 and not args.no_extra and args.workload == "c3" and nw > 1:
        _, extra["c1"] = one_window("c1", max(3, min(K, 10)))
        _, extra["c4"] = one_window("c4", 3)
        # config 2: 4096 key-frame pairs x 40 samples, inputs resident (device pointers) and end to end (host buffers)
        b = synth.make_imu_batch(n_pairs=4096, n_samples=40)
        dev = torch.device("cuda", local_rank)
        t_ = lambda a_, dt_=np.float64: torch.from_numpy(np.ascontiguousarray(a_, dt_).reshape(-1)).to(dev)  # noqa: E731
        sb, g_, a_, d_, bg_, ba_ = t_(b.sample_begin, np.int32), t_(b.gyro), t_(b.acc), t_(b.dt), t_(b.bg), t_(b.ba)
        o_ = torch.empty(4096 * 142, dtype=torch.float64, device=dev)
        cp = api.Context(local_rank)
        run_p = lambda: cp.preintegrate_batch_dev(4096, 4096 * 40, sb.data_ptr(), g_.data_ptr(), a_.data_ptr(), d_.data_ptr(),  # noqa: E731
                                                  bg_.data_ptr(), ba_.data_ptr(), o_.data_ptr())
        for _ in range(3):
            run_p()
        cp.reset_stats()
        cp.set_profiling(True)
        for _ in range(20):
            flush_l2()
            run_p()
        sp = cp.stats()
        cp.set_profiling(False)
        p_us = 1e3 * sp.preint_ms / max(1, sp.preint_launches)
        cp.preintegrate_batch(b.sample_begin, b.gyro, b.acc, b.dt, b.bg, b.ba)
        t0 = time.perf_counter()
        for _ in range(5):
            cp.preintegrate_batch(b.sample_begin, b.gyro, b.acc, b.dt, b.bg, b.ba)
        p_e2e = (time.perf_counter() - t0) / 5
        cp.close()
        p_bytes = 4096 * 40 * 56 + 4096 * (48 + 1136)  # SURVEY 8(d) B_pre
        p_flops = 4096 * 40 * 600.0
        peak_, _ = _peaks()
        extra["c2_preint"] = {
            "workload": "c2: 4096 key-frame pairs x 40 IMU samples (200 Hz), covariance + bias Jacobians",
            "pairs_per_sec": 4096 / (p_us * 1e-6), "updates_per_sec": 4096 * 40 / (p_us * 1e-6), "ms": p_us * 1e-3,
            "e2e_pairs_per_sec": 4096 / p_e2e, "e2e_ms": 1e3 * p_e2e,
            "roofline": {"kernel": "preint_batch_kernel", "bound": "fp64 latency", "algorithmic_bytes": p_bytes,
                         "achieved_gbs": p_bytes / (p_us * 1e-6) / 1e9, "frac_hbm": p_bytes / (p_us * 1e-6) / 1e9 / peak_,
                         "achieved_tflops": p_flops / (p_us * 1e-6) / 1e12, "frac_fp64": p_flops / (p_us * 1e-6) / 1e12 / FP64_PEAK_TFLOPS},
            "timing": "CUDA events on the library stream around the kernel, 20 launches, L2 flushed in between"}

    # config 4 sharded by map point over the ranks of this run (at N = 1: the collective path with one rank)
    sharded = None
    if not args.no_extra and args.workload == "c3" and nw > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(api.comm_unique_id()), dtype=torch.uint8))
        if world > 1:
            dist.broadcast(uid, 0)
        cs = api.Context(local_rank)
        cs.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
        w4 = synth.make_config("c4")
        sub, p0, p1, e0, e1 = sharding.shard_window(w4, rank, world)
        s_ms, s_it = 0.0, 0
        for it_ in range(1 + 3):
            barrier()
            r4 = cs.local_ba(sub)
            if it_ >= 1:
                s_ms += r4.solve_ms
                s_it += len(r4.trace)
        cs.reset_stats()
        cs.set_profiling(True)
        barrier()
        r4 = cs.local_ba(sub)
        sp = cs.stats()
        cs.set_profiling(False)
        (s_ms_max,), _ = sharding.reduce_bench([s_ms], [0.0], device="cuda")
        n4 = 15 * w4.n_free
        lds4 = (n4 + 3) & ~3
        sharded = {
            "workload": workload_desc(w4, "c4", 1) + f", map points sharded over {world} rank(s)",
            "lm_iters_per_sec": s_it / (s_ms_max * 1e-3), "ms_per_iter": s_ms_max / max(1, s_it), "ms_per_solve": s_ms_max / 3,
            "edges_per_rank": int(e1 - e0), "points_per_rank": int(p1 - p0),
            "nccl_calls_per_slot": 3, "nccl_calls_per_trial": 2,
            "allreduce_bytes_per_slot": int(8 * (n4 + world) + 8 * (lds4 * n4 + n4) + 16),
            "chol_ms": sp.solve_ms / max(1, sp.solve_launches), "schur_ms": sp.schur_ms / max(1, sp.schur_launches),
            "linearize_ms": sp.linearize_ms / max(1, sp.linearize_launches),
            "note": "one slot = linearise-if-needed + one LM trial; per trial allreduce(sum) of S|b_s and of chi2|scale, per "
                    "linearisation allreduce(sum) of [diag H_pp | per-rank max diag H_ll] (H_pp and b_p stay rank-local: every "
                    "rank folds its partial sums into its partial S); the reduced system is solved redundantly on every rank; "
                    "device time (CUDA events), max over ranks, 3 solves after 1 warm-up; per-kernel times from a profiled solve"}
        cs.close()

    # ---- reduce over ranks ----------------------------------------------------------------------------
    (dev_ms_max, e2e_s_max), (iters_all, e2e_iters_all, edges_all, launches_all) = sharding.reduce_bench(
        [dev_ms, e2e_s], [float(iters), float(e2e_iters), float(edges), float(st.kernel_launches)], device="cuda")

    if rank == 0:
        peak, peak_src = _peaks()
        n_red = 15 * wins[0].n_free
        per = lambda ms_, n_: 1e3 * ms_ / max(1, n_)  # noqa: E731
        lin_us, schur_us = per(stp.linearize_ms, stp.linearize_launches), per(stp.schur_ms, stp.schur_launches)
        chol_us, upd_us = per(stp.solve_ms, stp.solve_launches), per(stp.update_ms, stp.update_launches)
        b_lin = sum(algorithmic_bytes_linearize(w) for w in wins) / groups
        b_schur = sum(algorithmic_bytes_schur(w) for w in wins) / groups
        b_upd = sum(algorithmic_bytes_update(w) for w in wins) / groups
        f_chol = sum(solve_flops(w) for w in wins) / groups
        gbs = lambda b_, us_: b_ / (us_ * 1e-6) / 1e9 if us_ > 0 else 0.0  # noqa: E731
        lane_w = nw // groups
        roofline_all = [
            {"kernel": "linearize_v2 + reduce_partials + assemble_hpp (linearize_imu_v2 beside it)", "bound": "hbm",
             "bytes": int(b_lin), "us": lin_us, "achieved": gbs(b_lin, lin_us), "unit": "GB/s", "frac": gbs(b_lin, lin_us) / peak,
             "traffic": ncu_traffic("linearize_v2_kernel", lane_w, args.workload)},
            {"kernel": "schur_rec + schur_tile + schur_finish", "bound": "hbm", "bytes": int(b_schur), "us": schur_us,
             "achieved": gbs(b_schur, schur_us), "unit": "GB/s", "frac": gbs(b_schur, schur_us) / peak,
             "traffic": ncu_traffic("schur_tile_kernel", lane_w, args.workload)},
            {"kernel": "chol_la (reduced-system LDL^T + solve)", "bound": "fp64", "flops": f_chol, "us": chol_us,
             "achieved": f_chol / (max(1e-9, chol_us) * 1e-6) / 1e12, "unit": "TFLOP/s",
             "frac": f_chol / (max(1e-9, chol_us) * 1e-6) / 1e12 / FP64_PEAK_TFLOPS, "traffic": None},
            {"kernel": "update_eval (oplus + landmark back-substitution + residuals)", "bound": "hbm", "bytes": int(b_upd),
             "us": upd_us, "achieved": gbs(b_upd, upd_us), "unit": "GB/s", "frac": gbs(b_upd, upd_us) / peak,
             "traffic": ncu_traffic("update_eval_kernel", lane_w, args.workload)},
        ]
        dom = max(roofline_all, key=lambda r_: r_["us"])
        ws_mb = nw * 21 if args.workload == "c3" else None
        line = {
            "metric": METRIC, "value": iters_all / (dev_ms_max * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(wins[0], args.workload, nw),
            "run": {"l2": "flushed (256 MiB write) between steps" + (f"; working set ~{ws_mb} MB per GPU" if ws_mb else ""),
                    "windows_per_gpu": nw, "lanes": groups,
                    "parallelism": f"independent windows, {nw} per GPU x {world} GPUs, no collective; per GPU {groups} "
                                   f"concurrent lanes of ~{nw // groups} windows, one batched launch per kernel and lane",
                    "timing": "CUDA events on the library stream around each batched solve, summed over steps, max over ranks"},
            "edges_linearized_per_sec": edges_all / (dev_ms_max * 1e-3),
            "windows_per_sec": nw * world * K / (dev_ms_max * 1e-3),
            "lm_iters_per_step": iters / K,
            "wall_s_timed_region": wall,
            "e2e": {"value": e2e_iters_all / e2e_s_max, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s_max / K,
                    "call": "vilba_local_ba_batch (host buffers in and out)" if nw > 1 else "vilba_local_ba"},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            # the dominant kernel group of the timed configuration (largest device time per slot of one lane)
            "roofline": {"kernel": dom["kernel"], "bound": "hbm" if dom["bound"] == "hbm" else "tensor",
                         "achieved": dom["achieved"], "peak": peak if dom["bound"] == "hbm" else FP64_PEAK_TFLOPS,
                         "unit": dom["unit"], "frac": dom["frac"], "traffic": dom["traffic"], "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": dom.get("bytes"), "avg_launch_us": dom["us"],
                         "note": "algorithmic bytes = SURVEY 8(d) formula summed over the windows of one lane (one launch covers a "
                                 "lane); launch time = CUDA events between the kernel groups in a pass that runs the lanes "
                                 "concurrently like the timed pass (graphs off); traffic = dram read+write per launch from the "
                                 "committed ncu --set full capture (profiles/ncu_traffic.json)"
                                 + ("" if dom["bound"] == "hbm" else "; FP64 CUDA-core kernel: peak = measured DFMA throughput, "
                                    "reported under the contract's 'tensor' label (no tensor cores are used)")},
            "roofline_all": roofline_all,
            "linearize_accumulate_frac_hbm": roofline_all[0]["frac"],
            "kernels_us": {"linearize": lin_us, "schur": schur_us, "chol_solve": chol_us, "update_eval": upd_us,
                           "note": "per batched launch of one lane; CUDA events around each kernel group in a second, ungraphed pass "
                                   "with the lanes concurrent (the timed configuration)"},
        }
        line.update(extra)
        if sharded:
            line["sharded_c4"] = sharded
        if single:
            line["single_window"] = single
        if not args.no_cpu_baseline and world == 1:  # the CPU baseline is timed at N = 1 only
            threads = 1 if nw == 1 else min(host_threads(), nw)
            reps = 3
            val, n_solved, _, _, _ = cpu_throughput(wins[:max(threads, 1)], threads, reps)
            line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{n_solved} window solves of the same workload on {threads} host threads, one "
                                              "window per thread (oracle restatement; faster than real g2o's MatrixXd path, "
                                              "so the ratio is conservative)"}
            v1 = val
            if threads > 1:
                v1, _, _, _, _ = cpu_throughput(wins[:1], 1, 2)
                line["cpu_baseline"]["single_thread_value"] = v1
            if single:
                single["cpu_single_thread"] = v1
                single["ratio_vs_cpu_single_thread"] = single["value"] / v1
                single["e2e_ratio_vs_cpu_single_thread"] = single["e2e"] / v1
            # configs 1, 2 and 4 next to ONE host core running the restatement (C4: one solve of ~10 s)
            for name, reps_ in (("c1", 3), ("c4", 1)):
                if name in line:
                    vc, _, _, _, _ = cpu_throughput([synth.make_config(name)], 1, reps_, warm=name != "c4")
                    line[name]["cpu_single_thread"] = vc
                    line[name]["ratio_vs_cpu_single_thread"] = line[name]["value"] / vc
            if "c2_preint" in line:
                from oracle import pyoracle  # (the CPU baseline leg is where bench.py may execute oracle/)
                b2 = synth.make_imu_batch(n_pairs=512, n_samples=40)
                pyoracle.preintegrate_batch(b2.sample_begin, b2.gyro, b2.acc, b2.dt, b2.bg, b2.ba)
                t0 = time.perf_counter()
                pyoracle.preintegrate_batch(b2.sample_begin, b2.gyro, b2.acc, b2.dt, b2.bg, b2.ba)
                line["c2_preint"]["cpu_single_thread_pairs_per_sec"] = 512 / (time.perf_counter() - t0)
                line["c2_preint"]["cpu_sample"] = "512 of the 4096 pairs, one host thread, oracle restatement"
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
