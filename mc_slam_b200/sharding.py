"""Host-side sharding of the path over the GPUs of one node (one process per GPU).

Independent windows (BASELINE config 5) and key-frame pairs (config 2) partition with NO data-path
collective: every rank owns a contiguous slice; torch.distributed only carries the barrier and the
max-over-ranks / sum-over-ranks reductions of the bench (works with the gloo backend on CPUs, which is how
the logic is tested without GPUs)."""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import List, Sequence, Tuple

import numpy as np


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of `n_items` for `rank` (the first n_items % world ranks get one more)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    q, r = divmod(max(0, n_items), world)
    begin = rank * q + min(rank, r)
    return begin, begin + q + (1 if rank < r else 0)


def shard_pairs(sample_begin: Sequence[int], rank: int, world: int) -> Tuple[int, int, int, int]:
    """Slice of an IMU batch (CSR by key-frame pair): (pair_begin, pair_end, sample_begin, sample_end)."""
    n_pairs = len(sample_begin) - 1
    p0, p1 = shard_range(n_pairs, rank, world)
    return p0, p1, int(sample_begin[p0]), int(sample_begin[p1])


def reduce_bench(values_max: List[float], values_sum: List[float], device=None):
    """MAX over ranks of the timings, SUM over ranks of the work counters."""
    import torch
    import torch.distributed as dist

    tmax = torch.tensor(values_max, dtype=torch.float64, device=device)
    tsum = torch.tensor(values_sum, dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    return tmax.tolist(), tsum.tolist()


# ---- one large window sharded by map point (BASELINE config 4) ----------------------------------
def shard_points(win, rank: int, world: int) -> Tuple[int, int]:
    """[p_begin, p_end) of the map points `rank` owns: contiguous ranges balanced by edge count
    (vilba_shard_points of the C ABI; host code, no GPU needed)."""
    from . import capi
    lib = capi.load_library()
    cw = win.as_c()
    p0, p1 = C.c_int32(0), C.c_int32(0)
    if lib.vilba_shard_points(C.byref(cw), int(rank), int(world), C.byref(p0), C.byref(p1)) != 0:
        raise ValueError("bad rank/world")
    return int(p0.value), int(p1.value)


def shard_window(win, rank: int, world: int):
    """The sub-window `rank` solves: every key-frame and IMU edge, the points [p0, p1) with all their
    observations.  Returns (sub_window, p0, p1, e0, e1)."""
    p0, p1 = shard_points(win, rank, world)
    e0, e1 = int(win.pt_obs_begin[p0]), int(win.pt_obs_begin[p1])
    sub = dataclasses.replace(
        win, pt_xyz=win.pt_xyz[p0:p1].copy(), pt_obs_begin=(win.pt_obs_begin[p0:p1 + 1] - e0).astype(np.int32),
        obs_kf=win.obs_kf[e0:e1].copy(), obs_uv=win.obs_uv[e0:e1].copy(),
        obs_inv_sigma2=win.obs_inv_sigma2[e0:e1].copy(), truth={})
    return sub, p0, p1, e0, e1


def merge_sharded(win, parts):
    """Reassembles the result of a sharded solve.  `parts` = [(result, p0, p1, e0, e1)] in rank order; the
    key-frame states are identical on every rank, points / flags / chi2 are concatenated, the active-edge
    counts of the trace are summed."""
    from .capi import Result
    out = Result.alloc(win)
    first = parts[0][0]
    out.kf_state[:] = first.kf_state
    out.status, out.stage2_ran, out.solve_ms = first.status, first.stage2_ran, max(p[0].solve_ms for p in parts)
    out.n_outliers_stage1 = sum(p[0].n_outliers_stage1 for p in parts)
    for r, p0, p1, e0, e1 in parts:
        out.pt_xyz[p0:p1] = r.pt_xyz
        out.obs_outlier[e0:e1] = r.obs_outlier
        out.obs_chi2[e0:e1] = r.obs_chi2
    out.trace = [dict(t) for t in first.trace]
    for i, t in enumerate(out.trace):
        t["n_active_edges"] = sum(p[0].trace[i]["n_active_edges"] for p in parts)
    return out
