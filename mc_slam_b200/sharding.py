"""Host-side sharding of the path over the GPUs of one node (one process per GPU).

Independent windows (BASELINE config 5) and key-frame pairs (config 2) partition with NO data-path
collective: every rank owns a contiguous slice; torch.distributed only carries the barrier and the
max-over-ranks / sum-over-ranks reductions of the bench (works with the gloo backend on CPUs, which is how
the logic is tested without GPUs)."""
from __future__ import annotations

from typing import List, Sequence, Tuple


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
