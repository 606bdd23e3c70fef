"""N>1 host logic on CPUs: world_size-2 gloo processes shard independent windows / IMU pairs with no
data-path collective, and the bench reductions (max of times, sum of work) behave."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mc_slam_b200 import sharding, synth


def test_shard_range_partitions():
    for n in (0, 1, 5, 8, 512, 4097):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle  # the CPU checker stands in for the GPU path in this CPU-only test

    # --- independent windows: rank r solves windows [b, e) of 3 ---
    b, e = sharding.shard_range(3, rank, world)
    finals = {}
    for i in range(b, e):
        w = synth.make_config("tiny", window_index=i)
        finals[i] = pyoracle.local_ba(w).trace[-1]["chi2_final"]
    # --- IMU pairs: rank r integrates its slice, no exchange ---
    batch = synth.make_imu_batch(n_pairs=10, seed=21, ragged=True)
    p0, p1, s0, s1 = sharding.shard_pairs(batch.sample_begin, rank, world)
    out = pyoracle.preintegrate_batch(batch.sample_begin[p0:p1 + 1] - batch.sample_begin[p0], batch.gyro[s0:s1],
                                      batch.acc[s0:s1], batch.dt[s0:s1], batch.bg[p0:p1], batch.ba[p0:p1])
    np.save(os.path.join(out_dir, f"pre_{rank}.npy"), out)
    np.save(os.path.join(out_dir, f"win_{rank}.npy"), np.array(sorted(finals.items())))
    # --- bench reductions ---
    tmax, tsum = sharding.reduce_bench([1.0 + rank], [10.0 * (rank + 1)])
    assert tmax == [float(world)] and tsum == [10.0 * world * (world + 1) / 2]
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path, oracle):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    # sharded == unsharded
    batch = synth.make_imu_batch(n_pairs=10, seed=21, ragged=True)
    ref = oracle.preintegrate_batch(batch.sample_begin, batch.gyro, batch.acc, batch.dt, batch.bg, batch.ba)
    got = np.concatenate([np.load(tmp_path / f"pre_{r}.npy") for r in range(world)])
    assert np.array_equal(got, ref)
    wins = np.concatenate([np.load(tmp_path / f"win_{r}.npy") for r in range(world)])
    assert [int(i) for i in wins[:, 0]] == [0, 1, 2]
    for i, chi in wins:
        assert chi == oracle.local_ba(synth.make_config("tiny", window_index=int(i))).trace[-1]["chi2_final"]
