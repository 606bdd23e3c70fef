#!/usr/bin/env python
"""One rank of a point-sharded local-BA solve (BASELINE config 4); launched by torchrun, one process
per GPU.  Checks the merged result against the CPU oracle on the whole window and prints the device-timed
throughput (max over ranks).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/sharded_worker.py --config c4 [--check] [--steps K]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="small")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--keep-points", type=int, default=-1, help="keep only the first N map points (edge case: empty shards)")
    ap.add_argument("--then", default="", help="afterwards solve this other config on the SAME context (different P, E, n_free) and check it")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from mc_slam_b200 import api, sharding, synth

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(api.comm_unique_id()), dtype=torch.uint8))
    if world > 1:
        dist.broadcast(uid, 0)
    ctx = api.Context(local_rank)
    ctx.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)

    win = synth.make_config(args.config)
    if args.keep_points >= 0:
        import dataclasses
        e1 = int(win.pt_obs_begin[args.keep_points])
        win = dataclasses.replace(win, pt_xyz=win.pt_xyz[:args.keep_points].copy(), pt_obs_begin=win.pt_obs_begin[:args.keep_points + 1].copy(),
                                  obs_kf=win.obs_kf[:e1].copy(), obs_uv=win.obs_uv[:e1].copy(),
                                  obs_inv_sigma2=win.obs_inv_sigma2[:e1].copy(), truth={})
    sub, p0, p1, e0, e1 = sharding.shard_window(win, rank, world)
    res = None
    ms, iters = 0.0, 0
    for it in range(args.warmup + args.steps):
        if world > 1:
            dist.barrier()
        res = ctx.local_ba(sub)
        if it >= args.warmup:
            ms += res.solve_ms
            iters += len(res.trace)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    parts = [None] * world
    mine = (res, p0, p1, e0, e1)
    if world > 1:
        dist.all_gather_object(parts, mine)
    else:
        parts = [mine]
    ok = True
    if rank == 0:
        merged = sharding.merge_sharded(win, parts)
        for r, *_ in parts[1:]:  # the reduced system is solved redundantly: identical key-frame states
            ok = ok and np.array_equal(r.kf_state, parts[0][0].kf_state)
            ok = ok and [t_["trials"] for t_ in r.trace] == [t_["trials"] for t_ in parts[0][0].trace]
        line = {"config": args.config, "world": world, "lm_iters_per_sec": iters / (float(t.item()) * 1e-3),
                "ms_per_solve": float(t.item()) / args.steps, "lm_iters": len(merged.trace),
                "points_per_rank": [int(p[2] - p[1]) for p in parts], "edges_per_rank": [int(p[4] - p[3]) for p in parts],
                "ranks_identical": bool(ok)}
        if args.check:
            from oracle import pyoracle
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from parity_util import compare as _compare
            o = pyoracle.local_ba(win)
            _compare(merged, o, win)
            line["parity"] = "ok"
        print("SHARDED " + json.dumps(line), flush=True)
    if args.then:  # a second window of another shape on the same communicator context (the normal SLAM use)
        win2 = synth.make_config(args.then)
        sub2, q0, q1, f0, f1 = sharding.shard_window(win2, rank, world)
        if world > 1:
            dist.barrier()
        res2 = ctx.local_ba(sub2)
        parts2 = [None] * world
        if world > 1:
            dist.all_gather_object(parts2, (res2, q0, q1, f0, f1))
        else:
            parts2 = [(res2, q0, q1, f0, f1)]
        if rank == 0:
            from oracle import pyoracle
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from parity_util import compare as _compare
            _compare(sharding.merge_sharded(win2, parts2), pyoracle.local_ba(win2), win2)
            print("SHARDED_THEN " + json.dumps({"config": args.then, "parity": "ok"}), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    if not ok:
        raise SystemExit(3)


if __name__ == "__main__":
    main()
