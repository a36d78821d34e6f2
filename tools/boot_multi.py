#!/usr/bin/env python
"""BASELINE config 5 sharded: StatsEngine on 1e8 values (moments, 5 percentiles, BCa bootstrap) over N GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 \\
        tools/boot_multi.py [n] [B]

Every rank holds all of x (the bootstrap gathers from the whole population); the SLOT RANGE of the single index stream
is sharded (all-gather of accepted counts, all-reduce of the B sums).  Moments and percentiles run on each rank's shard
of x with the small collectives of DESIGN.md section 5.  Rank 0 prints one JSON line.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from mcframework_b200 import _stats_device as SD

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
group = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    group = dist.group.WORLD
x = torch.from_numpy(np.random.default_rng(0).normal(5, 2, n)).cuda()
shard = x[rank * n // world:(rank + 1) * n // world].contiguous()


def timed(fn, reps=2):
    best, out = 1e30, None
    for _ in range(reps):
        if group is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        if group is not None:
            dist.barrier()
        best = min(best, time.perf_counter() - t)
    return best, out


t_m, m = timed(lambda: SD.moments(shard, 5.0, group))
t_p, p = timed(lambda: SD.percentiles(shard, (1, 5, 50, 95, 99), n_valid=n, group=group))
t_b, means = timed(lambda: SD.bootstrap_means(x, B, np.random.default_rng(7), group=group), reps=1)
if rank == 0:
    print(json.dumps({"n": n, "B": B, "n_gpus": world, "moments_s": t_m, "mean": float(m.mean), "percentiles_s": t_p,
                      "percentiles": [float(v) for v in p], "bootstrap_s": t_b, "resamples_per_s": B / t_b,
                      "gathers_per_s": B * n / t_b, "mean_of_means": float(means.mean())}))
if group is not None:
    dist.destroy_process_group()
