#!/usr/bin/env python
"""Sharded Longstaff-Schwartz sweep under torchrun: fused (peer-memory exchange) vs staged (NCCL all-reduce per step)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import mcframework_b200 as mcf
from mcframework_b200 import _stats_device as SD


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_total, steps = int(float(os.environ.get("AB_PATHS", "1e7"))), 252
    n = n_total // world
    sim = mcf.BlackScholesPathSimulation()
    sim.set_seed(1 + 1000 * rank)
    paths = sim.simulate_paths(n, n_steps=steps, device=True, time_major=True)
    out = {"world": world, "paths_total": n * world}
    for name, fused in (("fused_peer", True), ("staged_nccl", False)):
        ts = []
        for _ in range(4):
            dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            price = SD.lsm_price(paths, 100.0, 0.05, 1.0 / steps, True, group=dist.group.WORLD, fused=fused)
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b)], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(float(t))
        out[name] = {"ms": min(ts[1:]), "price": float(price.cpu()[0])}
    out["peer_exchange_ok"] = SD.LsmPeerExchange.get(dist.group.WORLD).ok
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
