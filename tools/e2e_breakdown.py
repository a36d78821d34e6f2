#!/usr/bin/env python
"""Where the end-to-end time of bench.py's public-API step goes (host wall clock with syncs; development aid)."""
import cProfile
import io
import logging
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mcframework_b200 as mcf

logging.getLogger("mcframework_b200.core").setLevel(logging.WARNING)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
BS = dict(S0=100.0, K=100.0, T=1.0, r=0.05, sigma=0.20)
sim = mcf.BlackScholesSimulation()
fw = mcf.MonteCarloFramework()
fw.register_simulation(sim)


def step():
    sim.set_seed(42)
    t0 = time.perf_counter()
    res = fw.run_simulation(sim.name, n, parallel=True, n_workers=8, compute_stats=False, n_steps=252, **BS)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    g = sim.calculate_greeks(n, parallel=True, n_steps=252, n_workers=8, **BS)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, res.execution_time


for _ in range(3):
    before = torch.cuda.memory_stats()
    out = step()
    after = torch.cuda.memory_stats()
    print("run_simulation %.3f s (execution_time %.3f), calculate_greeks %.3f s" % tuple(out[i] for i in (0, 2, 1)),
          "| cudaMalloc calls:", after.get("num_device_alloc", 0) - before.get("num_device_alloc", 0), "cudaFree:",
          after.get("num_device_free", 0) - before.get("num_device_free", 0), "segments:",
          after.get("segment.all.allocated", 0) - before.get("segment.all.allocated", 0), "retries:",
          after.get("num_alloc_retries", 0) - before.get("num_alloc_retries", 0),
          "reserved GB: %.1f" % (after.get("reserved_bytes.all.current", 0) / 1e9))

# device-level pieces with explicit syncs
from mcframework_b200 import _device as D, _native as N, _streams as S
import numpy as np

lay = S.parallel_layout(np.random.SeedSequence(42), n, 8)
dl = D.DeviceLayout.of(lay)


def timed(label, fn):
    torch.cuda.synchronize()
    t = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    print("  %-28s %.1f ms" % (label, (time.perf_counter() - t) * 1e3))
    return r


for _ in range(2):
    plan = timed("normal_plan", lambda: D.normal_plan(dl, 252))
    one = timed("gbm_terminal x1", lambda: D.gbm_terminal(dl, 252, plan, [(N.BS_EURO_CALL, [100.0, 100.0, 1.0, 0.05, 0.2])]))
    eight = timed("gbm_terminal x8", lambda: D.gbm_terminal(dl, 252, plan, [(N.BS_EURO_CALL, [100.0 + k, 100.0, 1.0, 0.05, 0.2]) for k in range(8)]))
    host = timed("D2H 800 MB (pinned)", lambda: torch.empty(one.shape, dtype=torch.float64, pin_memory=True).copy_(one))
    host2 = timed("D2H 800 MB (.cpu())", lambda: one.cpu())
    del plan, one, eight, host, host2
pr = cProfile.Profile()
pr.enable()
step()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(25)
print(s.getvalue()[:6000])
