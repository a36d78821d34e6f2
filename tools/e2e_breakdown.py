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


for _ in range(2):
    print("run_simulation %.3f s (execution_time %.3f), calculate_greeks %.3f s" % tuple(step()[i] for i in (0, 2, 1)))
pr = cProfile.Profile()
pr.enable()
step()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(25)
print(s.getvalue()[:6000])
