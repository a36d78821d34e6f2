#!/usr/bin/env python
"""Size sweep of the Black-Scholes job (run + calculate_greeks through the public API, host-resident results) from the
reference's smallest parallel size to the headline size, and across host widths (n_workers sets the number of Philox
streams: 8 per worker):  python tools/size_sweep.py > gpurun_out/size_sweep.jsonl"""
import json
import logging
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mcframework_b200 as mcf

logging.getLogger("mcframework_b200.core").setLevel(logging.WARNING)
BS = dict(S0=100.0, K=100.0, T=1.0, r=0.05, sigma=0.20)


def job(sim, n, workers):
    sim.set_seed(42)
    res = sim.run(n, parallel=True, n_workers=workers, compute_stats=False, n_steps=252, **BS)
    g = sim.calculate_greeks(n, parallel=True, n_steps=252, n_workers=workers, **BS)
    return res, g


def main():
    sim = mcf.BlackScholesSimulation()
    rows = [(n, 8) for n in (20_000, 100_000, 1_000_000, 10_000_000, 100_000_000)]
    rows += [(1_000_000, w) for w in (32, 128, 256, 512)] + [(100_000_000, 256)]
    for n, workers in rows:
        job(sim, n, workers)
        torch.cuda.synchronize()
        reps = 5 if n <= 10_000_000 else 2
        t0 = time.perf_counter()
        for _ in range(reps):
            res, g = job(sim, n, workers)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        print(json.dumps({"paths": n, "n_workers": workers, "streams": -(-n // max(1, n // (workers * 8))), "s_per_job": dt,
                          "path_steps_per_s": n * 252 / dt, "price": g["price"]}), flush=True)


if __name__ == "__main__":
    main()
