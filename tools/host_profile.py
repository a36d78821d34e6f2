#!/usr/bin/env python
"""cProfile of the public-API Black-Scholes job (run + calculate_greeks) at a small size: where the HOST time goes.
    python tools/host_profile.py [paths] [n_workers]"""
import cProfile
import logging
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mcframework_b200 as mcf

logging.getLogger("mcframework_b200.core").setLevel(logging.WARNING)
BS = dict(S0=100.0, K=100.0, T=1.0, r=0.05, sigma=0.20)


def job(sim, n, workers):
    sim.set_seed(42)
    sim.run(n, parallel=True, n_workers=workers, compute_stats=False, n_steps=252, **BS)
    return sim.calculate_greeks(n, parallel=True, n_steps=252, n_workers=workers, **BS)


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
    w = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    sim = mcf.BlackScholesSimulation()
    for _ in range(3):
        job(sim, n, w)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(20):
        job(sim, n, w)
    torch.cuda.synchronize()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("cumulative").print_stats(35)


if __name__ == "__main__":
    main()
