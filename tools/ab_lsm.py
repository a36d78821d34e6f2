#!/usr/bin/env python
"""Longstaff-Schwartz sweep timing of whatever library MCF_B200_LIB names: one JSON line (development aid)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mcframework_b200 as mcf
from mcframework_b200 import _stats_device as SD


def main():
    label = sys.argv[1] if len(sys.argv) > 1 else "lib"
    n, steps = int(float(os.environ.get("AB_PATHS", "1e7"))), 252
    sim = mcf.BlackScholesPathSimulation()
    sim.set_seed(1)
    paths = sim.simulate_paths(n, n_steps=steps, device=True, time_major=True)
    ts = []
    for _ in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        price = SD.lsm_price(paths, 100.0, 0.05, 1.0 / steps, True)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = min(ts[1:])
    print(json.dumps({"label": label, "lsm_ms": ms, "GBps_algorithmic": (steps - 1) * n * 24 / ms / 1e6, "price": float(price.cpu()[0])}), flush=True)


if __name__ == "__main__":
    main()
