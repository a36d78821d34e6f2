#!/usr/bin/env python
"""Plain vs binned bootstrap gather at full population size (development aid)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mcframework_b200 import _stats_device as SD

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 200
variants = sys.argv[3].split(",") if len(sys.argv) > 3 else ["plain", "s22", "s23", "s21"]
x = torch.from_numpy(np.random.default_rng(0).normal(5, 2, n)).cuda()
KW = {"plain": dict(binned=False), "s22": dict(binned=True), "s23": dict(binned=True, region_shift=23),
      "s21": dict(binned=True, region_shift=21), "rpb16": dict(binned=True, resamples_per_batch=16)}
for label in variants:
    kw = KW[label]
    for rep in range(2):
        g = np.random.default_rng(7)
        torch.cuda.synchronize()
        t = time.perf_counter()
        m = SD.bootstrap_means(x, B, g, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
    print(f"{label:8s} n={n} B={B}: {dt * 1e3:8.1f} ms  {B / dt:8.1f} resamples/s  {B * n / dt / 1e9:6.1f} G gathers/s  mean of means {float(m.mean()):.12f}")
