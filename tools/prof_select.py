#!/usr/bin/env python
"""Percentiles of 1e8 values: wall time of the public call and (under ncu) its launch list.  python tools/prof_select.py [n]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mcframework_b200 import _stats_device as SD

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
g = torch.Generator(device="cuda")
g.manual_seed(1)
x = torch.empty(n, dtype=torch.float64, device="cuda").normal_(5.0, 2.0, generator=g)
pcts = (1, 5, 50, 95, 99)
for mode in ("bracketed", "radix"):
    SD.SELECT_BRACKET_MIN_N = (1 << 23) if mode == "bracketed" else (1 << 62)
    SD.percentiles(x, pcts)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        v = SD.percentiles(x, pcts)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    print(mode, "best %.3f ms" % (min(ts) * 1e3), [float(a) for a in v], flush=True)
