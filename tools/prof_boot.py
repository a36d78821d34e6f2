#!/usr/bin/env python
"""Binned bootstrap at n = 1e8: single-pass form against the two-pass form (second call of each, allocations warm)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mcframework_b200 import _stats_device as SD

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
g = torch.Generator(device="cuda")
g.manual_seed(1)
x = torch.empty(n, dtype=torch.float64, device="cuda").normal_(5.0, 2.0, generator=g)
forms = {"single": (True, True), "two": (False, False)}.get(sys.argv[3] if len(sys.argv) > 3 else "", (True, False, True, False))
for form in forms:
    before = dict(SD.BOOT_FORMS)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m = SD.bootstrap_means(x, B, np.random.default_rng(7), single_pass=form)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("single_pass" if form else "two_pass", "%.3f s  %.1f resamples/s" % (dt, B / dt), {k: SD.BOOT_FORMS[k] - before[k] for k in before},
          float(m.mean()), flush=True)
