#!/usr/bin/env python
"""Small driver for ncu captures of the statistics / LSM kernels at BASELINE sizes (development aid).

    python tools/prof_stats.py boot|select|lsm
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import mcframework_b200 as mcf
from mcframework_b200 import _device as D, _stats_device as SD, _streams as S

what = sys.argv[1] if len(sys.argv) > 1 else "boot"
if what in ("boot", "select"):
    n = 100_000_000
    x = torch.from_numpy(np.random.default_rng(0).normal(5, 2, n)).cuda()
    if what == "boot":
        m = SD.bootstrap_means(x, 32, np.random.default_rng(7))
        print(float(m.mean()))
    else:
        print(SD.percentiles(x, (1, 5, 50, 95, 99)))
else:
    n, steps = 10_000_000, 252
    sim = mcf.BlackScholesPathSimulation()
    sim.set_seed(1)
    dl = D.DeviceLayout.of(S.sequential_layout(sim.rng, n))
    plan = D.normal_plan(dl, steps)
    paths = D.gbm_paths(dl, steps, plan, 100.0, 0.05, 0.2, 1.0, time_major=True)
    del plan
    print(float(SD.lsm_price(paths, 100.0, 0.05, 1.0 / steps, True).cpu()[0]))
torch.cuda.synchronize()
