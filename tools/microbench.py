#!/usr/bin/env python
"""Kernel-level timings on one GPU (CUDA events, current stream). Not the headline bench; a development aid."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mcframework_b200 import _device as D, _native as N, _streams as S


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best


def main():
    out = {}
    n_pi = int(float(os.environ.get("MB_PI", "2e6")))
    for name, lay in (("pi_pcg64", S.sequential_layout(np.random.default_rng(1), n_pi)),
                      ("pi_philox", S.parallel_layout(np.random.SeedSequence(1), n_pi, 8))):
        dl = D.DeviceLayout.of(lay)
        t = timed(lambda: D.pi_hits(dl, 5000))
        out[name] = {"sims": n_pi, "n_points": 5000, "s": t, "samples_per_s": n_pi * 5000 / t}
    n = int(float(os.environ.get("MB_GBM", "1e7")))
    for name, lay in (("gbm_philox", S.parallel_layout(np.random.SeedSequence(42), n, 8)),
                      ("gbm_pcg64", S.sequential_layout(np.random.default_rng(42), n // 4))):
        dl = D.DeviceLayout.of(lay)
        nn = lay.n_total
        tp = timed(lambda: D.normal_plan(dl, 252))
        plan = D.normal_plan(dl, 252)
        ts = timed(lambda: D.gbm_terminal(dl, 252, plan, [(N.BS_EURO_CALL, [100, 100, 1, .05, .2])]))
        t8 = timed(lambda: D.gbm_terminal(dl, 252, plan, [(N.BS_EURO_CALL, [100 + k, 100, 1, .05, .2]) for k in range(8)]))
        out[name] = {"paths": nn, "steps": 252, "plan_s": tp, "sim_s": ts, "sim8_s": t8,
                     "path_steps_per_s_total": nn * 252 / (tp + ts), "path_steps_per_s_sim_only": nn * 252 / ts}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
