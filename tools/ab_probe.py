#!/usr/bin/env python
"""Kernel-group timings of whatever library MCF_B200_LIB names (default: the in-tree build): one JSON line.
Development aid for tools/ab_variants.sh-style comparisons; best of `reps` with CUDA events on the launch stream."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mcframework_b200 import _device as D, _native as N, _streams as S


def best(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
        del out
    return min(ts)


def main():
    label = sys.argv[1] if len(sys.argv) > 1 else "lib"
    n, steps = int(float(os.environ.get("AB_PATHS", "1e8"))), 252
    P = [100.0, 100.0, 1.0, 0.05, 0.2]
    sets = [(N.BS_EURO_CALL, [100.0 + k, 100.0, 1.0, 0.05, 0.2]) for k in range(8)]
    dl = D.DeviceLayout.of(S.parallel_layout(np.random.SeedSequence(42), n, 8))
    out = {"label": label, "plan_ms": best(lambda: D.normal_plan(dl, steps))}
    plan = D.normal_plan(dl, steps)
    one = D.gbm_terminal(dl, steps, plan, sets[:1])
    out["price"] = float(one[0].mean().item())
    del one
    out["replay1_ms"] = best(lambda: D.gbm_terminal(dl, steps, plan, sets[:1]))
    out["replay8_ms"] = best(lambda: D.gbm_terminal(dl, steps, plan, sets))
    del plan
    n_pi = 2_000_000
    dlp = D.DeviceLayout.of(S.parallel_layout(np.random.SeedSequence(1), n_pi, 8))
    t = best(lambda: D.pi_hits(dlp, 5000))
    out["pi_samples_per_s"] = n_pi * 5000 / (t * 1e-3)
    out["pi_hits_sum"] = int(D.pi_hits(dlp, 5000)[0].sum().item())
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
