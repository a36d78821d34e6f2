#!/usr/bin/env python
"""Per-config measurements for BASELINE.json configs 1, 3, 4, 5 on ONE GPU (config 2 is bench.py's headline).

    python tools/bench_configs.py [--scale 1.0] [--only pi,portfolio,lsm,stats] > gpurun_out/configs.json

CUDA events on the launch stream, 1 warm-up + best of 2 (these are records for profiles/, not the headline line).
``--scale`` shrinks the problem sizes (development runs).  Every number is a device-resident rate unless it says e2e.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import mcframework_b200 as mcf
from mcframework_b200 import _device as D, _native as N, _stats_device as SD, _streams as S

PEAK_HBM = 6553.9
try:
    PEAK_HBM = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:  # noqa: BLE001
    pass


def timed(fn, reps=2, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    out = None
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best, out


def cfg_pi(scale):
    import logging
    logging.getLogger("mcframework_b200.core").setLevel(logging.WARNING)
    out = {}
    sim = mcf.PiEstimationSimulation()
    fw = mcf.MonteCarloFramework()
    fw.register_simulation(sim)

    def as_is():
        sim.set_seed(123)
        return fw.run_simulation("Pi Estimation", 10_000, n_points=5000, parallel=True)

    as_is()
    t0 = time.perf_counter()
    res = as_is()
    wall = time.perf_counter() - t0
    out["as_is_e2e"] = {"wall_s": wall, "samples_per_s": 5e7 / wall, "sum_hits": int(round(res.results.sum() * 5000 / 4)),
                        "includes": "default 12-metric engine (bootstrap B=10000 over n=10000) + D2H"}
    t0 = time.perf_counter()
    sim.set_seed(123)
    res = fw.run_simulation("Pi Estimation", 10_000, n_points=5000, parallel=True, compute_stats=False)
    out["as_is_e2e_no_stats"] = {"wall_s": time.perf_counter() - t0, "samples_per_s": 5e7 / (time.perf_counter() - t0)}
    n = int(2e6 * scale)
    for name, lay in (("scaled_philox_blocks", S.parallel_layout(np.random.SeedSequence(123), n, 8)),
                      ("scaled_pcg64_sequential", S.sequential_layout(np.random.default_rng(123), n))):
        dl = D.DeviceLayout.of(lay)
        t, _ = timed(lambda: D.pi_hits(dl, 5000))
        out[name] = {"sims": n, "n_points": 5000, "s": t, "samples_per_s": n * 5000 / t}
    return out


def cfg_portfolio(scale):
    n, steps, stride = int(1e7 * scale), 2520, 21
    lay = S.parallel_layout(np.random.SeedSequence(42), n, 8)
    dl = D.DeviceLayout.of(lay)
    tp, plan = timed(lambda: D.normal_plan(dl, steps), reps=1)
    tt, _ = timed(lambda: D.gbm_terminal(dl, steps, plan, [(N.PORTFOLIO_GBM, [1e4, 0.07, 0.2])]), reps=1)
    ts, sl = timed(lambda: D.gbm_slices(dl, steps, plan, N.PORTFOLIO_GBM, [1e4, 0.07, 0.2], stride, True), reps=1)
    nbytes = sl.numel() * 8
    return {"paths": n, "steps": steps, "slice_stride": stride, "n_slices": int(sl.shape[0]), "plan_s": tp, "terminal_s": tt,
            "slices_s": ts, "path_steps_per_s_plan_plus_slices": n * steps / (tp + ts),
            "path_steps_per_s_plan_plus_terminal": n * steps / (tp + tt),
            "slice_store_GBps": nbytes / ts / 1e9, "slice_store_frac_of_hbm_peak": nbytes / ts / 1e9 / PEAK_HBM,
            "mean_terminal": float(SD.moments(sl[-1]).mean)}


def cfg_lsm(scale):
    n, steps = int(1e7 * scale), 252
    sim = mcf.BlackScholesPathSimulation()
    sim.set_seed(1)
    lay = S.sequential_layout(sim.rng, n)
    dl = D.DeviceLayout.of(lay)
    tp, plan = timed(lambda: D.normal_plan(dl, steps), reps=1)
    tg, paths = timed(lambda: D.gbm_paths(dl, steps, plan, 100.0, 0.05, 0.2, 1.0, time_major=True), reps=1, warm=0)
    del plan
    tl, price = timed(lambda: SD.lsm_price(paths, 100.0, 0.05, 1.0 / steps, True), reps=1)
    pbytes = paths.numel() * 8
    alg = steps * n * 24  # fused step kernel: S_t (decision) + S_{t-1} (regression) + discounted cash flow, 8 B each
    return {"paths": n, "steps": steps, "plan_s": tp, "paths_s": tg, "paths_store_GBps": pbytes / tg / 1e9, "lsm_s": tl,
            "lsm_path_steps_per_s": n * steps / tl, "lsm_algorithmic_GBps": alg / tl / 1e9,
            "lsm_frac_of_hbm_peak": alg / tl / 1e9 / PEAK_HBM, "american_put_price": float(price.cpu()[0])}


def cfg_stats(scale):
    n, B = int(1e8 * scale), int(max(100, 1e4 * scale))
    x = torch.from_numpy(np.random.default_rng(0).normal(5, 2, n)).cuda()
    out = {"n": n, "B": B}
    t, m = timed(lambda: SD.moments_raw(x, 5.0))
    out["moments"] = {"s": t, "GBps": n * 8 / t / 1e9, "frac_of_hbm_peak": n * 8 / t / 1e9 / PEAK_HBM}
    t, p = timed(lambda: SD.percentiles(x, (1, 5, 50, 95, 99)))
    out["percentiles"] = {"s": t, "algorithmic_GBps_8_passes": 8 * n * 8 / t / 1e9, "values": [float(v) for v in p]}
    gen = np.random.default_rng(7)
    t0 = time.perf_counter()
    means = SD.bootstrap_means(x, B, gen)
    torch.cuda.synchronize()
    tb = time.perf_counter() - t0
    out["bootstrap"] = {"s": tb, "resamples_per_s": B / tb, "gathers_per_s": B * n / tb,
                        "gather_GBps_algorithmic_8B": B * n * 8 / tb / 1e9, "gather_GBps_32B_sectors": B * n * 32 / tb / 1e9}
    eng = mcf.stats_engine.build_default_engine()
    ctx = mcf.StatsContext(n=n, percentiles=(1, 5, 50, 95, 99), bootstrap="bca", n_bootstrap=B, rng=7, target=5.0, eps=0.01)
    t0 = time.perf_counter()
    res = eng.compute(mcf.stats_engine.DeviceSample(x), ctx)
    out["engine_all_12_metrics_s"] = time.perf_counter() - t0
    out["bca_ci"] = res.metrics.get("ci_mean_bootstrap")
    out["mean_of_means"] = float(SD.moments(means).mean)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--only", default="pi,portfolio,lsm,stats")
    a = ap.parse_args()
    N.lib()
    out = {"gpu": torch.cuda.get_device_name(0), "scale": a.scale, "hbm_peak_GBps": PEAK_HBM}
    table = {"pi": cfg_pi, "portfolio": cfg_portfolio, "lsm": cfg_lsm, "stats": cfg_stats}
    for name in a.only.split(","):
        try:
            out[name] = table[name](a.scale)
        except Exception as exc:  # noqa: BLE001
            out[name] = {"error": repr(exc)}
        torch.cuda.empty_cache()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
