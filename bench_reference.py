"""CPU arm of the benchmark: the reference's own implementation of the hot path, timed on the host cores.

``load()`` returns the UNMODIFIED reference package when ``baseline/_ref`` holds an install of it (made by
``__graft_entry__.build()`` from ``/root/reference`` with ``pip install --no-index --no-deps --target baseline/_ref``;
git-ignored, it travels to the GPU box with the snapshot) -- ``kind = "reference"`` -- and otherwise the oracle's NumPy
restatement ``oracle/np_port.py`` -- ``kind = "port"`` (bit-identical results, ~2.5x faster than the reference because it
skips the reference's logging/validation layers; measured in DESIGN.md).  This file and ``bench.py`` are the only
places outside ``tests/`` that may execute anything under ``oracle/`` or ``baseline/_ref``, and only as the baseline.

Every function runs ONE bounded sample of a BASELINE.json config and returns ``(seconds, units, sample_text, check)``.
"""
from __future__ import annotations

import logging
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")

_ref = None


def load():
    """-> ("reference", package) or ("port", oracle.np_port)"""
    global _ref
    if _ref is not None:
        return _ref
    if os.path.isdir(os.path.join(REF_DIR, "mcframework")):
        sys.path.insert(0, REF_DIR)
        try:
            import mcframework  # the reference, unmodified

            for name in ("mcframework.core", "mcframework.stats_engine", "mcframework"):
                logging.getLogger(name).setLevel(logging.WARNING)
            _ref = ("reference", mcframework)
            return _ref
        except Exception:  # noqa: BLE001 - fall back to the port below
            pass
        finally:
            sys.path.remove(REF_DIR)
    from oracle import np_port

    _ref = ("port", np_port)
    return _ref


def cores() -> int:
    return os.cpu_count() or 1


BS = dict(S0=100.0, K=100.0, T=1.0, r=0.05, sigma=0.20)


# ------------------------------------------------------------------------------------------------ cfg 2 (headline)
def black_scholes_step(n_paths: int, n_workers: int, n_steps: int = 252):
    """run(parallel=True) + calculate_greeks (8 revaluations), the job of one bench step."""
    kind, m = load()
    t0 = time.perf_counter()
    if kind == "reference":
        sim = m.BlackScholesSimulation()
        sim.set_seed(42)
        res = sim.run(n_paths, parallel=True, n_workers=n_workers, compute_stats=False, n_steps=n_steps, **BS)
        greeks = sim.calculate_greeks(n_paths, parallel=True, n_steps=n_steps, **BS)  # n_workers = cpu_count inside
        price = float(res.mean)
    else:
        run = m.Runner("black_scholes", seed=42)
        payoffs = run.run(n_paths, parallel=True, n_workers=n_workers, n_steps=n_steps, **BS)
        greeks = m.greeks(n_paths, parallel=True, n_workers=n_workers, n_steps=n_steps, **BS)
        price = float(np.mean(payoffs))
    dt = time.perf_counter() - t0
    return dt, n_paths * n_steps, f"{n_paths} paths x {n_steps} steps: run(parallel=True) + calculate_greeks, {n_workers} threads", \
        {"price": price, "delta": float(greeks["delta"])}


# ------------------------------------------------------------------------------------------------ cfg 1
def pi_as_is(n_sims: int = 10_000, n_points: int = 5000):
    """README.md:86-93 as it stands: parallel=True below the threshold = the sequential PCG64 loop."""
    kind, m = load()
    t0 = time.perf_counter()
    if kind == "reference":
        sim = m.PiEstimationSimulation()
        sim.set_seed(123)
        fw = m.MonteCarloFramework()
        fw.register_simulation(sim)
        res = fw.run_simulation("Pi Estimation", n_sims, n_points=n_points, parallel=True, compute_stats=False)
        vals = res.results
    else:
        vals = m.Runner("pi", seed=123).run(n_sims, parallel=True, n_points=n_points)
    dt = time.perf_counter() - t0
    hits = int(round(float(np.sum(vals)) * n_points / 4.0))
    return dt, n_sims * n_points, f"{n_sims} sims x {n_points} points, seed 123, parallel=True (stats excluded), as the README runs it", \
        {"sum_hits": hits}


def pi_parallel(n_sims: int = 40_000, n_points: int = 5000):
    """The reference's block-parallel path (n >= 20 000) on all host cores."""
    kind, m = load()
    w = cores()
    t0 = time.perf_counter()
    if kind == "reference":
        sim = m.PiEstimationSimulation()
        sim.set_seed(123)
        vals = sim.run(n_sims, n_points=n_points, parallel=True, n_workers=w, compute_stats=False).results
    else:
        vals = m.Runner("pi", seed=123).run(n_sims, parallel=True, n_workers=w, n_points=n_points)
    dt = time.perf_counter() - t0
    return dt, n_sims * n_points, f"{n_sims} sims x {n_points} points, parallel=True on {w} threads (stats excluded)", \
        {"mean": float(np.mean(vals))}


# ------------------------------------------------------------------------------------------------ cfg 3
def portfolio(n_paths: int = 40_000, years: int = 10):
    """The reference stores no path slices (sims/portfolio.py:77-84): its terminal-wealth run is the comparable job."""
    kind, m = load()
    w = cores()
    t0 = time.perf_counter()
    if kind == "reference":
        sim = m.PortfolioSimulation()
        sim.set_seed(42)
        vals = sim.run(n_paths, parallel=True, n_workers=w, compute_stats=False, years=years).results
    else:
        vals = m.Runner("portfolio", seed=42).run(n_paths, parallel=True, n_workers=w, years=years)
    dt = time.perf_counter() - t0
    steps = int(years / (1.0 / 252.0))
    return dt, n_paths * steps, f"{n_paths} paths x {steps} steps terminal wealth (the reference stores no slices), parallel=True on {w} threads", \
        {"mean": float(np.mean(vals))}


# ------------------------------------------------------------------------------------------------ cfg 4
def lsm(n_paths: int = 20_000, n_steps: int = 252):
    """simulate_paths + _american_exercise_lsm (single thread in the reference); only the LSM sweep is the unit."""
    kind, m = load()
    if kind == "reference":
        from mcframework.sims.black_scholes import _american_exercise_lsm

        sim = m.BlackScholesPathSimulation()
        sim.set_seed(1)
        t0 = time.perf_counter()
        paths = sim.simulate_paths(n_paths, n_steps=n_steps)
        t_paths = time.perf_counter() - t0
        t0 = time.perf_counter()
        price = _american_exercise_lsm(paths, 100.0, 0.05, 1.0 / n_steps, "put")
        dt = time.perf_counter() - t0
    else:
        run = m.Runner("path_terminal", seed=1)
        t0 = time.perf_counter()
        paths = run.simulate_paths(n_paths, n_steps=n_steps)
        t_paths = time.perf_counter() - t0
        t0 = time.perf_counter()
        price, _ = m.lsm_price(paths, 100.0, 0.05, 1.0 / n_steps, "put")
        dt = time.perf_counter() - t0
    return dt, n_paths * n_steps, f"{n_paths} paths x {n_steps} steps, American put, LSM sweep only (1 thread; simulate_paths took {t_paths:.2f} s)", \
        {"price": float(price), "simulate_paths_s": t_paths}


# ------------------------------------------------------------------------------------------------ cfg 5
def bootstrap(n: int = 100_000, n_resamples: int = 1000):
    """ci_mean_bootstrap (BCa) -- stats_engine.py:1202-1301 -- on a reduced sample: the reference materialises (B, n)."""
    kind, m = load()
    x = np.random.default_rng(0).normal(5.0, 2.0, n)
    t0 = time.perf_counter()
    if kind == "reference":
        from mcframework.stats_engine import StatsContext, ci_mean_bootstrap

        ctx = StatsContext(n=n, percentiles=(1, 5, 50, 95, 99), n_bootstrap=n_resamples, bootstrap="bca", rng=7)
        ci = ci_mean_bootstrap(x, ctx)
        lo, hi = float(ci["low"]), float(ci["high"])
    else:
        means = m.bootstrap_means(x, n_resamples, np.random.default_rng(7))
        lo, hi = m.bca_interval(x, means, 0.95)
    dt = time.perf_counter() - t0
    return dt, n_resamples, f"BCa bootstrap, n = {n}, B = {n_resamples} (1 thread; the reference cannot allocate n = 1e8)", \
        {"low": float(lo), "high": float(hi)}


def percentiles(n: int = 10_000_000):
    kind, m = load()
    x = np.random.default_rng(0).normal(5.0, 2.0, n)
    t0 = time.perf_counter()
    if kind == "reference":
        from mcframework.stats_engine import StatsContext, percentiles as pct

        out = pct(x, StatsContext(n=n, percentiles=(1, 5, 50, 95, 99)))
    else:
        out = dict(zip((1, 5, 50, 95, 99), np.percentile(x, (1, 5, 50, 95, 99))))
    dt = time.perf_counter() - t0
    return dt, n, f"percentiles (1,5,50,95,99) of {n} values (1 thread)", {"p50": float(out[50])}
