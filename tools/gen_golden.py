#!/usr/bin/env python
"""Generate golden fixtures by running the UNMODIFIED reference (build container only).

    python tools/gen_golden.py            # needs /root/reference/src (read-only mount)

Writes ``tests/golden/reference_golden.json`` (scalars, dicts, provenance) and
``tests/golden/reference_golden.npz`` (result vectors).  The reference has no golden vectors
of its own for this path, so parity is pinned by these outputs; the
NumPy/SciPy versions that produced them are recorded because ``Generator`` streams are only
guaranteed stable per NumPy feature release (NEP 19).

Nothing under tests/, bench.py or smoke() reads /root/reference at run time; they read these
two files.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import scipy

REF_SRC = os.environ.get("MCF_REFERENCE_SRC", "/root/reference/src")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF_SRC)

import logging  # noqa: E402

import mcframework as ref  # noqa: E402
from mcframework.core import MonteCarloSimulation  # noqa: E402
from mcframework.sims import _american_exercise_lsm  # noqa: E402
from mcframework.stats_engine import StatsContext, build_default_engine  # noqa: E402

logging.getLogger("mcframework.core").setLevel(logging.WARNING)


class Dice(MonteCarloSimulation):
    def single_simulation(self, _rng=None, **kw):
        rng = self._rng(_rng, self.rng)
        return float(rng.integers(1, 7, size=2).sum())


def jsonable(o):
    if isinstance(o, dict):
        return {str(k): jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [jsonable(v) for v in o]
    if isinstance(o, (np.floating, np.integer)):
        return o.item()
    if isinstance(o, np.ndarray):
        return o.tolist()
    return o


def lsm_large(provenance, n_paths=1_000_000, n_steps=252):
    """G9: Longstaff-Schwartz at 1e6 paths x 252 steps on the UNMODIFIED reference (black_scholes.py:109-193, 403-418):
    ``simulate_paths(1_000_000)`` with seed 1, American put K=100.  The reference returns the price only; the exercise-step
    histogram comes from the oracle port run on the SAME matrix, after its price has been checked to equal the
    reference's bit for bit (which also pins the port at this size).  -> tests/golden/reference_lsm_1e6.json"""
    import time

    sys.path.insert(0, ROOT)
    from oracle import np_port

    ps = ref.BlackScholesPathSimulation()
    ps.set_seed(1)
    t0 = time.perf_counter()
    paths = ps.simulate_paths(n_paths, n_steps=n_steps)
    t_paths = time.perf_counter() - t0
    out = {"provenance": provenance, "n_paths": n_paths, "n_steps": n_steps, "seed": 1, "K": 100.0, "r": 0.05,
           "paths_checksum": float(paths.sum()), "paths_last": float(paths[-1, -1]), "paths_col_means": {
               str(t): float(paths[:, t].mean()) for t in (1, 63, 126, 252)}}
    for opt in ("put", "call"):
        t0 = time.perf_counter()
        price = _american_exercise_lsm(paths, 100.0, 0.05, 1.0 / n_steps, opt)
        t_lsm = time.perf_counter() - t0
        port_price, ex = np_port.lsm_price(paths, 100.0, 0.05, 1.0 / n_steps, opt)
        assert port_price == price, (opt, port_price, price)  # bit for bit on the same matrix
        out["lsm_" + opt] = {"price": float(price), "exercise_step_histogram": np.bincount(ex, minlength=n_steps + 1).tolist(),
                             "reference_seconds": t_lsm}
    out["reference_simulate_paths_seconds"] = t_paths
    with open(os.path.join(ROOT, "tests", "golden", "reference_lsm_1e6.json"), "w") as f:
        json.dump(jsonable(out), f, indent=1, sort_keys=True)
    print("wrote tests/golden/reference_lsm_1e6.json:", {k: v["price"] for k, v in out.items() if k.startswith("lsm_")})


def main():
    J, A = {}, {}
    J["provenance"] = {
        "reference": "milanfusco/mcFramework @ /root/reference (mcframework %s)" % ref.__version__,
        "numpy": np.__version__, "scipy": scipy.__version__, "python": sys.version.split()[0],
        "generator": "tools/gen_golden.py",
    }

    # ---- G0: seeding -------------------------------------------------------------------
    ss = np.random.SeedSequence(123)
    J["G0"] = {
        "seedseq123_state4": ss.generate_state(4, np.uint64).tolist(),
        "pcg64_123_first3": np.random.default_rng(123).bit_generator.random_raw(3).tolist(),
        "philox_child_keys": [c.generate_state(2, np.uint64).tolist() for c in np.random.SeedSequence(123).spawn(2)],
        "philox_child0_first4": np.random.Philox(np.random.SeedSequence(123).spawn(1)[0]).random_raw(4).tolist(),
    }

    # ---- G1: BASELINE config 1, as the README runs it ------------------------------------
    sim = ref.PiEstimationSimulation()
    sim.set_seed(123)
    fw = ref.MonteCarloFramework()
    fw.register_simulation(sim)
    res = fw.run_simulation("Pi Estimation", 10_000, n_points=5000, parallel=True,
                            extra_context={"rng": 7, "target": float(np.pi), "eps": 0.01})
    A["G1_pi_results"] = res.results
    J["G1"] = {"n": 10_000, "n_points": 5000, "seed": 123, "sum_hits": int(round(res.results.sum() * 5000 / 4)),
               "mean": res.mean, "std": res.std, "percentiles": res.percentiles, "stats": res.stats,
               "extra_context": {"rng": 7, "target": float(np.pi), "eps": 0.01}}

    # ---- G2: Philox block layout (n >= 20 000) -------------------------------------------
    sim = ref.PiEstimationSimulation()
    sim.set_seed(123)
    r2 = sim.run(20_000, n_points=100, parallel=True, n_workers=8, compute_stats=False)
    A["G2_pi_parallel_results"] = r2.results
    sim.set_seed(5)
    r2b = sim.run(20_001, n_points=33, antithetic=True, parallel=True, n_workers=3, compute_stats=False)
    A["G2b_pi_antithetic_parallel_results"] = r2b.results
    sim.set_seed(6)
    A["G2c_pi_antithetic_seq_results"] = sim.run(500, n_points=101, antithetic=True, compute_stats=False).results
    # second run WITHOUT re-seeding continues the PCG64 stream (core.py:593-594)
    A["G2d_pi_seq_second_run_results"] = sim.run(300, n_points=64, compute_stats=False).results
    J["G2"] = {"sum_hits": int(round(r2.results.sum() * 100 / 4))}

    # ---- G3: dice (user-defined integers sim) ----------------------------------------------
    d = Dice("2d6")
    d.set_seed(42)
    A["G3_dice_seq"] = d.run(10_000, compute_stats=False).results
    d.set_seed(42)
    A["G3_dice_par"] = d.run(20_000, parallel=True, n_workers=4, compute_stats=False).results

    # ---- G4: Black-Scholes -----------------------------------------------------------------
    bs = ref.BlackScholesSimulation()
    bs.set_seed(42)
    A["G4_bs_call_seq"] = bs.run(1000, compute_stats=False).results
    bs.set_seed(42)
    A["G4_bs_put_seq"] = bs.run(777, option_type="put", S0=95.0, K=105.0, T=0.5, r=0.03, sigma=0.35, n_steps=61,
                                compute_stats=False).results
    bs.set_seed(42)
    A["G4_bs_amer_put_seq"] = bs.run(1000, option_type="put", exercise_type="american", compute_stats=False).results
    bs.set_seed(42)
    A["G4_bs_amer_call_seq"] = bs.run(400, option_type="call", exercise_type="american", n_steps=40,
                                      compute_stats=False).results
    bs.set_seed(11)
    A["G4_bs_call_par"] = bs.run(20_000, parallel=True, n_workers=4, n_steps=64, compute_stats=False).results
    bs.set_seed(11)
    A["G4_bs_amer_put_par"] = bs.run(20_000, parallel=True, n_workers=2, n_steps=16, option_type="put",
                                     exercise_type="american", compute_stats=False).results
    bs.set_seed(9)
    J["G4_greeks_call"] = bs.calculate_greeks(1000)
    J["G4_greeks_put_par"] = bs.calculate_greeks(20_000, option_type="put", n_steps=12, parallel=True)
    # NOTE: parallel greeks use n_workers = cpu_count of THIS box
    J["G4_greeks_put_par_cpu_count"] = os.cpu_count()

    # ---- G5: Portfolio ---------------------------------------------------------------------
    pf = ref.PortfolioSimulation()
    pf.set_seed(42)
    A["G5_portfolio_gbm_seq"] = pf.run(500, years=5, compute_stats=False).results
    pf.set_seed(42)
    A["G5_portfolio_log1p_seq"] = pf.run(500, years=2, use_gbm=False, compute_stats=False).results
    pf.set_seed(3)
    A["G5_portfolio_gbm_par"] = pf.run(20_000, years=1, parallel=True, n_workers=8, compute_stats=False).results

    # ---- G6: paths + LSM -------------------------------------------------------------------
    ps = ref.BlackScholesPathSimulation()
    ps.set_seed(5)
    A["G6_path_terminal_seq"] = ps.run(300, n_steps=30, compute_stats=False).results
    ps.set_seed(1)
    paths = ps.simulate_paths(4000, n_steps=50)
    A["G6_paths_head"] = paths[:8]
    A["G6_paths_tail_row"] = paths[-1]
    J["G6"] = {
        "lsm_put": _american_exercise_lsm(paths, 100.0, 0.05, 1.0 / 50, "put"),
        "lsm_call": _american_exercise_lsm(paths, 100.0, 0.05, 1.0 / 50, "call"),
        "lsm_put_K110": _american_exercise_lsm(paths, 110.0, 0.05, 1.0 / 50, "put"),
        "paths_checksum": float(paths.sum()),
    }
    rng = np.random.default_rng(42)
    small = np.array([ref.sims._simulate_gbm_path(100.0, 0.05, 0.2, 1.0, 50, rng) for _ in range(100)])
    J["G6"]["lsm_small_put"] = _american_exercise_lsm(small, 100.0, 0.05, 1.0 / 50, "put")
    J["G6"]["lsm_small_call"] = _american_exercise_lsm(small, 100.0, 0.05, 1.0 / 50, "call")

    # ---- G7/G8: stats engine ---------------------------------------------------------------
    eng = build_default_engine()
    x = np.random.default_rng(0).normal(5, 2, 100_000)
    ctx = StatsContext(n=x.size, percentiles=(1, 5, 50, 95, 99), rng=7, n_bootstrap=1000, bootstrap="bca",
                       target=5.0, eps=0.01)
    J["G7"] = eng.compute(x, ctx).metrics
    xs = np.random.default_rng(3).exponential(2.0, 1000)
    xs[::97] = np.nan
    ctx = StatsContext(n=xs.size, nan_policy="omit", rng=11, n_bootstrap=500, bootstrap="percentile",
                       confidence=0.9, target=2.0, eps=0.1, ci_method="t")
    J["G7_omit"] = eng.compute(xs, ctx).metrics
    A["G7_omit_x"] = xs
    six = np.linspace(0, 1, 6)
    J["G8"] = {
        "percentile_rng123_B200": eng.compute(six, StatsContext(n=6, rng=123, n_bootstrap=200),
                                              select=["ci_mean_bootstrap"]).metrics,
        "bca_rng321_B200": eng.compute(six, StatsContext(n=6, rng=321, n_bootstrap=200, bootstrap="bca"),
                                       select=["ci_mean_bootstrap"]).metrics,
    }

    if "--lsm-large" in sys.argv:
        lsm_large(J["provenance"])
        return

    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    with open(os.path.join(ROOT, "tests", "golden", "reference_golden.json"), "w") as f:
        json.dump(jsonable(J), f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"), **A)
    print("wrote", len(J), "json groups and", len(A), "arrays")


if __name__ == "__main__":
    main()
