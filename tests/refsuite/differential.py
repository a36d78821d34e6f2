"""TEST INFRASTRUCTURE ONLY -- differential run: the UNMODIFIED reference (imported from ``/root/reference/src``) and
``mcframework_b200`` on the oracle-backed stand-in device (``fakedev``) are driven through the same public calls with
the same seeds; every output is compared.  Run as a script in its own process (``fakedev.install()`` patches globally):

    python tests/refsuite/differential.py [n_cases [seed]]

Exit code 0 and a last line ``differential: <k> comparisons, 0 mismatches`` on success.
"""
from __future__ import annotations

import math
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), "/root/reference/src"):
    if p not in sys.path:
        sys.path.insert(0, p)

import logging  # noqa: E402

import mcframework as ref  # noqa: E402  (the reference itself)
from refsuite import fakedev  # noqa: E402

fakedev.install()
import mcframework_b200 as new  # noqa: E402

for name in ("mcframework.core", "mcframework_b200.core", "mcframework.stats_engine", "mcframework_b200.stats_engine"):
    logging.getLogger(name).setLevel(logging.CRITICAL)

RTOL = 1e-9  # host-side assembly; the stand-in sums in NumPy order, the reference in its own
mismatches: list[str] = []
n_cmp = 0


def same(a, b, where: str, rtol: float = RTOL) -> None:
    """Deep comparison: structure, keys, types of leaves (str/None/bool exactly, numbers within rtol, NaN == NaN)."""
    global n_cmp
    if isinstance(a, dict) or isinstance(b, dict):
        if not (isinstance(a, dict) and isinstance(b, dict)) or set(a) != set(b):
            mismatches.append(f"{where}: keys {sorted(map(str, a)) if isinstance(a, dict) else a!r} vs "
                              f"{sorted(map(str, b)) if isinstance(b, dict) else b!r}")
            return
        for k in a:
            same(a[k], b[k], f"{where}[{k!r}]", rtol)
        return
    if isinstance(a, (list, tuple)) or isinstance(b, (list, tuple)):
        if not (isinstance(a, (list, tuple)) and isinstance(b, (list, tuple))) or len(a) != len(b):
            mismatches.append(f"{where}: {a!r} vs {b!r}")
            return
        for i, (x, y) in enumerate(zip(a, b)):
            same(x, y, f"{where}[{i}]", rtol)
        return
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
        n_cmp += 1
        if a.shape != b.shape or not np.allclose(a, b, rtol=rtol, atol=1e-12, equal_nan=True):
            mismatches.append(f"{where}: arrays differ (shapes {a.shape} {b.shape})")
        return
    n_cmp += 1
    if a is None or b is None or isinstance(a, (str, bool)) or isinstance(b, (str, bool)):
        if a != b:
            mismatches.append(f"{where}: {a!r} vs {b!r}")
        return
    try:
        fa, fb = float(a), float(b)
    except (TypeError, ValueError):
        if a != b:
            mismatches.append(f"{where}: {a!r} vs {b!r}")
        return
    if math.isnan(fa) and math.isnan(fb):
        return
    if fa == fb or abs(fa - fb) <= rtol * max(abs(fa), abs(fb)) + 1e-12:
        return
    mismatches.append(f"{where}: {fa!r} vs {fb!r}")


def outcome(fn):
    """(kind, payload): the value, or the exception type name and message."""
    try:
        return "ok", fn()
    except Exception as exc:  # noqa: BLE001
        return "raise", (type(exc).__name__, str(exc))


def both(where, f_ref, f_new, rtol: float = RTOL):
    a, b = outcome(f_ref), outcome(f_new)
    if a[0] != b[0]:
        mismatches.append(f"{where}: reference {a!r} vs drop-in {b!r}")
        return None, None
    if a[0] == "raise":
        # same exception family (MissingContextError subclasses ValueError in both) and same message
        same(a[1][1], b[1][1], where + " (message)")
        if a[1][0] != b[1][0]:
            mismatches.append(f"{where}: exception {a[1][0]} vs {b[1][0]}")
        return None, None
    same(a[1], b[1], where, rtol)
    return a[1], b[1]


def result_fields(r, seeded_bootstrap: bool):
    meta = {k: v for k, v in r.metadata.items() if k != "timestamp"}
    stats = dict(r.stats)
    if not seeded_bootstrap and isinstance(stats.get("ci_mean_bootstrap"), dict):  # fresh OS entropy on both sides
        stats["ci_mean_bootstrap"] = {k: v for k, v in stats["ci_mean_bootstrap"].items() if k not in ("low", "high")}
    return {"results": np.asarray(r.results), "n": r.n_simulations, "mean": r.mean, "std": r.std,
            "percentiles": r.percentiles, "stats": stats, "metadata": meta}


# ------------------------------------------------------------------------------------------------ scenarios
def engine_cases(rng: np.random.Generator, n_cases: int):
    for case in range(n_cases):
        n = int(rng.choice([0, 1, 2, 3, 4, 7, 29, 30, 31, 200]))
        x = rng.normal(3.0, 2.0, n) if rng.random() < 0.7 else np.round(rng.normal(0, 3, n))  # ties
        taint = rng.random()
        if n >= 3 and taint < 0.25:
            x[rng.integers(0, n)] = np.nan
        elif n >= 3 and taint < 0.35:
            x[rng.integers(0, n)] = np.inf
        kw = dict(n=n, confidence=float(rng.choice([0.5, 0.9, 0.95, 0.99])),
                  ci_method=str(rng.choice(["auto", "z", "t", "bootstrap"])),
                  nan_policy=str(rng.choice(["propagate", "omit"])),
                  bootstrap=str(rng.choice(["percentile", "bca"])), n_bootstrap=int(rng.choice([100, 128])),
                  percentiles=tuple(int(p) for p in rng.choice([1, 5, 25, 50, 75, 95, 99], size=int(rng.integers(0, 4)),
                                                                replace=False)))
        if rng.random() < 0.6:
            kw["target"] = float(rng.normal(3.0, 1.0))
        if rng.random() < 0.6:
            kw["eps"] = float(rng.choice([0.01, 0.5]))
        if rng.random() < 0.3:
            kw["ess"] = int(rng.integers(1, max(2, n + 1)))
        if rng.random() < 0.3:
            kw["ddof"] = int(rng.choice([0, 1]))
        seed = int(rng.integers(0, 2 ** 31))
        where = f"engine[{case}] n={n} {kw}"

        def run(mod, as_dict):
            ctx_kw = dict(kw, rng=seed)
            ctx = ctx_kw if as_dict else mod.StatsContext(**ctx_kw)
            res = mod.stats_engine.build_default_engine().compute(x.copy(), ctx)
            return {"metrics": res.metrics, "skipped": sorted(n_ for n_, _ in res.skipped),
                    "errors": sorted(n_ for n_, _ in res.errors)}

        as_dict = bool(rng.random() < 0.3)
        both(where, lambda: run(ref, as_dict), lambda: run(new, as_dict))
        sel = [str(s) for s in rng.choice(["mean", "std", "percentiles", "skew", "kurtosis", "ci_mean", "ci_mean_chebyshev",
                                           "markov_error_prob", "bias_to_target", "mse_to_target", "chebyshev_required_n"],
                                          size=3, replace=False)]

        def run_sel(mod):
            res = mod.stats_engine.build_default_engine().compute(x.copy(), mod.StatsContext(**dict(kw, rng=seed)), select=sel)
            return res.metrics

        both(where + f" select={sel}", lambda: run_sel(ref), lambda: run_sel(new))


def run_cases(rng: np.random.Generator, n_cases: int):
    def sims(mod):
        return {"pi": mod.PiEstimationSimulation, "portfolio": mod.PortfolioSimulation, "bs": mod.BlackScholesSimulation,
                "path": mod.BlackScholesPathSimulation}

    for case in range(n_cases):
        which = str(rng.choice(["pi", "portfolio", "bs", "path"]))
        seed = int(rng.integers(0, 10_000))
        parallel = bool(rng.random() < 0.3)
        n = int(rng.choice([20_000, 20_011])) if parallel else int(rng.choice([1, 2, 17, 300]))
        sim_kw = {"pi": dict(n_points=int(rng.choice([1, 2, 7, 50])), antithetic=bool(rng.random() < 0.5)),
                  "portfolio": dict(years=float(rng.choice([0, 0.01, 0.1, 1])), use_gbm=bool(rng.random() < 0.5),
                                    volatility=float(rng.choice([0.0, 0.2]))),
                  "bs": dict(option_type=str(rng.choice(["call", "put"])), exercise_type=str(rng.choice(["european", "american"])),
                             n_steps=int(rng.choice([1, 5, 16, 33])), K=float(rng.choice([90.0, 100.0, 130.0]))),
                  "path": dict(n_steps=int(rng.choice([1, 12])), sigma=float(rng.choice([0.1, 0.4])))}[which]
        if parallel and which == "pi":
            sim_kw["n_points"] = min(sim_kw["n_points"], 7)
        run_kw = dict(parallel=parallel, compute_stats=bool(rng.random() < 0.8), confidence=float(rng.choice([0.9, 0.95])),
                      ci_method=str(rng.choice(["auto", "z", "t", "bootstrap"])))
        if parallel:
            run_kw["n_workers"] = int(rng.choice([2, 3, 8]))
        if rng.random() < 0.6:
            run_kw["percentiles"] = [int(p) for p in rng.choice([1, 5, 10, 50, 90, 99], size=int(rng.integers(0, 4)), replace=False)]
            if run_kw["percentiles"] and rng.random() < 0.3:
                run_kw["percentiles"] = run_kw["percentiles"] + run_kw["percentiles"][:1]  # a duplicate, unsorted
        if rng.random() < 0.5:
            run_kw["extra_context"] = dict(target=float(rng.normal(3, 1)), eps=0.05, n_bootstrap=100, rng=int(seed),
                                           bootstrap=str(rng.choice(["percentile", "bca"])))
        where = f"run[{case}] {which} seed={seed} n={n} {sim_kw} {run_kw}"

        def go(mod):
            sim = sims(mod)[which]()
            sim.set_seed(seed)
            fw = mod.MonteCarloFramework()
            fw.register_simulation(sim)
            ticks = []
            first = fw.run_simulation(sim.name, n, progress_callback=lambda d, t: ticks.append((d, t)), **run_kw, **sim_kw)
            second = sim.run(3, compute_stats=False, **sim_kw)  # un-reseeded continuation of the stream
            cmp_ = {}
            for metric in ("mean", "std", "var", "se", "p50", "p7"):
                cmp_[metric] = outcome(lambda: fw.compare_results([sim.name], metric))
            # the report prints floats at full precision: keep its structure, round the numbers to 9 significant digits
            text = [re.sub(r"-?\d+\.\d+(?:e[-+]?\d+)?", lambda m_: f"{float(m_.group()):.9g}", ln)
                    for ln in first.result_to_string().splitlines()
                    if "Execution time" not in ln and "timestamp" not in ln and "ci_mean_bootstrap" not in ln]
            return {"first": result_fields(first, "extra_context" in run_kw), "second": np.asarray(second.results),
                    "compare": cmp_, "text": text, "spawned": sim.seed_seq.n_children_spawned,
                    # thread completion order decides the intermediate counts of a parallel run (ragged last block)
                    "ticks": ticks if not parallel else [len(ticks), max(ticks), sorted({t for _, t in ticks})]}

        import contextlib
        import io

        def quiet(mod):
            with contextlib.redirect_stdout(io.StringIO()):
                return go(mod)

        both(where, lambda: quiet(ref), lambda: quiet(new), rtol=1e-9)


def hook_cases(rng: np.random.Generator):
    """User plugin code (arbitrary Python hooks): looped on the host by both, on the same streams."""
    def make(mod):
        class Dice(mod.MonteCarloSimulation):
            def single_simulation(self, _rng=None, n_dice=2, **kw):
                g = self._rng(_rng, self.rng)
                return float(g.integers(1, 7, size=n_dice).sum())

        return Dice("Dice")

    for n, par, workers, backend in ((50, False, None, "auto"), (20_000, True, 2, "thread"), (20_003, True, 3, "auto"),
                                     (19_999, True, 4, "thread"), (20_000, True, 1, "thread")):
        seed = int(rng.integers(0, 1000))

        def go(mod):
            sim = make(mod)
            sim.parallel_backend = backend
            sim.set_seed(seed)
            kw = dict(parallel=par, percentiles=(10, 90), n_dice=3)
            if workers:
                kw["n_workers"] = workers
            r = sim.run(n, **kw)
            r2 = sim.run(7, compute_stats=False)
            return {"a": result_fields(r, False), "b": np.asarray(r2.results)}

        both(f"hook n={n} parallel={par} workers={workers} backend={backend}", lambda: go(ref), lambda: go(new))


def api_surface():
    """Every public name of the reference (module ``__all__`` lists) exists in the drop-in with the same call signature:
    functions, class constructors and public methods (parameter names, kinds and defaults; extra keyword-only
    parameters with defaults are allowed on the drop-in side, they cannot break a caller of the reference)."""
    import dataclasses
    import inspect
    from importlib import import_module

    def params(fn):
        try:
            sig = inspect.signature(fn)
        except (TypeError, ValueError):
            return None
        return [(p.name, str(p.kind), None if p.default is inspect._empty else repr(p.default)) for p in sig.parameters.values()]

    def compatible(where, pr, pn):
        global n_cmp
        n_cmp += 1
        if pr is None or pn is None:
            return
        extra = [q for q in pn[len(pr):]] if pn[:len(pr)] == pr else None
        if pn == pr:
            return
        names_r = [q[0] for q in pr]
        rest = [q for q in pn if q[0] not in names_r]
        kept = [q for q in pn if q[0] in names_r]
        ok = kept == pr and all(q[1] in ("KEYWORD_ONLY", "VAR_KEYWORD") or q[2] is not None for q in rest) and \
            all(q[1] == "KEYWORD_ONLY" or q[1] == "VAR_KEYWORD" for q in rest)
        if not ok:
            mismatches.append(f"{where}: signature {pr} vs {pn}")

    for modname in ("", ".core", ".stats_engine", ".utils", ".sims", ".sims.pi", ".sims.portfolio", ".sims.black_scholes"):
        mr, mn = import_module("mcframework" + modname), import_module("mcframework_b200" + modname)
        names = list(getattr(mr, "__all__", []))
        same(sorted(n_ for n_ in names if not hasattr(mn, n_)), [], f"api{modname}: names missing in the drop-in")
        for name in names:
            a, b = getattr(mr, name), getattr(mn, name, None)
            if b is None:
                continue
            if inspect.isclass(a):
                if dataclasses.is_dataclass(a):
                    same([(f.name, repr(f.default) if f.default is not dataclasses.MISSING else None) for f in dataclasses.fields(a)],
                         [(f.name, repr(f.default) if f.default is not dataclasses.MISSING else None) for f in dataclasses.fields(b)],
                         f"api{modname}.{name} fields")
                elif issubclass(a, BaseException) or hasattr(a, "__members__"):
                    same([c.__name__ for c in a.__mro__[1:3]], [c.__name__ for c in b.__mro__[1:3]], f"api{modname}.{name} bases")
                    if hasattr(a, "__members__"):
                        same({k: v.value for k, v in a.__members__.items()}, {k: v.value for k, v in b.__members__.items()},
                             f"api{modname}.{name} members")
                else:
                    compatible(f"api{modname}.{name}.__init__", params(a.__init__), params(b.__init__))
                for meth, fn in inspect.getmembers(a, predicate=inspect.isfunction):
                    if meth.startswith("_") and meth not in ("__call__",):
                        continue
                    other = getattr(b, meth, None)
                    if other is None:
                        mismatches.append(f"api{modname}.{name}.{meth}: missing in the drop-in")
                        continue
                    compatible(f"api{modname}.{name}.{meth}", params(fn), params(other))
                for attr in ("_PCTS", "_PARALLEL_THRESHOLD", "_CHUNKS_PER_WORKER"):
                    if hasattr(a, attr):
                        same(getattr(a, attr), getattr(b, attr, None), f"api{modname}.{name}.{attr}")
            elif callable(a):
                compatible(f"api{modname}.{name}", params(a), params(b))


def declared_cases(rng: np.random.Generator):
    """Declared draw programs (a drop-in extra): the reference runs the Python hook, the drop-in the declaration of the
    same hook through the device entry points -- results, statistics and the state the generator is left in must agree."""
    from mcframework_b200 import programs

    hooks = {
        "dice": (lambda self, g, kw: float(g.integers(1, 7, size=kw.get("n_dice", 2)).sum()),
                 lambda kw: programs.integers_sum(1, 7, size=kw.get("n_dice", 2))),
        "streak": (lambda self, g, kw: float(max((len(r) for r in "".join(map(str, g.integers(0, 2, size=31))).split("0")), default=0)),
                   lambda kw: programs.integers_longest_run(0, 2, size=31, value=1)),
        "normal": (lambda self, g, kw: float(g.normal(kw.get("loc", 1.5), 2.0)),
                   lambda kw: programs.normal(kw.get("loc", 1.5), 2.0)),
        "wide": (lambda self, g, kw: float(g.integers(0, 3_000_000_000, size=3).sum()),  # rejections: host loop on both sides
                 lambda kw: programs.integers_sum(0, 3_000_000_000, size=3)),
    }

    def make(mod, which, declare):
        hook, decl = hooks[which]

        class Sim(mod.MonteCarloSimulation):
            def single_simulation(self, _rng=None, **kw):
                return hook(self, self._rng(_rng, self.rng), kw)

        if declare:
            Sim.draw_program = lambda self, **kw: decl(kw)
        return Sim(which)

    for which in hooks:
        for n, par, workers in ((1, False, None), (37, False, None), (20_000, True, 2), (20_005, True, 5)):
            seed = int(rng.integers(0, 1000))
            kw = {"dice": dict(n_dice=3), "normal": dict(loc=-0.5)}.get(which, {})

            def go(mod, declare):
                sim = make(mod, which, declare)
                sim.set_seed(seed)
                run_kw = dict(parallel=par, percentiles=(25,), **kw)
                if workers:
                    run_kw["n_workers"] = workers
                first = sim.run(n, **run_kw)
                second = sim.run(5, compute_stats=False, **kw)  # continues the stream (odd slot counts included)
                return {"first": result_fields(first, False), "second": np.asarray(second.results),
                        "next_raw": int(sim.rng.bit_generator.random_raw())}

            both(f"declared {which} n={n} parallel={par} workers={workers}", lambda: go(ref, False), lambda: go(new, True))


def misc_cases():
    for conf, n, method in [(0.95, 10, "auto"), (0.95, 30, "auto"), (0.9, 5, "z"), (0.99, 100, "t"), (0.95, 10, "bootstrap"),
                            (0.95, 10, "nope")]:
        both(f"autocrit{(conf, n, method)}", lambda: ref.utils.autocrit(conf, n, method), lambda: new.utils.autocrit(conf, n, method))
    both("z_crit", lambda: ref.utils.z_crit(0.9), lambda: new.utils.z_crit(0.9))
    both("t_crit", lambda: ref.utils.t_crit(0.9, 7), lambda: new.utils.t_crit(0.9, 7))
    for bad in (dict(n=5, confidence=1.5), dict(n=5, percentiles=(101,)), dict(n=5, n_bootstrap=0), dict(n=5, ddof=-1),
                dict(n=5, eps=-1.0), dict(n=5, ess=0), dict(n=5, nan_policy="zap")):
        both(f"StatsContext{bad}", lambda: vars(ref.StatsContext(**bad)) and None, lambda: vars(new.StatsContext(**bad)) and None)
    for call in (dict(n_simulations=0), dict(n_simulations=5, n_workers=0), dict(n_simulations=5, confidence=1.0),
                 dict(n_simulations=5, ci_method="x")):
        both(f"run{call}", lambda: ref.PiEstimationSimulation().run(**call) and None,
             lambda: new.PiEstimationSimulation().run(**call) and None)
    greeks_kw = dict(n_simulations=200, n_steps=8, option_type="put")
    for par in (False,):
        def g(mod):
            s = mod.BlackScholesSimulation()
            s.set_seed(5)
            before = s.rng.bit_generator.state["state"]["state"]
            out = s.calculate_greeks(parallel=par, **greeks_kw)
            return {"greeks": out, "rng_restored": s.rng.bit_generator.state["state"]["state"] == before}
        both(f"greeks parallel={par}", lambda: g(ref), lambda: g(new), rtol=1e-7)

    # input types the engine accepts: lists, tuples, integer / bool / float32 arrays, empty input.  (Known superset, not
    # compared: a 2-D array -- the reference records errors for skew, kurtosis and the bootstrap interval because SciPy
    # and the index gather work along axis 0 there; the drop-in treats the sample as flat and returns all twelve.)
    for label, data in (("list", [1.0, 2.5, 4.0, 8.0]), ("ints", np.arange(12)),
                        ("tuple", (3.0, 1.0, 2.0)), ("bools", np.array([True, False, True, True])),
                        ("float32", np.linspace(0, 1, 9, dtype=np.float32)), ("empty", [])):
        def shaped(mod, data=data):
            n = int(np.asarray(data).size)
            res = mod.stats_engine.build_default_engine().compute(
                np.asarray(data) if label != "list" and label != "tuple" else data,
                mod.StatsContext(n=n, rng=5, n_bootstrap=100, target=1.0, eps=0.5, percentiles=(10, 50)))
            return {"metrics": res.metrics, "skipped": sorted(k for k, _ in res.skipped), "errors": sorted(k for k, _ in res.errors)}
        both(f"engine input {label}", lambda: shaped(ref), lambda: shaped(new))
    for conf in (0.001, 0.999, 0.5):
        def extreme(mod, conf=conf):
            x = np.random.default_rng(3).standard_t(3, 400)
            return mod.stats_engine.build_default_engine().compute(
                x, mod.StatsContext(n=400, confidence=conf, rng=9, n_bootstrap=200, bootstrap="bca", target=0.1, eps=0.2)).metrics
        both(f"engine confidence {conf}", lambda: extreme(ref), lambda: extreme(new))

    def pickled(mod):
        import pickle

        sim = mod.PiEstimationSimulation()
        sim.set_seed(77)
        sim.run(5, n_points=10, compute_stats=False)  # moves self.rng; the copy restarts from seed_seq (:318-330)
        copy = pickle.loads(pickle.dumps(sim))
        return {"copy": np.asarray(copy.run(4, n_points=10, compute_stats=False).results),
                "orig": np.asarray(sim.run(4, n_points=10, compute_stats=False).results), "name": copy.name,
                "entropy": copy.seed_seq.entropy}
    both("pickle round trip", lambda: pickled(ref), lambda: pickled(new))

    def unseeded(mod):
        sim = mod.PiEstimationSimulation()
        r = sim.run(3, n_points=5, compute_stats=False)
        return {"seed_entropy": r.metadata.get("seed_entropy"), "seq": sim.seed_seq is None, "n": len(r.results)}
    both("unseeded run", lambda: unseeded(ref), lambda: unseeded(new))

    def paths(mod):
        s = mod.BlackScholesPathSimulation()
        s.set_seed(9)
        p = s.simulate_paths(40, n_steps=20)
        from importlib import import_module
        bsm = import_module(mod.__name__ + ".sims.black_scholes")
        return {"paths": p, "put": bsm._american_exercise_lsm(p, 100.0, 0.05, 1.0 / 20, "put"),
                "call": bsm._american_exercise_lsm(p, 100.0, 0.05, 1.0 / 20, "call"),
                "one": bsm._simulate_gbm_path(100.0, 0.05, 0.2, 1.0, 6, np.random.default_rng(4))}
    both("paths+lsm", lambda: paths(ref), lambda: paths(new), rtol=1e-9)


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 20240601)
    api_surface()
    misc_cases()
    hook_cases(rng)
    declared_cases(rng)
    engine_cases(rng, n_cases)
    run_cases(rng, max(8, n_cases // 3))
    for m in mismatches[:40]:
        print("MISMATCH", m)
    print(f"differential: {n_cmp} comparisons, {len(mismatches)} mismatches")
    return 1 if mismatches else 0


if __name__ == "__main__":
    sys.exit(main())
