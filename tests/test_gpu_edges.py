"""Edge cases of the hot path through the public API: smallest sizes, ragged block layouts, odd step counts,
non-finite data, many percentiles, repeated / un-reseeded runs.  The expected values come from the oracle
(NumPy-level port of the reference, oracle/np_port.py) on the same seeds."""
import logging
import math

import numpy as np
import pytest

from conftest import have_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_cuda(), reason="needs a CUDA device")]
RTOL, ATOL = 1e-12, 1e-11


@pytest.fixture(scope="module")
def m():
    import mcframework_b200 as mod
    logging.getLogger("mcframework_b200.core").setLevel(logging.WARNING)
    return mod


def _port(sim, seed, n, parallel=False, n_workers=None, **kw):
    from oracle import np_port
    r = np_port.Runner(sim, seed=seed)
    return r.run(n, parallel=parallel, n_workers=n_workers, **kw)


@pytest.mark.parametrize("n,n_points,anti", [(1, 1, False), (1, 1, True), (3, 2, True), (7, 3, True), (2, 100_000, False), (40, 65, True)])
def test_pi_smallest_and_odd_sizes(m, n, n_points, anti):
    sim = m.PiEstimationSimulation()
    sim.set_seed(9)
    got = sim.run(n, n_points=n_points, antithetic=anti, compute_stats=False).results
    assert np.array_equal(got, _port("pi", 9, n, n_points=n_points, antithetic=anti))


@pytest.mark.parametrize("n,workers,steps", [(20_000, 7, 1), (20_001, 8, 3), (23_456, 5, 64), (20_000, 64, 65), (20_000, 3000, 5)])
def test_ragged_block_layouts_and_odd_step_counts(m, n, workers, steps):
    """Last block shorter than the others, more workers than make sense, 1-step and 65-step paths."""
    bs = m.BlackScholesSimulation()
    bs.set_seed(4)
    got = bs.run(n, parallel=True, n_workers=workers, n_steps=steps, option_type="put", compute_stats=False).results
    want = _port("black_scholes", 4, n, parallel=True, n_workers=workers, n_steps=steps, option_type="put")
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=ATOL)


def test_single_replication_and_direct_hook_calls(m):
    from oracle import np_port
    bs = m.BlackScholesSimulation()
    bs.set_seed(2)
    got = [bs.single_simulation(n_steps=17), bs.single_simulation(option_type="put", exercise_type="american", n_steps=9),
           float(bs.run(1, n_steps=5, compute_stats=False).results[0])]
    rng = np.random.default_rng(np.random.SeedSequence(2))
    want = [np_port.black_scholes_once(rng, n_steps=17), np_port.black_scholes_once(rng, option_type="put", exercise_type="american", n_steps=9),
            np_port.black_scholes_once(rng, n_steps=5)]
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=ATOL)
    pf = m.PortfolioSimulation()
    pf.set_seed(3)
    rng = np.random.default_rng(np.random.SeedSequence(3))
    np.testing.assert_allclose(pf.single_simulation(years=0.5), np_port.portfolio_once(rng, years=0.5), rtol=RTOL)
    np.testing.assert_allclose(pf.single_simulation(years=1, use_gbm=False), np_port.portfolio_once(rng, years=1, use_gbm=False), rtol=RTOL)
    # zero-step horizon (reference tests/test_errors_and_edge_cases.py::test_zero_year_portfolio): the initial value, no draw
    before = pf.rng.bit_generator.state
    assert pf.single_simulation(initial_value=10_000, annual_return=0.07, volatility=0.2, years=0) == 10_000.0
    r0 = pf.run(50, years=0, initial_value=123.5, percentiles=(50,))
    assert np.array_equal(r0.results, np.full(50, 123.5)) and r0.mean == 123.5 and r0.percentiles[50] == 123.5
    rp = pf.run(20_000, years=0, parallel=True, n_workers=2, compute_stats=False)
    assert np.array_equal(rp.results, np.full(20_000, 10_000.0))
    assert pf.rng.bit_generator.state == before
    # a Philox generator handed in through _rng (what a reference worker thread does)
    g1, g2 = (np.random.Generator(np.random.Philox(5)) for _ in range(2))
    pi = m.PiEstimationSimulation()
    assert pi.single_simulation(n_points=333, _rng=g1) == np_port.pi_once(g2, n_points=333)
    assert pi.single_simulation(n_points=10, antithetic=True, _rng=g1) == np_port.pi_once(g2, n_points=10, antithetic=True)
    assert g1.bit_generator.state["state"]["counter"].tolist() == g2.bit_generator.state["state"]["counter"].tolist()


def test_stats_engine_edge_inputs(m):
    from mcframework_b200 import stats_engine as se
    eng = se.build_default_engine()
    # n = 1 and n = 2
    one = eng.compute(np.array([3.5]), m.StatsContext(n=1, rng=1, n_bootstrap=100))
    assert one.metrics["mean"] == 3.5 and "std" not in one.metrics and one.metrics["skew"] == 0.0
    assert one.metrics["percentiles"] == {5: 3.5, 25: 3.5, 50: 3.5, 75: 3.5, 95: 3.5}
    assert math.isnan(one.metrics["ci_mean"]["low"]) and one.metrics["ci_mean_bootstrap"]["low"] == 3.5
    two = eng.compute(np.array([1.0, 2.0]), m.StatsContext(n=2, rng=1, n_bootstrap=100))
    assert two.metrics["std"] == pytest.approx(np.std([1.0, 2.0], ddof=1)) and two.metrics["ci_mean"]["method"] == "t"
    # constant data: scipy returns NaN for skew / kurtosis, std 0, degenerate intervals
    const = eng.compute(np.full(1000, 2.25), m.StatsContext(n=1000, rng=1, n_bootstrap=100))
    assert const.metrics["std"] == 0.0 and math.isnan(const.metrics["skew"]) and math.isnan(const.metrics["kurtosis"])
    assert const.metrics["ci_mean"]["low"] == const.metrics["ci_mean"]["high"] == 2.25
    assert const.metrics["ci_mean_bootstrap"]["low"] == pytest.approx(2.25, rel=1e-15)
    # non-finite values under both policies
    x = np.random.default_rng(1).normal(0, 1, 5000)
    x[[1, 50]] = [np.nan, np.inf]
    prop = eng.compute(x, m.StatsContext(n=x.size, rng=1, n_bootstrap=100)).metrics
    assert math.isnan(prop["mean"]) and math.isnan(prop["std"]) and all(math.isnan(v) for v in prop["percentiles"].values())
    # ... and with BCa the reference's np.percentile rejects the NaN-adjusted percentiles: the engine re-raises ValueError
    with pytest.raises(ValueError, match=r"Percentiles must be in the range \[0, 100\]"):
        eng.compute(x, m.StatsContext(n=x.size, rng=1, n_bootstrap=100, bootstrap="bca"))
    omit = eng.compute(x, m.StatsContext(n=x.size, rng=1, n_bootstrap=100, nan_policy="omit")).metrics
    f = x[np.isfinite(x)]
    assert omit["mean"] == pytest.approx(f.mean(), rel=1e-13) and omit["percentiles"][50] == np.percentile(f, 50)
    assert omit["ci_mean"]["se"] == pytest.approx(f.std(ddof=1) / math.sqrt(f.size), rel=1e-12)
    # 40 percentiles in one call (more than one selection launch), including 0 and 100
    pcts = tuple(range(0, 101, 5)) + (1, 2, 3, 4, 96, 97, 98, 99, 33, 34, 35, 36, 37, 38, 39, 41, 42, 43, 44)
    got = se.percentiles(f, m.StatsContext(n=f.size, percentiles=pcts))
    assert got == dict(zip(pcts, map(float, np.percentile(f, pcts))))
    assert se.percentiles(np.array([]), m.StatsContext(n=0, percentiles=(5,))) != se.percentiles(np.array([]), m.StatsContext(n=0, percentiles=(5,)))  # NaN != NaN


def test_unreseeded_runs_continue_streams(m):
    """Sequential runs continue self.rng; parallel runs spawn NEW children from the same SeedSequence (the reference's
    n_children_spawned side effect)."""
    from oracle import np_port
    bs, port = m.BlackScholesSimulation(), np_port.Runner("black_scholes", seed=8)
    bs.set_seed(8)
    for n, par in ((300, False), (20_000, True), (150, False), (20_000, True)):
        got = bs.run(n, parallel=par, n_workers=4, n_steps=11, compute_stats=False).results
        np.testing.assert_allclose(got, port.run(n, parallel=par, n_workers=4, n_steps=11), rtol=RTOL, atol=ATOL)


def test_percentile_merging_and_metadata(m):
    """User percentiles merged over the engine defaults (core.py:452-484 of the reference), with and without stats."""
    sim = m.PiEstimationSimulation()
    sim.set_seed(3)
    r = sim.run(2000, n_points=200, percentiles=[10, 90, 50])
    assert set(r.percentiles) == {5, 25, 50, 75, 95, 10, 90}
    for p, v in r.percentiles.items():
        assert v == np.percentile(r.results, p)
    assert r.metadata["requested_percentiles"] == [10, 90, 50] and r.metadata["engine_defaults_used"] is True
    sim.set_seed(3)
    r2 = sim.run(2000, n_points=200, percentiles=[10, 90], compute_stats=False)
    assert set(r2.percentiles) == {10, 90} and r2.stats == {} and r2.metadata["engine_defaults_used"] is False
    assert np.array_equal(r.results, r2.results)
    sim.set_seed(3)
    r3 = sim.run(2000, n_points=200, compute_stats=False)
    assert r3.percentiles == {} and r3.mean == pytest.approx(r.results.mean(), rel=1e-14)
    fw = m.MonteCarloFramework()
    fw.register_simulation(sim, "pi")
    fw.results["pi"] = r
    assert fw.compare_results(["pi"], "p90") == {"pi": r.percentiles[90]}
    with pytest.raises(ValueError, match="Percentile 25 not computed"):
        fw.compare_results(["pi"], "p25")  # engine default, but not among the REQUESTED ones
    text = r.result_to_string()
    assert "Results for simulation 'Pi Estimation':" in text and "Number of simulations: 2000" in text
