"""Public-API parity on the GPU: the scenarios of tools/gen_golden.py (run there on the UNMODIFIED reference)
replayed through ``mcframework_b200`` -- same calls, same seeds -- and compared with the committed fixtures.

Bars: Pi results (multiples of 4/n_points, i.e. integer hit counts) bit-exact; every float within RES_RTOL of
the reference result vector; statistics within STAT_RTOL (one-pass device sums vs NumPy pairwise sums);
bootstrap intervals within BOOT_RTOL (identical index stream, different summation order).
"""
import logging
import math

import numpy as np
import pytest

from conftest import have_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_cuda(), reason="needs a CUDA device")]

RES_RTOL, PAY_ATOL = 1e-12, 1e-11
STAT_RTOL = 1e-10
BOOT_RTOL = 1e-11


@pytest.fixture(scope="module")
def m():
    import mcframework_b200 as mod
    logging.getLogger("mcframework_b200.core").setLevel(logging.WARNING)
    return mod


def _close_stats(got, want, rtol=STAT_RTOL):
    assert set(got) == set(want), (sorted(got), sorted(want))
    for k, w in want.items():
        g = got[k]
        if isinstance(w, dict):
            assert set(g) == set(w), k
            for kk, ww in w.items():
                if isinstance(ww, str):
                    assert g[kk] == ww, (k, kk)
                else:
                    np.testing.assert_allclose(g[kk], ww, rtol=BOOT_RTOL if "bootstrap" in k else rtol, err_msg=f"{k}.{kk}")
        elif isinstance(w, (list, tuple)):
            np.testing.assert_allclose(g, w, rtol=rtol, err_msg=k)
        elif isinstance(w, str):
            assert g == w, k
        else:
            np.testing.assert_allclose(g, w, rtol=rtol, atol=1e-13, err_msg=k)


def test_config1_pi_run_simulation_end_to_end(m, golden):
    """BASELINE config 1 exactly as the README runs it (parallel=True falls back to the sequential stream)."""
    j, a = golden
    sim = m.PiEstimationSimulation()
    sim.set_seed(123)
    fw = m.MonteCarloFramework()
    fw.register_simulation(sim)
    seen = []
    res = fw.run_simulation("Pi Estimation", 10_000, n_points=5000, parallel=True,
                            extra_context=j["G1"]["extra_context"], progress_callback=lambda d, t: seen.append((d, t)))
    assert isinstance(res.results, np.ndarray) and res.results.dtype == np.float64
    assert np.array_equal(res.results, a["G1_pi_results"])  # hit counts bit-exact
    assert int(round(res.results.sum() * 5000 / 4)) == j["G1"]["sum_hits"] == 39_268_872
    np.testing.assert_allclose(res.mean, j["G1"]["mean"], rtol=1e-14)
    np.testing.assert_allclose(res.std, j["G1"]["std"], rtol=1e-12)
    assert res.percentiles == {int(k): v for k, v in j["G1"]["percentiles"].items()}  # order statistics bit-exact
    _close_stats(res.stats, j["G1"]["stats"])
    assert seen[-1] == (10_000, 10_000) and len(seen) == 100
    assert res.metadata["seed_entropy"] == 123 and res.metadata["engine_defaults_used"] is True
    assert fw.compare_results(["Pi Estimation"], "mean")["Pi Estimation"] == res.mean


def test_pi_layouts_and_stream_continuation(m, golden):
    _, a = golden
    sim = m.PiEstimationSimulation()
    sim.set_seed(123)
    assert np.array_equal(sim.run(20_000, n_points=100, parallel=True, n_workers=8, compute_stats=False).results,
                          a["G2_pi_parallel_results"])
    sim.set_seed(5)
    assert np.array_equal(sim.run(20_001, n_points=33, antithetic=True, parallel=True, n_workers=3, compute_stats=False).results,
                          a["G2b_pi_antithetic_parallel_results"])
    sim.set_seed(6)
    assert np.array_equal(sim.run(500, n_points=101, antithetic=True, compute_stats=False).results,
                          a["G2c_pi_antithetic_seq_results"])
    # a second run WITHOUT re-seeding continues the PCG64 stream where the first one stopped
    assert np.array_equal(sim.run(300, n_points=64, compute_stats=False).results, a["G2d_pi_seq_second_run_results"])


def test_user_hook_runs_as_plugin_code(m, golden):
    """Arbitrary Python hooks are user code: looped on the host with the reference's stream layout (G3)."""
    _, a = golden

    class Dice(m.MonteCarloSimulation):
        def single_simulation(self, _rng=None, **kw):
            rng = self._rng(_rng, self.rng)
            return float(rng.integers(1, 7, size=2).sum())

    d = Dice("2d6")
    d.set_seed(42)
    assert np.array_equal(d.run(10_000, compute_stats=False).results, a["G3_dice_seq"])
    d.set_seed(42)
    r = d.run(20_000, parallel=True, n_workers=4)
    assert np.array_equal(r.results, a["G3_dice_par"])
    np.testing.assert_allclose(r.mean, a["G3_dice_par"].mean(), rtol=1e-13)  # statistics still come from the kernels
    assert r.percentiles[50] == np.percentile(a["G3_dice_par"], 50)


def test_black_scholes_runs(m, golden):
    _, a = golden
    bs = m.BlackScholesSimulation()
    bs.set_seed(42)
    np.testing.assert_allclose(bs.run(1000, compute_stats=False).results, a["G4_bs_call_seq"], rtol=RES_RTOL, atol=PAY_ATOL)
    bs.set_seed(42)
    r = bs.run(777, option_type="put", S0=95.0, K=105.0, T=0.5, r=0.03, sigma=0.35, n_steps=61, compute_stats=False)
    np.testing.assert_allclose(r.results, a["G4_bs_put_seq"], rtol=RES_RTOL, atol=PAY_ATOL)
    bs.set_seed(42)
    r = bs.run(1000, option_type="put", exercise_type="american", compute_stats=False)
    np.testing.assert_allclose(r.results, a["G4_bs_amer_put_seq"], rtol=RES_RTOL, atol=PAY_ATOL)
    bs.set_seed(11)
    r = bs.run(20_000, parallel=True, n_workers=4, n_steps=64, compute_stats=False)
    np.testing.assert_allclose(r.results, a["G4_bs_call_par"], rtol=RES_RTOL, atol=PAY_ATOL)
    with pytest.raises(ValueError, match="exercise_type must be 'european' or 'american'"):
        bs.run(10, exercise_type="bermudan")


def test_pipelined_block_runs_equal_plain_runs(m, golden):
    """Large block-parallel runs copy their result to the host group by group behind the next group's kernels
    (core._PIPELINE_MIN_SIMS).  Forced on at fixture size: same vectors bit for bit, same statistics, same spawn
    side effects, with 1..5 groups and more groups than streams; Pi and Black-Scholes programs."""
    _, a = golden
    bs = m.BlackScholesSimulation()
    bs.set_seed(11)
    plain = bs.run(20_000, parallel=True, n_workers=4, n_steps=64, percentiles=(10, 90))
    np.testing.assert_allclose(plain.results, a["G4_bs_call_par"], rtol=RES_RTOL, atol=PAY_ATOL)
    for groups in (1, 2, 3, 5, 64):
        piped = m.BlackScholesSimulation()
        piped._PIPELINE_MIN_SIMS, piped._PIPELINE_GROUPS = 1, groups
        piped.set_seed(11)
        r = piped.run(20_000, parallel=True, n_workers=4, n_steps=64, percentiles=(10, 90))
        assert isinstance(r.results, np.ndarray) and r.results.dtype == np.float64 and r.results.flags.writeable
        assert np.array_equal(r.results, plain.results), groups
        assert (r.mean, r.std, r.percentiles) == (plain.mean, plain.std, plain.percentiles)
        assert piped.seed_seq.n_children_spawned == bs.seed_seq.n_children_spawned
    pi = m.PiEstimationSimulation()
    pi._PIPELINE_MIN_SIMS, pi._PIPELINE_GROUPS = 1, 3
    pi.set_seed(123)
    assert np.array_equal(pi.run(20_000, n_points=100, parallel=True, n_workers=8, compute_stats=False).results,
                          a["G2_pi_parallel_results"])
    pi.set_seed(5)  # ragged last block (20_001 = 3*8 blocks of 833 + 9)
    assert np.array_equal(pi.run(20_001, n_points=33, antithetic=True, parallel=True, n_workers=3, compute_stats=False).results,
                          a["G2b_pi_antithetic_parallel_results"])


def test_results_can_stay_on_the_device_until_somebody_reads_them(m, golden):
    """opt-in device-resident results.  Statistics, percentiles and compare_results need no host copy;
    the ndarray appears (once) when the caller touches it, and equals the one an ordinary run returns."""
    from mcframework_b200.core import DeviceResults

    _, a = golden
    bs = m.BlackScholesSimulation()
    fw = m.MonteCarloFramework()
    fw.register_simulation(bs)
    bs.set_seed(11)
    plain = fw.run_simulation(bs.name, 20_000, parallel=True, n_workers=4, n_steps=64, percentiles=(10, 90))
    bs.keep_results_on_device = True
    bs.set_seed(11)
    lazy = fw.run_simulation(bs.name, 20_000, parallel=True, n_workers=4, n_steps=64, percentiles=(10, 90))
    assert isinstance(lazy.results, DeviceResults) and not lazy.results.materialised
    assert lazy.results.device.is_cuda and len(lazy.results) == 20_000 and lazy.results.size == 20_000
    assert (lazy.mean, lazy.std, lazy.percentiles) == (plain.mean, plain.std, plain.percentiles)
    assert lazy.stats["ci_mean"] == plain.stats["ci_mean"]
    assert fw.compare_results([bs.name], "p90")[bs.name] == plain.percentiles[90]
    assert not lazy.results.materialised  # nothing above needed the host copy
    assert np.array_equal(np.asarray(lazy.results), plain.results) and lazy.results.materialised
    assert lazy.results[3] == plain.results[3] and float(np.mean(lazy.results)) == float(np.mean(plain.results))
    np.testing.assert_allclose(np.asarray(lazy.results), a["G4_bs_call_par"], rtol=RES_RTOL, atol=PAY_ATOL)
    bs.keep_results_on_device = False


def test_declared_draw_programs_run_as_kernels_and_match_the_python_hooks(m, golden):
    """user simulations that DECLARE their draws (dice of core.py:22-25, coin-flip streak of the
    getting-started guide, SimpleSim of the reference tests) run on the device, bit-identical to looping their Python
    hook like the reference does -- sequential PCG64 layout, Philox blocks, odd sizes (half-draw buffer), and a span
    whose Lemire rejections force the host loop."""
    from mcframework_b200 import _device as D, programs

    _, a = golden

    class Dice(m.MonteCarloSimulation):
        size = 2

        def single_simulation(self, _rng=None, **kw):
            return float(self._rng(_rng, self.rng).integers(1, 7, size=self.size).sum())

    class DeclaredDice(Dice):
        def draw_program(self, **kw):
            return programs.integers_sum(1, 7, size=self.size)

    class Streak(m.MonteCarloSimulation):
        def single_simulation(self, n_flips=100, _rng=None, **kw):
            best = cur = 0
            for f in self._rng(_rng, self.rng).integers(0, 2, size=n_flips):
                cur = cur + 1 if f == 1 else 0
                best = max(best, cur)
            return float(best)

    class DeclaredStreak(Streak):
        def draw_program(self, n_flips=100, **kw):
            return programs.integers_longest_run(0, 2, size=n_flips, value=1)

    class Simple(m.MonteCarloSimulation):
        def single_simulation(self, mean=0.0, std=1.0, _rng=None, **kw):
            return float(self._rng(_rng, self.rng).normal(mean, std))

    class DeclaredSimple(Simple):
        def draw_program(self, mean=0.0, std=1.0, **kw):
            return programs.normal(mean, std)

    class Wide(Dice):  # 3e9 values: 30 % of the slots are rejected -> the engine must fall back to the hook
        def single_simulation(self, _rng=None, **kw):
            return float(self._rng(_rng, self.rng).integers(0, 3_000_000_000, size=3).sum())

        def draw_program(self, **kw):
            return programs.integers_sum(0, 3_000_000_000, size=3)

    before = D.LAUNCHES["pi"]
    d = DeclaredDice()
    d.set_seed(42)
    assert np.array_equal(d.run(10_000, compute_stats=False).results, a["G3_dice_seq"])
    d.set_seed(42)
    assert np.array_equal(d.run(20_000, parallel=True, n_workers=4, compute_stats=False).results, a["G3_dice_par"])
    assert D.LAUNCHES["pi"] == before + 2  # both ran as kernels
    for size, n in ((3, 1001), (1, 777)):  # odd slot counts: replications start in the middle of a 64-bit draw
        ref, dev = Dice(), DeclaredDice()
        ref.size = dev.size = size
        ref.set_seed(7)
        dev.set_seed(7)
        r1, r2 = ref.run(n, compute_stats=False), dev.run(n, compute_stats=False)
        assert np.array_equal(r1.results, r2.results)
        assert np.array_equal(ref.run(5, compute_stats=False).results, dev.run(5, compute_stats=False).results)  # stream continues
    for kw in (dict(), dict(parallel=True, n_workers=3)):
        n = 20_011 if kw else 3000
        ref, dev = Streak(), DeclaredStreak()
        ref.set_seed(5)
        dev.set_seed(5)
        assert np.array_equal(ref.run(n, n_flips=37, compute_stats=False, **kw).results,
                              dev.run(n, n_flips=37, compute_stats=False, **kw).results)
        ref, dev = Simple(), DeclaredSimple()
        ref.set_seed(9)
        dev.set_seed(9)
        assert np.array_equal(ref.run(n, mean=5.0, std=2.0, compute_stats=False, **kw).results,
                              dev.run(n, mean=5.0, std=2.0, compute_stats=False, **kw).results)
    ref, dev = Wide(), Wide()
    ref.draw_program = lambda **kw: None
    ref.set_seed(3)
    dev.set_seed(3)
    assert np.array_equal(ref.run(500, compute_stats=False).results, dev.run(500, compute_stats=False).results)
    assert np.array_equal(ref.run(3, compute_stats=False).results, dev.run(3, compute_stats=False).results)


def test_greeks(m, golden):
    j, _ = golden
    bs = m.BlackScholesSimulation()
    bs.set_seed(9)
    before = bs.rng.bit_generator.state
    g = bs.calculate_greeks(1000)
    for k, w in j["G4_greeks_call"].items():
        np.testing.assert_allclose(g[k], w, rtol=1e-9, err_msg=k)  # differences of means: cancellation
    assert bs.rng.bit_generator.state == before and bs.seed_seq.entropy == 42
    g = bs.calculate_greeks(20_000, option_type="put", n_steps=12, parallel=True, n_workers=j["G4_greeks_put_par_cpu_count"])
    for k, w in j["G4_greeks_put_par"].items():
        np.testing.assert_allclose(g[k], w, rtol=1e-9, err_msg=k)


def test_portfolio_and_paths_and_lsm(m, golden):
    j, a = golden
    pf = m.PortfolioSimulation()
    pf.set_seed(42)
    np.testing.assert_allclose(pf.run(500, years=5, compute_stats=False).results, a["G5_portfolio_gbm_seq"], rtol=RES_RTOL)
    pf.set_seed(42)
    np.testing.assert_allclose(pf.run(500, years=2, use_gbm=False, compute_stats=False).results,
                               a["G5_portfolio_log1p_seq"], rtol=RES_RTOL)
    ps = m.BlackScholesPathSimulation()
    ps.set_seed(5)
    np.testing.assert_allclose(ps.run(300, n_steps=30, compute_stats=False).results, a["G6_path_terminal_seq"], rtol=RES_RTOL)
    ps.set_seed(1)
    paths = ps.simulate_paths(4000, n_steps=50)
    assert paths.shape == (4000, 51)
    np.testing.assert_allclose(paths[:8], a["G6_paths_head"], rtol=RES_RTOL)
    np.testing.assert_allclose(paths[-1], a["G6_paths_tail_row"], rtol=RES_RTOL)
    from mcframework_b200.sims import _american_exercise_lsm, _simulate_gbm_path

    np.testing.assert_allclose(_american_exercise_lsm(paths, 100.0, 0.05, 1.0 / 50, "put"), j["G6"]["lsm_put"], rtol=1e-9)
    np.testing.assert_allclose(_american_exercise_lsm(paths, 100.0, 0.05, 1.0 / 50, "call"), j["G6"]["lsm_call"], rtol=1e-9)
    np.testing.assert_allclose(_american_exercise_lsm(paths, 110.0, 0.05, 1.0 / 50, "put"), j["G6"]["lsm_put_K110"], rtol=1e-9)
    rng = np.random.default_rng(42)
    small = np.array([_simulate_gbm_path(100.0, 0.05, 0.2, 1.0, 50, rng) for _ in range(100)])
    np.testing.assert_allclose(_american_exercise_lsm(small, 100.0, 0.05, 1.0 / 50, "put"), j["G6"]["lsm_small_put"], rtol=1e-9)
    np.testing.assert_allclose(_american_exercise_lsm(small, 100.0, 0.05, 1.0 / 50, "call"), j["G6"]["lsm_small_call"], rtol=1e-9)
    # device-resident variant: time-major paths straight into the LSM sweep
    ps.set_seed(1)
    tm = ps.simulate_paths(4000, n_steps=50, device=True, time_major=True)
    np.testing.assert_allclose(_american_exercise_lsm(tm, 100.0, 0.05, 1.0 / 50, "put", time_major=True), j["G6"]["lsm_put"], rtol=1e-9)


def test_stats_engine_golden(m, golden):
    j, a = golden
    from mcframework_b200.stats_engine import build_default_engine

    eng = build_default_engine()
    x = np.random.default_rng(0).normal(5, 2, 100_000)
    ctx = m.StatsContext(n=x.size, percentiles=(1, 5, 50, 95, 99), rng=7, n_bootstrap=1000, bootstrap="bca", target=5.0, eps=0.01)
    out = eng.compute(x, ctx)
    assert not out.errors and not out.skipped
    want = dict(j["G7"])
    want["percentiles"] = {int(k): v for k, v in want["percentiles"].items()}
    assert out.metrics["percentiles"] == want["percentiles"]
    _close_stats(out.metrics, want)
    xs = a["G7_omit_x"]
    ctx = m.StatsContext(n=xs.size, nan_policy="omit", rng=11, n_bootstrap=500, bootstrap="percentile", confidence=0.9,
                         target=2.0, eps=0.1, ci_method="t")
    want = dict(j["G7_omit"])
    want["percentiles"] = {int(k): v for k, v in want["percentiles"].items()}
    _close_stats(eng.compute(xs, ctx).metrics, want)
    six = np.linspace(0, 1, 6)
    got = eng.compute(six, m.StatsContext(n=6, rng=123, n_bootstrap=200), select=["ci_mean_bootstrap"]).metrics
    _close_stats(got, j["G8"]["percentile_rng123_B200"])
    got = eng.compute(six, m.StatsContext(n=6, rng=321, n_bootstrap=200, bootstrap="bca"), select=["ci_mean_bootstrap"]).metrics
    _close_stats(got, j["G8"]["bca_rng321_B200"])


def test_engine_taxonomy_and_small_samples(m):
    from mcframework_b200 import stats_engine as se

    eng = se.build_default_engine()
    out = eng.compute(np.array([1.0, 2.0, 3.0, 4.0]), m.StatsContext(n=4, rng=1, n_bootstrap=100))
    skipped = dict(out.skipped)
    assert set(skipped) == {"chebyshev_required_n", "markov_error_prob", "bias_to_target", "mse_to_target"}
    want = dict(zip((5, 25, 50, 75, 95), map(float, np.percentile([1.0, 2.0, 3.0, 4.0], (5, 25, 50, 75, 95)))))
    assert out.metrics["mean"] == 2.5 and out.metrics["percentiles"] == want
    assert out.metrics["skew"] == 0.0 and out.metrics["ci_mean"]["method"] == "t"
    assert se.std(np.array([1.0]), m.StatsContext(n=1)) is None and se.mean(np.array([]), {"n": 0}) is None
    assert se.skew(np.array([1.0, 2.0]), None) == 0.0 and se.kurtosis(np.array([1.0, 2.0, 3.0]), None) == 0.0
    assert math.isnan(se.mean(np.array([1.0, np.nan]), None))
    assert se.mean(np.array([1.0, np.nan, 3.0]), {"nan_policy": "omit"}) == 2.0
    with pytest.raises(ValueError, match="method must be 'auto', 'z', or 't'"):
        se.ci_mean(np.arange(10.0), m.StatsContext(n=10, ci_method="bootstrap"))
    assert se.ci_mean_bootstrap(np.array([]), {"n": 0}) is None
    # a user metric gets the host array and can sit next to device metrics
    eng2 = m.StatsEngine([m.FnMetric("mean", se.mean), m.FnMetric("peak", lambda x, ctx: float(np.max(x)))])
    assert eng2.compute(np.array([1.0, 5.0, 3.0])).metrics == {"mean": 3.0, "peak": 5.0}


def test_portfolio_simulate_slices_extension(golden):
    """``PortfolioSimulation.simulate_slices`` (extension; closest reference shape ``simulate_paths``,
    black_scholes.py:403-418): the last slice of a block-parallel call equals the reference's ``run`` results (fixture
    G5_portfolio_gbm_par), both layouts agree, ragged path counts and strides fall back to the per-draw kernel, fp32
    narrows only the store, the sequential form consumes ``self.rng`` like ``run`` does."""
    import torch

    import mcframework_b200 as m

    _, a = golden
    pf = m.PortfolioSimulation()
    pf.set_seed(3)
    sl = pf.simulate_slices(20_000, years=1, slice_stride=21, parallel=True, n_workers=8)
    assert sl.shape == (13, 20_000) and sl.dtype == np.float64
    np.testing.assert_allclose(sl[-1], a["G5_portfolio_gbm_par"], rtol=1e-12)
    assert np.all(sl[0] == 10_000.0) and np.all(sl > 0)
    pf.set_seed(3)
    rm = pf.simulate_slices(20_000, years=1, slice_stride=21, parallel=True, n_workers=8, time_major=False)
    assert rm.shape == (20_000, 13) and np.array_equal(rm.T, sl)
    pf.set_seed(3)
    f32 = pf.simulate_slices(20_000, years=1, slice_stride=21, parallel=True, n_workers=8, dtype="float32", device=True)
    assert f32.dtype == torch.float32 and np.array_equal(f32.cpu().numpy(), sl.astype(np.float32))
    pf.set_seed(3)
    odd = pf.simulate_slices(20_001, years=1, slice_stride=21, parallel=True, n_workers=3)  # 20 001 % 2 == 1: scalar stores
    pf.set_seed(3)
    want = pf.run(20_001, years=1, parallel=True, n_workers=3, compute_stats=False).results
    np.testing.assert_allclose(odd[-1], want, rtol=1e-12)
    pf.set_seed(3)
    s5 = pf.simulate_slices(20_000, years=1, slice_stride=5, parallel=True, n_workers=8)  # stride < 16: every draw replayed
    assert s5.shape == (51, 20_000)
    np.testing.assert_allclose(s5[21], sl[5], rtol=1e-12)  # step 105 is a slice of both
    # sequential form: one PCG64 stream, generator left where ``run`` leaves it
    p1, p2 = m.PortfolioSimulation(), m.PortfolioSimulation()
    p1.set_seed(42), p2.set_seed(42)
    seq = p1.simulate_slices(500, years=5, slice_stride=252)
    np.testing.assert_allclose(seq[-1], a["G5_portfolio_gbm_seq"], rtol=1e-12)
    p2.run(500, years=5, compute_stats=False)
    assert p1.rng.bit_generator.state == p2.rng.bit_generator.state
    with pytest.raises(ValueError):
        pf.simulate_slices(10, dtype="float16")
