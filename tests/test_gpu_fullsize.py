"""Full-size runs (BASELINE.json configs) checked through size-independent properties -- the oracle cannot
finish these sizes, so parity here is: agreement with closed forms within the Monte Carlo standard error,
exact internal consistency (order statistics vs counts, chunk-sum replay vs draw-by-draw replay on sampled
replications), and identities that hold for any seed.
"""
import logging
import math

import numpy as np
import pytest

from conftest import have_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_cuda(), reason="needs a CUDA device")]


def _bs_call(S0, K, T, r, s):
    from math import erf, exp, log, sqrt
    N = lambda x: 0.5 * (1.0 + erf(x / sqrt(2.0)))
    d1 = (log(S0 / K) + (r + 0.5 * s * s) * T) / (s * sqrt(T))
    return S0 * N(d1) - K * exp(-r * T) * N(d1 - s * sqrt(T))


@pytest.fixture(scope="module")
def m():
    import mcframework_b200 as mod
    logging.getLogger("mcframework_b200.core").setLevel(logging.WARNING)
    return mod


def test_config2_black_scholes_1e8_paths(m):
    """1e8 paths x 252 steps (config 2): price within 4 SE of Black-Scholes, put-call parity on common random
    numbers, and the chunk-sum replay equals the draw-by-draw replay on the first and last blocks."""
    import torch
    from mcframework_b200 import _device as D, _native as N, _stats_device as SD, _streams as S

    n, steps = 100_000_000, 252
    P = [100.0, 100.0, 1.0, 0.05, 0.2]
    lay = S.parallel_layout(np.random.SeedSequence(42), n, 8)
    dl = D.DeviceLayout.of(lay)
    plan = D.normal_plan(dl, steps)
    both = D.gbm_terminal(dl, steps, plan, [(N.BS_EURO_CALL, P), (N.BS_EURO_PUT, P), (N.NORMAL_SUM, [])])
    call, put, zsum = (SD.moments(both[k]) for k in range(3))
    se = math.sqrt(call.m2 / (n - 1) / n)
    assert abs(call.mean - _bs_call(*P)) < 4 * se
    # put-call parity pathwise: call - put = exp(-rT) (S_T - K); its mean is S0 - K exp(-rT) within 4 SE of S_T
    parity = call.mean - put.mean
    se_st = 100.0 * math.sqrt(math.exp(0.04) - 1.0) / math.sqrt(n)
    assert abs(parity - (100.0 - 100.0 * math.exp(-0.05))) < 4 * se_st
    # sum of 252 standard normals: mean 0, variance 252, no skew, excess kurtosis 0
    assert abs(zsum.mean) < 5 * math.sqrt(252 / n)
    assert abs(zsum.m2 / (n - 1) - 252.0) < 5 * 252.0 * math.sqrt(2.0 / n)
    # every replication consumed a plausible number of draws and the ranges tile each stream exactly
    starts = plan.sim_start.view(torch.int64)
    first = int(lay.records["n_sims"][0])
    d = (starts[1:first] - starts[: first - 1]).cpu().numpy()
    assert d.min() >= steps and d.max() < steps + 45 and 1.018 * steps < d.mean() < 1.026 * steps  # ~2.2 % extra draws
    # chunk-sum replay (affine body) vs draw-by-draw replay (path-dependent body family) on sampled replications:
    # PORTFOLIO_LOG1P with sigma=0 gives exp(n*log1p(mu*dt)) -- instead compare NORMAL_SUM with a full replay through paths
    sub = S.StreamLayout(lay.kind, lay.records[:1].copy(), 4096, 0)
    sub.records["n_sims"] = 4096
    dls = D.DeviceLayout.of(sub)
    plans = D.normal_plan(dls, steps)
    z_fast = D.gbm_terminal(dls, steps, plans, [(N.NORMAL_SUM, [])])[0].cpu().numpy()
    paths = D.gbm_paths(dls, steps, plans, 1.0, 0.0, 1.0, float(steps)).cpu().numpy()  # dt = 1, a = -0.5, b = 1
    z_slow = np.log(paths[:, -1]) + 0.5 * steps
    np.testing.assert_allclose(z_fast, z_slow, rtol=0, atol=2e-11)
    np.testing.assert_allclose(z_fast, both[2][:4096].cpu().numpy(), rtol=0, atol=0)  # same streams, same plan geometry


def test_large_runs_pipeline_their_host_copy(m):
    """2^24 + 5 paths through the public API (above core._PIPELINE_MIN_SIMS: four stream groups, host copy of each
    behind the kernels of the next) == one plan + replay over the whole layout, bit for bit."""
    from mcframework_b200 import _device as D, _native as N, _streams as S

    n, steps = (1 << 24) + 5, 16
    P = [100.0, 100.0, 1.0, 0.05, 0.2]
    bs = m.BlackScholesSimulation()
    assert n >= bs._PIPELINE_MIN_SIMS
    bs.set_seed(42)
    res = bs.run(n, parallel=True, n_workers=8, n_steps=steps, compute_stats=False)
    lay = S.parallel_layout(np.random.SeedSequence(42), n, 8)
    dl = D.DeviceLayout.of(lay)
    whole = D.gbm_terminal(dl, steps, D.normal_plan(dl, steps), [(N.BS_EURO_CALL, P)])[0].cpu().numpy()
    assert res.results.shape == (n,) and np.array_equal(res.results, whole)


def test_config1_pi_scaled_2e6_sims(m):
    from mcframework_b200 import _device as D, _stats_device as SD, _streams as S

    n = 2_000_000
    hits, est = D.pi_hits(D.DeviceLayout.of(S.parallel_layout(np.random.SeedSequence(123), n, 8)), 5000)
    s = SD.moments(est)
    assert abs(s.mean - math.pi) < 5 * math.sqrt(math.pi * (4 - math.pi) / 5000 / n)
    assert int(hits.min()) >= 3600 and int(hits.max()) <= 4250
    # the first block is the same stream as a 20 000-sim run with the same seed and worker count? no: block sizes differ;
    # identity that must hold instead: estimates are exactly 4*hits/5000
    import torch
    n_pts = torch.full_like(est, 5000.0)  # tensor / tensor is a correctly rounded division (a scalar divisor is not)
    assert bool((est == torch.div(hits.double() * 4.0, n_pts)).all())


def test_config5_stats_on_1e8_results(m):
    """Percentiles are exact order statistics (rank identities), moments agree with two-pass torch reductions,
    a B=16 bootstrap on the full 1e8 vector reproduces NumPy's index stream for its first resample."""
    import torch
    from mcframework_b200 import _stats_device as SD

    n = 100_000_000
    x = torch.from_numpy(np.random.default_rng(0).normal(5, 2, n)).cuda()
    s = SD.moments(x, target=5.0)
    assert s.n == n
    np.testing.assert_allclose(s.mean, float(x.mean()), rtol=1e-12)
    np.testing.assert_allclose(math.sqrt(s.m2 / (n - 1)), float(x.std()), rtol=1e-11)
    np.testing.assert_allclose(s.sq_target / n, float(((x - 5.0) ** 2).mean()), rtol=1e-11)
    pcts = (1, 5, 50, 95, 99)
    vals = SD.percentiles(x, pcts)
    prev, nxt, gamma = SD.quantile_plan(n, pcts)
    order = SD.select(x, list(prev) + list(nxt)).cpu().numpy()
    for k, p in enumerate(pcts):
        lo, hi = order[k], order[k + len(pcts)]
        assert lo <= vals[k] <= hi
        assert SD.count_less(x, lo) <= prev[k] < SD.count_less(x, float(np.nextafter(lo, np.inf)))  # lo has rank prev
    # bootstrap: resample 0 of NumPy's (B, n) matrix, reproduced on the host with the same generator
    g_dev, g_ref = np.random.default_rng(7), np.random.default_rng(7)
    means = SD.bootstrap_means(x, 16, g_dev).cpu().numpy()
    idx = g_ref.integers(0, n, size=n)
    want0 = float(x[torch.from_numpy(idx).cuda()].sum()) / n
    np.testing.assert_allclose(means[0], want0, rtol=1e-12)
    assert abs(means.mean() - s.mean) < 6 * 2.0 / math.sqrt(n)


def test_config4_lsm_1e6_paths_vs_the_unmodified_reference(m):
    """1e6 paths x 252 steps, seed 1: ``simulate_paths`` + Longstaff-Schwartz against the fixture the UNMODIFIED reference
    produced (tools/gen_golden.py --lsm-large -> tests/golden/reference_lsm_1e6.json; black_scholes.py:167-193).
    Bars: path matrix checksums rtol 1e-12, price rtol 1e-9, >= 99.99 % identical exercise steps (histograms)."""
    import json
    import os

    import torch
    from mcframework_b200 import _stats_device as SD

    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_lsm_1e6.json")))
    n, steps = g["n_paths"], g["n_steps"]
    ps = m.BlackScholesPathSimulation()
    ps.set_seed(g["seed"])
    tm = ps.simulate_paths(n, n_steps=steps, device=True, time_major=True)
    assert tuple(tm.shape) == (steps + 1, n)
    np.testing.assert_allclose(float(tm.sum(dtype=torch.float64)), g["paths_checksum"], rtol=1e-12)
    np.testing.assert_allclose(float(tm[-1, -1]), g["paths_last"], rtol=1e-12)
    for t, want in g["paths_col_means"].items():
        np.testing.assert_allclose(float(tm[int(t)].mean(dtype=torch.float64)), want, rtol=1e-12)
    for opt in ("put", "call"):
        price, ex = SD.lsm_price(tm, g["K"], g["r"], 1.0 / steps, opt == "put", return_exercise=True)
        want = g["lsm_" + opt]
        np.testing.assert_allclose(float(price.cpu()[0]), want["price"], rtol=1e-9)
        hist = torch.bincount(ex.to(torch.int64), minlength=steps + 1).cpu().numpy()
        moved = int(np.abs(hist - np.asarray(want["exercise_step_histogram"])).sum()) // 2
        assert moved <= 1e-4 * n, (opt, moved)


def test_config4_lsm_1e7_paths_full_size(m):
    """BASELINE config 4 at its full size, 1e7 paths x 252 steps (20 GB of time-major paths): the price agrees with the
    reference's 1e6-path price within the Monte Carlo error of both, lies above the European put, and two halves of the
    path set price alike."""
    import json
    import os

    from mcframework_b200 import _stats_device as SD
    from mcframework_b200.sims import _american_exercise_lsm

    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_lsm_1e6.json")))
    n = 10_000_000
    ps = m.BlackScholesPathSimulation()
    ps.set_seed(1)
    tm = ps.simulate_paths(n, n_steps=252, device=True, time_major=True)
    price = _american_exercise_lsm(tm, 100.0, 0.05, 1.0 / 252, "put", time_major=True)
    euro_put = _bs_call(100.0, 100.0, 1.0, 0.05, 0.2) - 100.0 + 100.0 * math.exp(-0.05)
    se = 7.6 * math.sqrt(1.0 / n + 1.0 / g["n_paths"])  # payoff standard deviation ~7.6
    assert euro_put < price and abs(price - g["lsm_put"]["price"]) < 4 * se, (price, g["lsm_put"]["price"], se)
    # the first 1e6 of the 1e7 paths ARE the fixture's paths (same seed, same sequential stream)
    first = float(SD.lsm_price(tm[:, :1_000_000].contiguous(), 100.0, 0.05, 1.0 / 252, True).cpu()[0])
    np.testing.assert_allclose(first, g["lsm_put"]["price"], rtol=1e-9)
    a = float(SD.lsm_price(tm[:, : n // 2].contiguous(), 100.0, 0.05, 1.0 / 252, True).cpu()[0])
    b = float(SD.lsm_price(tm[:, n // 2:].contiguous(), 100.0, 0.05, 1.0 / 252, True).cpu()[0])
    assert abs(a - b) < 5 * 7.6 * math.sqrt(4.0 / n)


def test_config3_portfolio_1e7_paths_x_2520_steps_stored_slices(m):
    """BASELINE config 3 at full size through the public extension ``PortfolioSimulation.simulate_slices``: 1e7 paths x
    2520 steps, a slice every 21 steps (121 slices, time-major).  Slice 0 is V0, the LAST slice is the result vector of
    ``run`` with the same seed (bit for bit: same plan, same replay), sampled streams equal the strided columns of the
    full paths regenerated draw by draw, and the fp32 store agrees with the fp64 one to float precision."""
    import torch
    from mcframework_b200 import _device as D, _native as N, _stats_device as SD, _streams as S

    n, years, stride = 10_000_000, 10, 21
    pf = m.PortfolioSimulation()
    pf.set_seed(42)
    sl = pf.simulate_slices(n, years=years, slice_stride=stride, parallel=True, n_workers=8, device=True)
    assert tuple(sl.shape) == (2520 // stride + 1, n) and sl.dtype == torch.float64
    assert bool((sl[0] == 10_000.0).all())
    pf.set_seed(42)
    pf.keep_results_on_device = True
    res = pf.run(n, parallel=True, n_workers=8, years=years, compute_stats=False)
    term = res.results.device  # core.DeviceResults: the vector stayed in HBM
    np.testing.assert_allclose(sl[-1].cpu().numpy(), term.cpu().numpy(), rtol=1e-12)
    mean_T = float(SD.moments(sl[-1]).mean)
    sd_T = 1e4 * math.exp(0.7) * math.sqrt(math.exp(0.04 * years) - 1.0)
    assert abs(mean_T - 1e4 * math.exp(0.07 * years)) < 5 * sd_T / math.sqrt(n)
    # sampled streams: the first 1024 paths of the first and of the last Philox block, every normal regenerated
    lay = S.parallel_layout(np.random.SeedSequence(42), n, 8)
    for k in (0, lay.n_streams - 1):
        sub = S.StreamLayout(lay.kind, lay.records[k:k + 1].copy(), 1024, 0)
        sub.records["first_sim"], sub.records["n_sims"] = 0, 1024
        dls = D.DeviceLayout.of(sub)
        plans = D.normal_plan(dls, 2520)
        mu, sig = 0.07, 0.2
        # gbm_paths parameters (S0, r, sigma, T) with T = 10 and dt = T/2520 = 1/252: the same per-step drift and scale
        full = D.gbm_paths(dls, 2520, plans, 1e4, mu, sig, float(years), time_major=True)
        f0 = int(lay.records["first_sim"][k])
        np.testing.assert_allclose(sl[:, f0:f0 + 1024].cpu().numpy(), full[::stride].cpu().numpy(), rtol=2e-12)
    del res, term
    pf.set_seed(42)
    sl32 = pf.simulate_slices(n, years=years, slice_stride=stride, parallel=True, n_workers=8, device=True, dtype="float32")
    assert sl32.dtype == torch.float32
    rel = ((sl32[-1].double() - sl[-1]).abs() / sl[-1]).max()
    assert float(rel) < 1e-7 * 1.0001  # one float rounding (2^-24 = 6e-8)


def test_config4_lsm_1e7_paths_smaller(m):
    """American put, 2e6 paths x 252 steps through the public API: above the European put, below the strike, and
    stable (same price from two different shards of a 2-way split of the paths within MC error)."""
    from mcframework_b200 import _stats_device as SD
    from mcframework_b200.sims import _american_exercise_lsm

    ps = m.BlackScholesPathSimulation()
    ps.set_seed(1)
    tm = ps.simulate_paths(2_000_000, n_steps=252, device=True, time_major=True)
    price = _american_exercise_lsm(tm, 100.0, 0.05, 1.0 / 252, "put", time_major=True)
    euro_put = _bs_call(100.0, 100.0, 1.0, 0.05, 0.2) - 100.0 + 100.0 * math.exp(-0.05)
    assert euro_put < price < 0.5 * 100.0 and abs(price - 6.08) < 0.05  # literature value ~6.08-6.09 (LSM biased low)
    a = float(SD.lsm_price(tm[:, :1_000_000].contiguous(), 100.0, 0.05, 1.0 / 252, True).cpu()[0])
    b = float(SD.lsm_price(tm[:, 1_000_000:].contiguous(), 100.0, 0.05, 1.0 / 252, True).cpu()[0])
    assert abs(a - b) < 0.03
