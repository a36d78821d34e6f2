"""GPU parity tests: CUDA kernels (through the C ABI) vs the CPU oracle and the reference's golden outputs.

Bars: integer outputs (Pi hits, replication start positions) bit-exact; floating-point values within
FP_RTOL relative (+ PAY_ATOL absolute on option payoffs, where S_T - K cancels).  The only differences
left are the summation order (sequential on device, NumPy pairwise in the reference) and <=1 ulp in exp/log1p.
"""
import numpy as np
import pytest

from conftest import have_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_cuda(), reason="needs a CUDA device")]

FP_RTOL = 1e-12
PAY_ATOL = 1e-11
BS = [100.0, 100.0, 1.0, 0.05, 0.2]


@pytest.fixture(scope="module")
def dev():
    from mcframework_b200 import _device, _native, _streams
    _native.lib()
    return _device, _native, _streams


def _seq(S, seed, n):
    return S.sequential_layout(np.random.default_rng(np.random.SeedSequence(seed)), n)


def test_abi_version(dev):
    _, N, _ = dev
    assert N.lib().mcf_abi_version() == 4
    assert b"sm_100a" in N.lib().mcf_build_info()


def test_pi_config1_bit_exact(dev, golden):
    D, _, S = dev
    j, a = golden
    hits, out = D.pi_hits(D.DeviceLayout.of(_seq(S, 123, 10_000)), 5000)
    hits, out = hits.cpu().numpy(), out.cpu().numpy()
    assert int(hits.sum()) == 39_268_872 == j["G1"]["sum_hits"]
    assert np.array_equal(out, a["G1_pi_results"])


def test_pi_philox_blocks_and_antithetic_bit_exact(dev, golden):
    D, _, S = dev
    _, a = golden
    lay = S.parallel_layout(np.random.SeedSequence(123), 20_000, 8)
    assert np.array_equal(D.pi_hits(D.DeviceLayout.of(lay), 100)[1].cpu().numpy(), a["G2_pi_parallel_results"])
    lay = S.parallel_layout(np.random.SeedSequence(5), 20_001, 3)
    assert np.array_equal(D.pi_hits(D.DeviceLayout.of(lay), 33, True)[1].cpu().numpy(),
                          a["G2b_pi_antithetic_parallel_results"])
    assert np.array_equal(D.pi_hits(D.DeviceLayout.of(_seq(S, 6, 500)), 101, True)[1].cpu().numpy(),
                          a["G2c_pi_antithetic_seq_results"])


@pytest.mark.parametrize("n_points", [1, 2, 31, 33, 8192, 8193, 20_000])
def test_pi_slicing_vs_oracle(dev, n_points):
    from oracle import c_oracle as co
    D, _, S = dev
    for anti in (False, True):
        hits = D.pi_hits(D.DeviceLayout.of(_seq(S, 77, 37)), n_points, anti)[0].cpu().numpy()
        assert np.array_equal(hits, co.Gen.pcg64(77).pi(37, n_points, anti)[1])


def test_normal_plan_positions_bit_exact(dev):
    from oracle import c_oracle as co
    D, _, S = dev
    for seed, n, steps in [(42, 1000, 252), (3, 5000, 7), (8, 3, 100_000)]:
        dl = D.DeviceLayout.of(_seq(S, seed, n))
        plan = D.normal_plan(dl, steps)
        g, want = co.Gen.pcg64(seed), []
        for _ in range(n):
            want.append(g.consumed)
            g.standard_normal(steps)
        assert np.array_equal(plan.sim_start.cpu().numpy().view(np.uint64), np.array(want, dtype=np.uint64))
        assert int(plan.stream_end.cpu().numpy()[0]) == g.consumed
    lay = S.parallel_layout(np.random.SeedSequence(11), 40_000, 4)
    dl = D.DeviceLayout.of(lay)
    plan = D.normal_plan(dl, 64)
    starts = plan.sim_start.cpu().numpy().view(np.uint64)
    for k in (0, 7, lay.n_streams - 1):
        rec = lay.records[k]
        g, want = co.Gen.philox_from_key(rec["s"][:2]), []
        for _ in range(int(rec["n_sims"])):
            want.append(g.consumed)
            g.standard_normal(64)
        f = int(rec["first_sim"])
        assert np.array_equal(starts[f:f + len(want)], np.array(want, dtype=np.uint64))


def _terminal(D, lay, steps, mode, p):
    dl = D.DeviceLayout.of(lay)
    return D.gbm_terminal(dl, steps, D.normal_plan(dl, steps), [(mode, p)])[0].cpu().numpy()


PF = [1e4, 0.07, 0.2]
TERMINAL_CASES = {
    "G4_bs_call_seq": (("seq", 42, 1000), 252, "BS_EURO_CALL", BS),
    "G4_bs_put_seq": (("seq", 42, 777), 61, "BS_EURO_PUT", [95, 105, .5, .03, .35]),
    "G4_bs_amer_put_seq": (("seq", 42, 1000), 252, "BS_AMER_PUT", BS),
    "G4_bs_amer_call_seq": (("seq", 42, 400), 40, "BS_AMER_CALL", BS),
    "G4_bs_call_par": (("par", 11, 20_000, 4), 64, "BS_EURO_CALL", BS),
    "G4_bs_amer_put_par": (("par", 11, 20_000, 2), 16, "BS_AMER_PUT", BS),
    "G5_portfolio_gbm_seq": (("seq", 42, 500), 1260, "PORTFOLIO_GBM", PF),
    "G5_portfolio_log1p_seq": (("seq", 42, 500), 504, "PORTFOLIO_LOG1P", PF),
    "G5_portfolio_gbm_par": (("par", 3, 20_000, 8), 252, "PORTFOLIO_GBM", PF),
    "G6_path_terminal_seq": (("seq", 5, 300), 30, "PATH_TERMINAL", [100, .05, .2, 1.0]),
}


@pytest.mark.parametrize("key", sorted(TERMINAL_CASES))
def test_terminal_values_golden(dev, golden, key):
    """Payoffs / terminal values of the reference run (tests/golden, produced by tools/gen_golden.py)."""
    D, N, S = dev
    _, a = golden
    how, steps, mode, p = TERMINAL_CASES[key]
    lay = _seq(S, how[1], how[2]) if how[0] == "seq" else S.parallel_layout(np.random.SeedSequence(how[1]), how[2], how[3])
    got = _terminal(D, lay, steps, getattr(N, mode), p)
    np.testing.assert_allclose(got, a[key], rtol=FP_RTOL, atol=PAY_ATOL if key.startswith("G4") else 0)


def test_paths_golden_both_layouts(dev, golden):
    D, _, S = dev
    j, a = golden
    dl = D.DeviceLayout.of(_seq(S, 1, 4000))
    plan = D.normal_plan(dl, 50)
    rows = D.gbm_paths(dl, 50, plan, 100.0, 0.05, 0.2, 1.0).cpu().numpy()
    np.testing.assert_allclose(rows[:8], a["G6_paths_head"], rtol=FP_RTOL)
    np.testing.assert_allclose(rows[-1], a["G6_paths_tail_row"], rtol=FP_RTOL)
    assert abs(rows.sum() - j["G6"]["paths_checksum"]) <= 1e-10 * abs(j["G6"]["paths_checksum"])
    cols = D.gbm_paths(dl, 50, plan, 100.0, 0.05, 0.2, 1.0, time_major=True).cpu().numpy()
    assert np.array_equal(cols.T, rows)


def test_common_random_numbers_specs(dev):
    """Several parameter sets in one launch (Greeks) == separate launches, within FP_RTOL."""
    D, N, S = dev
    dl = D.DeviceLayout.of(S.parallel_layout(np.random.SeedSequence(42), 30_000, 8))
    plan = D.normal_plan(dl, 32)
    sets = [(N.BS_EURO_CALL, BS), (N.BS_EURO_CALL, [101.0, 100, 1, .05, .2]), (N.BS_EURO_PUT, [100, 100, 1 - 1 / 365, .05, .2]),
            (N.BS_EURO_CALL, [100, 100, 1, .0505, .202])]
    fused = D.gbm_terminal(dl, 32, plan, sets).cpu().numpy()
    for k, s in enumerate(sets):
        single = D.gbm_terminal(dl, 32, plan, [s])[0].cpu().numpy()
        np.testing.assert_allclose(fused[k], single, rtol=FP_RTOL, atol=PAY_ATOL)


def test_large_run_sum_of_normals_matches_oracle_blockwise(dev):
    """Size-independent property at a larger size: per-replication normal sums of whole Philox blocks."""
    from oracle import c_oracle as co
    D, N, S = dev
    lay = S.parallel_layout(np.random.SeedSequence(2024), 400_000, 8)
    dl = D.DeviceLayout.of(lay)
    plan = D.normal_plan(dl, 252)
    sums = D.gbm_terminal(dl, 252, plan, [(N.NORMAL_SUM, [])])[0].cpu().numpy()
    for k in (0, 31, lay.n_streams - 1):
        rec = lay.records[k]
        n = int(rec["n_sims"])
        z = co.Gen.philox_from_key(rec["s"][:2]).standard_normal(n * 252).reshape(n, 252)
        f = int(rec["first_sim"])
        np.testing.assert_allclose(sums[f:f + n], z.sum(axis=1), rtol=0, atol=1e-11)
    assert abs(sums.mean()) < 5 * np.sqrt(252 / sums.size)


@pytest.mark.parametrize("steps", [15, 16, 17, 31, 32, 33, 47, 100])
def test_boundary_replay_sums_match_oracle_normals(dev, steps):
    """Affine bodies on aligned Philox streams regenerate only the shorter side of every replication boundary and take
    the rest from the stored sub-chunk sums (>= 16 steps; below that the direct replay runs): per-replication sums of
    NumPy's normals, for step counts around the 16-draw sub-chunk size and ragged stream lengths."""
    from oracle import c_oracle as co
    D, N, S = dev
    lay = S.parallel_layout(np.random.SeedSequence(700 + steps), 20_003, 5)
    dl = D.DeviceLayout.of(lay)
    plan = D.normal_plan(dl, steps)
    sums = D.gbm_terminal(dl, steps, plan, [(N.NORMAL_SUM, []), (N.BS_EURO_PUT, BS)]).cpu().numpy()
    for k in (0, lay.n_streams // 2, lay.n_streams - 1):
        rec = lay.records[k]
        n, f = int(rec["n_sims"]), int(rec["first_sim"])
        z = co.Gen.philox_from_key(rec["s"][:2]).standard_normal(n * steps).reshape(n, steps)
        np.testing.assert_allclose(sums[0, f:f + n], z.sum(axis=1), rtol=0, atol=1e-12 * max(1.0, steps ** 0.5))
    single = D.gbm_terminal(dl, steps, plan, [(N.BS_EURO_PUT, BS)])[0].cpu().numpy()
    assert np.array_equal(single, sums[1])  # deterministic: same launch shape or not, same bits


# ======================================================================= StatsEngine kernels
STAT_RTOL = 1e-11  # device one-pass sums vs NumPy's pairwise sums / SciPy's two-pass moments


def _dev_tensor(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


@pytest.fixture(scope="module")
def sd():
    from mcframework_b200 import _stats_device
    return _stats_device


@pytest.mark.parametrize("n", [1, 2, 3, 31, 1000, 100_003, 4_000_001])
def test_moments_vs_numpy_scipy(sd, n):
    import scipy.stats as st
    x = np.random.default_rng(n).normal(5.0, 2.0, n) ** 1
    m = sd.moments(_dev_tensor(x), target=4.5)
    assert m.n == n and m.n_nonfinite == 0
    np.testing.assert_allclose(m.mean, x.mean(), rtol=1e-14)
    np.testing.assert_allclose(m.total, x.sum(), rtol=1e-13)
    np.testing.assert_allclose(m.sq_target, ((x - 4.5) ** 2).sum(), rtol=1e-13)
    if n > 1:
        np.testing.assert_allclose(np.sqrt(m.m2 / (n - 1)), x.std(ddof=1), rtol=1e-13)
    if n > 3:
        m2, m3, m4 = m.m2 / n, m.m3 / n, m.m4 / n
        g1 = np.sqrt(n * (n - 1.0)) / (n - 2.0) * m3 / m2 ** 1.5
        g2 = 1.0 / (n - 2) / (n - 3) * ((n * n - 1.0) * m4 / m2 ** 2.0 - 3 * (n - 1) ** 2.0)
        np.testing.assert_allclose(g1, st.skew(x, bias=False), rtol=STAT_RTOL, atol=1e-13)
        np.testing.assert_allclose(g2, st.kurtosis(x, fisher=True, bias=False), rtol=STAT_RTOL, atol=1e-12)


def test_moments_nonfinite_and_skewed(sd):
    x = np.random.default_rng(0).lognormal(0, 1.5, 300_001)
    x[[5, 77, 1000]] = [np.nan, np.inf, -np.inf]
    x[2000] = np.nan
    m = sd.moments(_dev_tensor(x[1:]), target=0.0)  # odd offset: exercises the unaligned path
    f = x[1:][np.isfinite(x[1:])]
    assert (m.n, m.n_nan, m.n_posinf, m.n_neginf) == (f.size, 2, 1, 1)
    np.testing.assert_allclose(m.mean, f.mean(), rtol=1e-13)
    np.testing.assert_allclose(m.m2, ((f - f.mean()) ** 2).sum(), rtol=1e-12)
    np.testing.assert_allclose(m.m3, ((f - f.mean()) ** 3).sum(), rtol=1e-10)
    np.testing.assert_allclose(m.m4, ((f - f.mean()) ** 4).sum(), rtol=1e-10)


@pytest.mark.parametrize("n,dist", [(1, "normal"), (2, "normal"), (1001, "normal"), (250_000, "payoff"), (3_000_000, "normal"),
                                    (100_000, "const"), (65_537, "ints"),
                                    # >= 2^20 values: candidates are compacted after the second radix pass; tied data
                                    # (zeros, a constant, few distinct integers) overflow the scratch and keep reading x
                                    (2_500_000, "payoff"), (1_100_000, "const"), (1_300_001, "ints"), (1_048_576, "normal")])
def test_percentiles_bit_exact_vs_numpy(sd, n, dist):
    rng = np.random.default_rng(n)
    if dist == "normal":
        x = rng.normal(5, 2, n)
    elif dist == "payoff":
        x = np.maximum(rng.normal(0, 10, n), 0.0)  # ~50% exact zeros, like option payoffs
        x[::7] *= -0.0
    elif dist == "const":
        x = np.full(n, 3.25)
    else:
        x = rng.integers(-5, 5, n).astype(float)
    pcts = (0, 1, 5, 25, 50, 75, 95, 99, 100, 33.3)
    got = sd.percentiles(_dev_tensor(x), pcts)
    want = np.percentile(x, pcts)
    assert np.array_equal(got, want), (got, want)


def test_percentiles_nan_policies(sd):
    x = np.random.default_rng(3).normal(0, 1, 50_000)
    x[[3, 99, 4000]] = [np.nan, np.inf, -np.inf]
    xd = _dev_tensor(x)
    assert np.all(np.isnan(sd.percentiles(xd, (5, 50), has_nan=True)))
    f = x[np.isfinite(x)]
    assert np.array_equal(sd.percentiles(xd, (5, 50, 95), n_valid=f.size, omit_nonfinite=True), np.percentile(f, (5, 50, 95)))
    y = x.copy()
    y[3] = 1.0  # infinities only: they take part in the order
    with np.errstate(invalid="ignore"):  # lerp between two equal infinities is NaN, in NumPy too
        assert np.array_equal(sd.percentiles(_dev_tensor(y), (0, 50, 100)), np.percentile(y, (0, 50, 100)), equal_nan=True)


@pytest.mark.parametrize("n,B", [(6, 200), (1000, 150), (10_000, 100), (65_536, 30), (100_000, 50), (3_000_000, 4)])
def test_bootstrap_means_vs_numpy_stream(sd, n, B):
    """Indices are bit-exact (same accepted slots), so the means agree to summation-order rounding;
    the generator is left exactly where NumPy leaves it."""
    x = np.random.default_rng(1).normal(5, 2, n)
    g_ref, g_dev = np.random.default_rng(321), np.random.default_rng(321)
    want = np.stack([x[g_ref.integers(0, n, size=n)].mean() for _ in range(B)])
    got = sd.bootstrap_means(_dev_tensor(x), B, g_dev).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-13, atol=0)
    assert g_dev.bit_generator.state == g_ref.bit_generator.state
    # integer-valued data: sums are exact in fp64, so any index difference would show
    xi = np.random.default_rng(2).integers(0, 1 << 20, n).astype(float)
    g_ref, g_dev = np.random.default_rng(5), np.random.default_rng(5)
    g_ref.integers(0, 10, size=3), g_dev.integers(0, 10, size=3)  # leaves a buffered upper half (has_uint32 = 1)
    want = np.stack([xi[g_ref.integers(0, n, size=n)].sum() for _ in range(B)])
    got = sd.bootstrap_means(_dev_tensor(xi), B, g_dev).cpu().numpy() * n
    assert np.array_equal(np.rint(got), want)
    assert g_dev.bit_generator.state == g_ref.bit_generator.state


def test_bootstrap_philox_generator(sd):
    n, B = 5000, 120
    x = np.random.default_rng(1).integers(0, 1000, n).astype(float)
    g_ref, g_dev = (np.random.Generator(np.random.Philox(9)) for _ in range(2))
    g_ref.random(3), g_dev.random(3)  # mid-buffer start
    want = np.stack([x[g_ref.integers(0, n, size=n)].sum() for _ in range(B)])
    got = sd.bootstrap_means(_dev_tensor(x), B, g_dev).cpu().numpy() * n
    assert np.array_equal(np.rint(got), want)
    assert np.array_equal(g_dev.integers(0, 1 << 30, 5), g_ref.integers(0, 1 << 30, 5))


@pytest.mark.parametrize("n_paths,n_steps,put", [(100, 50, True), (100, 50, False), (20_000, 252, True), (3, 10, True), (5000, 12, False),
                                                 (800_001, 8, True), (1_000_000, 6, False)])
def test_lsm_vs_oracle(sd, n_paths, n_steps, put):
    """Odd path counts take the 8-byte form of the step kernel, even ones the 16-byte form; the two large cases span several
    bands per thread (unchecked full trips, the partial last band, both sweep directions)."""
    from oracle import np_port
    rng = np.random.default_rng(42)
    T, r, sigma, S0, K = 1.0, 0.05, 0.2, 100.0, 100.0
    dt = T / n_steps
    z = rng.standard_normal((n_paths, n_steps))
    paths = S0 * np.exp(np.concatenate([np.zeros((n_paths, 1)), np.cumsum((r - 0.5 * sigma ** 2) * dt + sigma * np.sqrt(dt) * z, axis=1)], axis=1))
    want, want_ex = np_port.lsm_price(paths, K, r, dt, "put" if put else "call")
    tm = sd.transpose(_dev_tensor(paths))
    assert np.array_equal(tm.cpu().numpy(), paths.T)
    price, ex = sd.lsm_price(tm, K, r, dt, put, return_exercise=True)
    np.testing.assert_allclose(float(price.cpu().numpy()[0]), want, rtol=1e-9)
    assert np.mean(ex.cpu().numpy() == want_ex) >= 0.9999  # exercise decisions: borderline flips only


def test_lsm_deep_otm_is_zero(sd):
    paths = np.full((500, 20), 100.0) * np.exp(np.random.default_rng(0).normal(0, 0.01, (500, 20)))
    got = float(sd.lsm_price(sd.transpose(_dev_tensor(paths)), 10.0, 0.05, 0.05, True).cpu().numpy()[0])
    assert got == 0.0


def test_count_less_and_compact(sd):
    x = np.random.default_rng(8).normal(0, 1, 70_001)
    x[::13] = np.nan
    x[5::101] = np.inf
    xd = _dev_tensor(x)
    assert sd.count_less(xd, 0.25) == int(np.sum(x < 0.25))
    assert np.array_equal(sd.compact_finite(xd).cpu().numpy(), x[np.isfinite(x)])


def test_path_slices_are_strided_paths(dev):
    """mcf_gbm_slices == every stride-th column of mcf_gbm_paths (same stream, same plan); portfolio mode
    == V0*exp(cumulative log-return) and its last slice equals the PortfolioSimulation terminal value."""
    D, N, S = dev
    dl = D.DeviceLayout.of(S.parallel_layout(np.random.SeedSequence(21), 20_000, 4))
    plan = D.normal_plan(dl, 48)
    full = D.gbm_paths(dl, 48, plan, 100.0, 0.05, 0.2, 1.0).cpu().numpy()
    tm = D.gbm_slices(dl, 48, plan, N.PATH_TERMINAL, [100.0, 0.05, 0.2, 1.0], 12, time_major=True).cpu().numpy()
    rm = D.gbm_slices(dl, 48, plan, N.PATH_TERMINAL, [100.0, 0.05, 0.2, 1.0], 12, time_major=False).cpu().numpy()
    assert tm.shape == (5, 20_000) and np.array_equal(tm.T, rm) and np.array_equal(rm, full[:, ::12])
    odd = D.gbm_slices(dl, 48, plan, N.PATH_TERMINAL, [100.0, 0.05, 0.2, 1.0], 7).cpu().numpy()
    assert np.array_equal(odd.T, full[:, ::7])
    pf = D.gbm_slices(dl, 48, plan, N.PORTFOLIO_GBM, [1e4, 0.07, 0.2], 48).cpu().numpy()
    term = D.gbm_terminal(dl, 48, plan, [(N.PORTFOLIO_GBM, [1e4, 0.07, 0.2])])[0].cpu().numpy()
    assert np.all(pf[0] == 1e4)
    np.testing.assert_allclose(pf[1], term, rtol=FP_RTOL)  # strictly sequential sum vs chunk sums


@pytest.mark.parametrize("steps,stride", [(96, 24), (64, 16), (2520, 21), (85, 17)])
def test_path_slices_from_segment_sums(dev, steps, stride):
    """Strides of at least one sub-chunk that divide the step count take the segment form (plan re-located for
    `stride` normals per pseudo-replication, boundary-form sums, one accumulation pass): same values as every
    stride-th column of the full paths, up to the summation order."""
    D, N, S = dev
    n = 20_000 if steps < 1000 else 20_000
    lay = S.parallel_layout(np.random.SeedSequence(33), n + 7, 3)
    dl = D.DeviceLayout.of(lay)
    plan = D.normal_plan(dl, steps)
    before = D.LAUNCHES["plan"]
    tm = D.gbm_slices(dl, steps, plan, N.PATH_TERMINAL, [100.0, 0.05, 0.2, 1.0], stride, time_major=True).cpu().numpy()
    assert D.LAUNCHES["plan"] == before + 1  # the re-locate: this call took the segment form
    rm = D.gbm_slices(dl, steps, plan, N.PATH_TERMINAL, [100.0, 0.05, 0.2, 1.0], stride, time_major=False).cpu().numpy()
    assert tm.shape == (steps // stride + 1, lay.n_total) and np.array_equal(tm.T, rm)
    full = D.gbm_paths(dl, steps, plan, 100.0, 0.05, 0.2, 1.0).cpu().numpy()
    np.testing.assert_allclose(rm, full[:, ::stride], rtol=1e-12)
    if steps == 2520:
        pf = D.gbm_slices(dl, steps, plan, N.PORTFOLIO_GBM, [1e4, 0.07, 0.2], stride).cpu().numpy()
        term = D.gbm_terminal(dl, steps, plan, [(N.PORTFOLIO_GBM, [1e4, 0.07, 0.2])])[0].cpu().numpy()
        assert np.all(pf[0] == 1e4)
        np.testing.assert_allclose(pf[-1], term, rtol=1e-11)


@pytest.mark.parametrize("n,B,shift,cs", [(50_000, 40, 12, 64), (120_000, 25, 13, 256), (3_000_000, 6, 18, 1024)])
def test_bootstrap_binned_path_is_the_same_index_stream(sd, n, B, shift, cs):
    """The binned gather (queues per resample and per region of x, regions summed one after the other) draws the
    very same indices: integer-valued data give exact sums, the generator ends in NumPy's state."""
    xi = np.random.default_rng(2).integers(0, 1 << 20, n).astype(float)
    g_ref, g_dev, g_dev2 = (np.random.default_rng(5) for _ in range(3))
    want = np.stack([xi[g_ref.integers(0, n, size=n)].sum() for _ in range(B)])
    before = dict(sd.BOOT_FORMS)
    got = sd.bootstrap_means(_dev_tensor(xi), B, g_dev, binned=True, region_shift=shift, chunk_slots=cs,
                             resamples_per_batch=7).cpu().numpy() * n
    assert np.array_equal(np.rint(got), want)
    assert g_dev.bit_generator.state == g_ref.bit_generator.state
    # the single-pass form (safe CTAs bin without exact ordinals, the rest are counted first) decided this run ...
    assert sd.BOOT_FORMS["single_pass"] == before["single_pass"] + 1 and sd.BOOT_FORMS["two_pass"] == before["two_pass"]
    # ... and the two-pass form (count pass over every slot first) gives the same sums and the same generator state
    g_two = np.random.default_rng(5)
    two = sd.bootstrap_means(_dev_tensor(xi), B, g_two, binned=True, region_shift=shift, chunk_slots=cs, resamples_per_batch=7,
                             single_pass=False).cpu().numpy() * n
    assert np.array_equal(np.rint(two), want) and g_two.bit_generator.state == g_ref.bit_generator.state
    plain = sd.bootstrap_means(_dev_tensor(xi), B, g_dev2, binned=False, chunk_slots=cs).cpu().numpy() * n
    assert np.array_equal(np.rint(plain), want)


@pytest.mark.parametrize("case", ["normal", "lognormal_heavy_tail", "half_zeros", "sorted", "constant", "with_infs", "few_levels"])
def test_bracketed_selection_is_exact(sd, case):
    """Large samples take the one-pass bracketed selection (mcf_select_bracketed: sample brackets, exact counts, radix
    passes on the candidates only).  Bit-exact against np.percentile whatever the shape of the data; heavy ties and
    constant data are the cases it hands back to the plain radix path (same answer either way)."""
    n = 9_000_001
    rng = np.random.default_rng(17)
    if case == "normal":
        x = rng.normal(5, 2, n)
    elif case == "lognormal_heavy_tail":
        x = np.exp(rng.normal(0, 3, n))
    elif case == "half_zeros":  # option payoffs: a point mass at 0 heavier than the candidate scratch
        x = np.maximum(rng.normal(0, 1, n), 0.0)
    elif case == "sorted":
        x = np.sort(rng.normal(0, 1, n))
    elif case == "constant":
        x = np.full(n, 3.25)
    elif case == "with_infs":
        x = rng.normal(0, 1, n)
        x[::1000] = np.inf
        x[5::1000] = -np.inf
    else:
        x = rng.integers(0, 7, n).astype(float)
    xd = _dev_tensor(x)
    pcts = (0, 0.001, 1, 5, 25, 50, 75, 95, 99, 99.999, 100)[:8] if case != "normal" else (1, 5, 25, 50, 75, 95, 99)
    got = sd.percentiles(xd, pcts)
    want = np.percentile(x, pcts)
    assert np.array_equal(got, want, equal_nan=True), (case, got, want)  # (inf - inf in NumPy's lerp is NaN on both sides)
    # the direct entry: order statistics at ranks including both ends
    ranks = [0, 1, n // 3, n // 2, n - 2, n - 1]
    direct = sd.select_bracketed(xd, ranks)
    ref = np.sort(x)[ranks]
    if direct is not None:
        assert np.array_equal(direct.numpy(), ref), case
    else:
        assert case in ("half_zeros", "constant", "few_levels", "with_infs", "sorted", "lognormal_heavy_tail", "normal")
    assert np.array_equal(sd.select(xd, ranks).cpu().numpy(), ref)
    if case == "normal":  # the smooth case must really be decided by the one-pass form
        assert sd.select_bracketed(xd, [n // 100, n // 2, (99 * n) // 100]) is not None
