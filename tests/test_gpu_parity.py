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
    assert N.lib().mcf_abi_version() == 1
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


def test_black_scholes_golden(dev, golden):
    D, N, S = dev
    _, a = golden

    def run(lay, steps, mode, p):
        dl = D.DeviceLayout.of(lay)
        return D.gbm_terminal(dl, steps, D.normal_plan(dl, steps), [(mode, p)])[0].cpu().numpy()

    np.testing.assert_allclose(run(_seq(S, 42, 1000), 252, N.BS_EURO_CALL, BS), a["G4_bs_call_seq"],
                               rtol=FP_RTOL, atol=PAY_ATOL)
    np.testing.assert_allclose(run(_seq(S, 42, 777), 61, N.BS_EURO_PUT, [95, 105, .5, .03, .35]), a["G4_bs_put_seq"],
                               rtol=FP_RTOL, atol=PAY_ATOL)
    np.testing.assert_allclose(run(_seq(S, 42, 1000), 252, N.BS_AMER_PUT, BS), a["G4_bs_amer_put_seq"],
                               rtol=FP_RTOL, atol=PAY_ATOL)
    np.testing.assert_allclose(run(_seq(S, 42, 400), 40, N.BS_AMER_CALL, BS), a["G4_bs_amer_call_seq"],
                               rtol=FP_RTOL, atol=PAY_ATOL)
    np.testing.assert_allclose(run(S.parallel_layout(np.random.SeedSequence(11), 20_000, 4), 64, N.BS_EURO_CALL, BS),
                               a["G4_bs_call_par"], rtol=FP_RTOL, atol=PAY_ATOL)
    np.testing.assert_allclose(run(S.parallel_layout(np.random.SeedSequence(11), 20_000, 2), 16, N.BS_AMER_PUT, BS),
                               a["G4_bs_amer_put_par"], rtol=FP_RTOL, atol=PAY_ATOL)


def test_portfolio_and_path_golden(dev, golden):
    D, N, S = dev
    _, a = golden

    def run(lay, steps, mode, p):
        dl = D.DeviceLayout.of(lay)
        return D.gbm_terminal(dl, steps, D.normal_plan(dl, steps), [(mode, p)])[0].cpu().numpy()

    PF = [1e4, 0.07, 0.2]
    np.testing.assert_allclose(run(_seq(S, 42, 500), 1260, N.PORTFOLIO_GBM, PF), a["G5_portfolio_gbm_seq"], rtol=FP_RTOL)
    np.testing.assert_allclose(run(_seq(S, 42, 500), 504, N.PORTFOLIO_LOG1P, PF), a["G5_portfolio_log1p_seq"],
                               rtol=FP_RTOL)
    np.testing.assert_allclose(run(S.parallel_layout(np.random.SeedSequence(3), 20_000, 8), 252, N.PORTFOLIO_GBM, PF),
                               a["G5_portfolio_gbm_par"], rtol=FP_RTOL)
    np.testing.assert_allclose(run(_seq(S, 5, 300), 30, N.PATH_TERMINAL, [100, .05, .2, 1.0]),
                               a["G6_path_terminal_seq"], rtol=FP_RTOL)


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
