"""CPU-only tests: host logic, C-ABI surface, and the device functions compiled for the host (tests/hostemu).

Nothing here launches a kernel.  The hostemu harness compiles the SAME ``__host__ __device__`` functions the
kernels call (mcf_device.cuh) with g++ and walks them with the kernels' work decomposition, so stream
positions, jump-ahead and the planning pass are checked against the oracle without a GPU.
"""
import os
import re

import numpy as np
import pytest

from conftest import ROOT


# ------------------------------------------------------------------------------------- C ABI surface
def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mcf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from mcframework_b200 import _native, build

    build.build_native()
    lib = _native.load_library()  # dlopen only; no CUDA call
    names = _declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/mcf_b200.h but not exported"
        assert name in _native.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.mcf_abi_version() == 4
    assert lib.mcf_select_workspace_bytes() > 0 and lib.mcf_lsm_workspace_bytes(1000) > 0
    assert lib.mcf_bootstrap_workspace_bytes(10 ** 6, 64) > 0 and lib.mcf_moments_workspace_bytes(5) > 0


def test_product_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import mcframework_b200 as m

    sim = m.PiEstimationSimulation()
    sim.set_seed(1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sim.run(10, n_points=100)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.DEFAULT_ENGINE.compute(np.arange(10.0), m.StatsContext(n=10))


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mcframework_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


# ------------------------------------------------------------------------------------- stream layouts
def test_layouts_match_reference_block_rule():
    from mcframework_b200 import _streams as S

    lay = S.parallel_layout(np.random.SeedSequence(123), 20_000, 8)
    assert lay.n_streams == 65 and lay.hint == 312 and int(lay.records["n_sims"][-1]) == 32
    kids = np.random.SeedSequence(123).spawn(65)
    assert np.array_equal(lay.records["s"][:, :2], np.stack([k.generate_state(2, np.uint64) for k in kids]))
    parts = [lay.shard(r, 4) for r in range(4)]
    assert sum(p.n_total for p in parts) == 20_000 and all(int(p.records["first_sim"][0]) == 0 for p in parts)
    seq = S.sequential_layout(np.random.default_rng(5), 10)
    st = np.random.default_rng(5).bit_generator.state["state"]
    assert int(seq.records["s"][0, 0]) == st["state"] >> 64 and int(seq.records["s"][0, 3]) == st["inc"] & (2 ** 64 - 1)


def test_consume_draws_matches_numpy():
    from mcframework_b200._streams import consume_draws

    for mk in (lambda: np.random.default_rng(5), lambda: np.random.Generator(np.random.Philox(7))):
        for pre in (0, 1, 3, 6):
            for d in (1, 2, 4, 5, 9, 1000):
                a, b = mk(), mk()
                a.bit_generator.random_raw(pre), b.bit_generator.random_raw(pre)
                want = int(np.atleast_1d(a.bit_generator.random_raw(d))[-1])
                assert consume_draws(b, d) == want
                assert a.bit_generator.random_raw(6).tolist() == b.bit_generator.random_raw(6).tolist()


# ------------------------------------------------------------------------------------- quantile rule
def test_quantile_plan_and_lerp_are_numpy_percentile():
    from mcframework_b200._stats_device import lerp_like_numpy, quantile_plan

    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 10, 1001, 65_536):
        x = np.sort(rng.normal(0, 3, n))
        pcts = [0, 1, 2.5, 5, 25, 33.3, 50, 75, 95, 97.5, 99, 100]
        lo, hi, g = quantile_plan(n, pcts)
        assert np.array_equal(lerp_like_numpy(x[lo], x[hi], g), np.percentile(x, pcts))


def test_summary_merge_is_chan():
    from mcframework_b200._stats_device import MomentSummary, merge_summaries

    rng = np.random.default_rng(1)
    x = rng.lognormal(0, 1, 10_000)

    def summ(a):
        m = a.mean()
        return MomentSummary(a.size, 0, 0, 0, m, ((a - m) ** 2).sum(), ((a - m) ** 3).sum(), ((a - m) ** 4).sum(),
                             ((a - 1.0) ** 2).sum(), a.sum())

    whole = summ(x)
    merged = merge_summaries([summ(x[:100]), summ(x[100:7000]), summ(x[7000:7000]) if False else summ(x[7000:])])
    for f in ("n", "mean", "m2", "m3", "m4", "sq_target", "total"):
        np.testing.assert_allclose(getattr(merged, f), getattr(whole, f), rtol=1e-11)


# ------------------------------------------------------------------------------------- host API logic
def test_validation_messages_and_context():
    import mcframework_b200 as m
    from mcframework_b200.core import make_blocks

    sim = m.PiEstimationSimulation()
    for kw, msg in ((dict(n_simulations=0), "n_simulations must be positive"),
                    (dict(n_simulations=5, n_workers=0), "n_workers must be positive"),
                    (dict(n_simulations=5, confidence=1.0), r"confidence must be in the interval \(0, 1\)"),
                    (dict(n_simulations=5, ci_method="x"), "ci_method must be one of")):
        with pytest.raises(ValueError, match=msg):
            sim.run(**kw)
    assert make_blocks(10, 4) == [(0, 4), (4, 8), (8, 10)] and make_blocks(0, 3) == []
    with pytest.raises(ValueError, match="n_bootstrap .* should be >= 100"):
        m.StatsContext(n=10, n_bootstrap=50)
    with pytest.raises(ValueError, match="ess .* cannot exceed n"):
        m.StatsContext(n=10, ess=11)
    ctx = m.StatsContext(n=100, confidence=0.9, nan_policy="omit")
    assert ctx.q_bound() == (100.0 * ((1 - 0.9) / 2), 100.0 * (1 - (1 - 0.9) / 2)) and ctx.eff_n(7, 5) == 5
    fw = m.MonteCarloFramework()
    with pytest.raises(ValueError, match="Simulation 'nope' not found"):
        fw.run_simulation("nope", 10)
    fw.results["a"] = m.SimulationResult(np.zeros(4), 4, 0.0, 1.0, 2.0, {50: 1.5}, {}, {"requested_percentiles": [50]})
    assert fw.compare_results(["a"], "var") == {"a": 4.0} and fw.compare_results(["a"], "p50") == {"a": 1.5}
    with pytest.raises(ValueError, match="Percentile 75 not computed"):
        fw.compare_results(["a"], "p75")
    assert m.z_crit(0.95) == pytest.approx(1.959964, abs=1e-5) and m.autocrit(0.95, 10)[1] == "t"


def test_builtin_hooks_map_to_programs():
    import mcframework_b200 as m
    from mcframework_b200 import _native as N

    p = m.PortfolioSimulation()._program_for(dict(years=5, use_gbm=False))
    assert p.kind == "normals" and p.n_steps == int(5 / (1.0 / 252.0)) and p.specs[0][0] == N.PORTFOLIO_LOG1P
    # a horizon below one trading day has no steps: `rng.normal(size=0)` in the reference, a constant program here
    # (reference tests/test_errors_and_edge_cases.py::test_zero_year_portfolio)
    z = m.PortfolioSimulation()._program_for(dict(years=0, initial_value=10_000))
    assert z.kind == "constant" and z.value == 10_000.0
    assert m.PortfolioSimulation()._program_for(dict(years=0.003, use_gbm=False)).kind == "constant"
    with pytest.raises(ValueError):  # NumPy: "negative dimensions are not allowed"
        m.PortfolioSimulation()._program_for(dict(years=-1))
    b = m.BlackScholesSimulation()._program_for(dict(option_type="put", exercise_type="american", n_steps=10))
    assert b.specs[0][0] == N.BS_AMER_PUT and b.n_steps == 10
    with pytest.raises(ValueError, match="option_type must be 'call' or 'put'"):
        m.BlackScholesSimulation()._program_for(dict(option_type="straddle"))

    class Custom(m.PiEstimationSimulation):
        def single_simulation(self, **kw):
            return 1.0

    assert Custom()._program_for({}) is None  # replaced hook => user Python code, looped on the host


def test_declared_draw_programs_host_logic():
    """`draw_program` declarations: validation, and that a declaration wins over the 'hook was
    replaced => arbitrary Python' rule while an undeclared subclass still loops on the host."""
    import mcframework_b200 as mcf
    from mcframework_b200 import _native as N, programs

    p = programs.integers_sum(1, 7, size=2)
    assert (p.kind, p.low, p.high, p.size, p.op) == ("integers", 1, 7, 2, 0)
    q = programs.integers_longest_run(0, 2, size=100, value=1)
    assert (q.op, q.match, q.size) == (1, 1, 100)
    z = programs.normal(5.0, 2.0)
    assert z.kind == "normals" and z.n_steps == 1 and z.specs == [(N.NORMAL_LOC_SCALE, [5.0, 2.0])]
    for bad in (dict(low=0, high=1, size=3), dict(low=0, high=1 << 33, size=1), dict(low=0, high=6, size=0)):
        with pytest.raises(ValueError):
            programs.integers_sum(**bad)

    class Plain(mcf.MonteCarloSimulation):
        def single_simulation(self, _rng=None, **kw):
            return 1.0

    class Declared(Plain):
        def draw_program(self, **kw):
            return p

    assert Plain().draw_program() is None and Plain()._program_for({}) is None
    assert Declared()._program_for({}) is p


# ------------------------------------------------------------------------------------- hostemu
@pytest.fixture(scope="module")
def emu():
    import hostemu

    hostemu.build()
    return hostemu


def test_hostemu_pi_matches_oracle(emu):
    from mcframework_b200 import _streams as S
    from oracle import c_oracle as co

    for n_points, anti in ((100, False), (33, True), (8193, False)):
        lay = S.sequential_layout(np.random.default_rng(np.random.SeedSequence(77)), 23)
        assert np.array_equal(emu.pi_hits(lay, n_points, anti), co.Gen.pcg64(77).pi(23, n_points, anti)[1])
    lay = S.parallel_layout(np.random.SeedSequence(123), 20_000, 8)
    hits = emu.pi_hits(lay, 100)
    assert int(hits.sum()) == 1_570_354


def test_hostemu_normal_plan_matches_oracle(emu):
    from mcframework_b200 import _streams as S
    from oracle import c_oracle as co

    for seed, n, steps in ((42, 300, 252), (3, 2000, 7)):
        lay = S.sequential_layout(np.random.default_rng(np.random.SeedSequence(seed)), n)
        rc, starts, end, _ = emu.normal_plan(lay, steps)
        assert rc == 0
        g, want = co.Gen.pcg64(seed), []
        for _ in range(n):
            want.append(g.consumed)
            g.standard_normal(steps)
        assert np.array_equal(starts, np.array(want, dtype=np.uint64)) and int(end[0]) == g.consumed


def test_hostemu_terminal_values_match_reference_golden(emu, golden):
    """Plan + replay (block-synchronous Philox consumption, ziggurat state machine) on the host build of the
    device functions reproduces the reference's result vectors."""
    from mcframework_b200 import _native as N, _streams as S

    _, a = golden
    BS = [100.0, 100.0, 1.0, 0.05, 0.2]

    def seq(seed, n):
        return S.sequential_layout(np.random.default_rng(np.random.SeedSequence(seed)), n)

    cases = [(seq(42, 1000), 252, N.BS_EURO_CALL, BS, "G4_bs_call_seq"),
             (seq(42, 1000), 252, N.BS_AMER_PUT, BS, "G4_bs_amer_put_seq"),
             (S.parallel_layout(np.random.SeedSequence(11), 20_000, 4), 64, N.BS_EURO_CALL, BS, "G4_bs_call_par"),
             (S.parallel_layout(np.random.SeedSequence(11), 20_000, 2), 16, N.BS_AMER_PUT, BS, "G4_bs_amer_put_par"),
             (seq(42, 500), 504, N.PORTFOLIO_LOG1P, [1e4, .07, .2], "G5_portfolio_log1p_seq")]
    for lay, steps, mode, p, key in cases:
        plan = emu.normal_plan(lay, steps)
        assert plan.rc == 0
        masked = emu.gbm_terminal(lay, steps, plan, [(mode, p)])[0]  # what the kernels do: replay from the emit masks
        np.testing.assert_allclose(masked, a[key], rtol=1e-12, atol=1e-11)
        fsm = emu.gbm_terminal(lay, steps, plan, [(mode, p)], masked=False)[0]  # state machine, strictly sequential sums
        if mode in (N.BS_AMER_PUT, N.BS_AMER_CALL, N.PORTFOLIO_LOG1P):
            assert np.array_equal(masked, fsm)  # path-dependent bodies: same normals, same order, same bits
        else:  # affine bodies: chunk sums + partial chunks => different (deterministic) summation order
            np.testing.assert_allclose(masked, fsm, rtol=1e-12, atol=1e-11)
    # Greeks-style launch: several parameter sets from one replay of the normals
    lay = S.parallel_layout(np.random.SeedSequence(11), 20_000, 4)
    plan = emu.normal_plan(lay, 64)
    both = emu.gbm_terminal(lay, 64, plan, [(N.BS_EURO_CALL, BS), (N.BS_EURO_PUT, [101.0, 100.0, 1.0, 0.05, 0.2])])
    np.testing.assert_allclose(both[0], a["G4_bs_call_par"], rtol=1e-12, atol=1e-11)
    lay = S.parallel_layout(np.random.SeedSequence(5), 20_001, 3)
    assert np.array_equal(emu.pi_hits(lay, 33, True) * 4.0 / 33, a["G2b_pi_antithetic_parallel_results"])
    g = np.random.Generator(np.random.Philox(9))
    g.random(3)  # mid-buffer Philox state
    _, rec = S.generator_record(g)
    z, _ = emu.normals(rec, 0, 1000)
    assert np.array_equal(z, g.standard_normal(1000))


def test_hostemu_ziggurat_with_certified_wedge_bounds_is_numpy_bit_for_bit(emu):
    """2e6 normals per generator (~3e4 wedge tests, ~600 tail draws): every value equals NumPy's, so every
    accept/reject decision taken from the polynomial bounds instead of exp() was NumPy's decision."""
    from mcframework_b200 import _streams as S

    for gen in (np.random.default_rng(2024), np.random.Generator(np.random.Philox(77))):
        _, rec = S.generator_record(gen)
        z, end = emu.normals(rec, 0, 2_000_000)
        before = gen.bit_generator.state
        assert np.array_equal(z, gen.standard_normal(2_000_000))
        probe = np.random.Generator(type(gen.bit_generator)())
        probe.bit_generator.state = before
        S.consume_draws(probe, int(end))
        assert probe.bit_generator.random_raw(4).tolist() == gen.bit_generator.random_raw(4).tolist()


@pytest.mark.parametrize("steps", [1, 7, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 252, 1000])
def test_hostemu_chunk_sum_replay_equals_sequential_replay(emu, steps):
    """Affine bodies take whole chunks from the sums stored by the planning pass: same normals, different
    (deterministic) summation order.  Checked against the strictly sequential state-machine replay for
    replications shorter than / equal to / longer than a chunk, on both generators."""
    from mcframework_b200 import _native as N, _streams as S

    n = max(50, 30_000 // steps)
    for lay in (S.sequential_layout(np.random.default_rng(steps), n),
                S.parallel_layout(np.random.SeedSequence(steps), max(20_000, n), 3) if steps <= 100 else
                S.parallel_layout(np.random.SeedSequence(steps), 20_000, 64)):
        if lay.n_total * steps > 3_000_000:
            continue
        plan = emu.normal_plan(lay, steps)
        assert plan.rc == 0
        got = emu.gbm_terminal(lay, steps, plan, [(N.NORMAL_SUM, [])])[0]
        want = emu.gbm_terminal(lay, steps, plan, [(N.NORMAL_SUM, [])], masked=False)[0]
        np.testing.assert_allclose(got, want, rtol=0, atol=5e-13 * max(1.0, np.sqrt(steps)))


def test_hostemu_fast_philox_scan_positions_match_oracle(emu):
    """Block-aligned Philox streams take the SIMT-friendly chunk walk (parked wedge tests, skip-counter chain,
    invalid chunks re-walked through the queue): replication starts and stream ends must still be exact."""
    from mcframework_b200 import _streams as S
    from oracle import c_oracle as co

    for seed, n, workers, steps in ((11, 40_000, 4, 64), (5, 20_000, 8, 252), (9, 20_011, 2, 17)):
        lay = S.parallel_layout(np.random.SeedSequence(seed), n, workers)
        plan = emu.normal_plan(lay, steps)
        assert plan.rc == 0
        for k in (0, lay.n_streams // 2, lay.n_streams - 1):
            rec = lay.records[k]
            g, want = co.Gen.philox_from_key(rec["s"][:2]), []
            for _ in range(int(rec["n_sims"])):
                want.append(g.consumed)
                g.standard_normal(steps)
            f = int(rec["first_sim"])
            assert np.array_equal(plan.sim_start[f:f + len(want)], np.array(want, dtype=np.uint64))
            assert int(plan.stream_end[k]) == g.consumed


def test_hostemu_folded_philox_blocks_equal_numpy_raw_draws(emu):
    """Rounds 0 and 1 with their launch-constant halves folded (18 wide multiplications instead of 20) give NumPy's
    Philox4x64-10 output bit for bit -- fresh streams, spawned keys, and a generator advanced far into its stream."""
    from mcframework_b200 import _streams as S

    gens = [np.random.Generator(np.random.Philox(5)), np.random.Generator(np.random.Philox(np.random.SeedSequence(42).spawn(3)[2])),
            np.random.Generator(np.random.Philox(key=np.array([2**64 - 1, 12345], dtype=np.uint64))),
            np.random.Generator(np.random.Philox(9).advance(2**70 + 12345))]
    for g in gens:
        _, rec = S.generator_record(g)
        assert int(rec["buf_pos"][0]) % 4 == 0
        b0 = int(rec["buf_pos"][0]) // 4  # NumPy increments the counter before each block: a fresh buffer starts at block 1
        rc, got = emu.philox_folded(rec, b0, 64)
        assert rc == 0
        assert np.array_equal(got, g.bit_generator.random_raw(256))
        rc, got = emu.philox_folded(rec, b0 + (1 << 40), 8)  # far ahead: compare with NumPy's own jump
        bg = type(g.bit_generator)()
        bg.state = g.bit_generator.state
        bg.advance((1 << 40) - 64)  # the generator above already consumed 64 blocks
        assert rc == 0 and np.array_equal(got, bg.random_raw(32))


def test_child_philox_keys_equal_numpy_generate_state():
    """``_streams.child_philox_keys`` evaluates SeedSequence's hash for all spawned children at once (NumPy
    bit_generator.pyx: mix_entropy / generate_state; reference call site core.py:637,659): bit-identical keys for small
    and 128-bit entropies, sequence entropies, nested spawn keys and children that do not start at index 0; shapes it
    does not cover return None (the caller then asks NumPy)."""
    from mcframework_b200 import _streams as S

    for ent in (0, 1, 42, 123, 2**32 + 5, 2**100 + 12345, 2**127 - 1, [1, 2, 3], [5], (2**40, 7), np.random.SeedSequence().entropy):
        for prefix in ((), (3,), (0, 17), (2**31, 4, 5, 6, 7)):
            ss = np.random.SeedSequence(ent, spawn_key=prefix)
            ss.spawn(3)
            children = ss.spawn(70)
            want = np.array([c.generate_state(2, np.uint64) for c in children])
            got = S.child_philox_keys(children)
            assert got is not None and np.array_equal(got, want), (ent, prefix)
    assert S.child_philox_keys(np.random.SeedSequence(1).spawn(3)) is None  # too few: not worth it
    mixed = np.random.SeedSequence(1).spawn(8) + np.random.SeedSequence(2).spawn(8)
    assert S.child_philox_keys(mixed) is None  # children of different parents
    assert S.child_philox_keys(np.random.SeedSequence(1, pool_size=8).spawn(9)) is None
    blocks = [(i * 10, i * 10 + 10) for i in range(40)]
    kids = np.random.SeedSequence(9).spawn(40)
    rec = S.philox_records(kids, blocks)
    assert np.array_equal(rec["s"][:, :2], np.array([c.generate_state(2, np.uint64) for c in kids]))
    assert rec["first_sim"].tolist() == [a for a, _ in blocks] and rec["n_sims"].tolist() == [10] * 40


def test_result_copies_beyond_the_pinned_budget_are_pageable(monkeypatch):
    """Results the caller keeps stay page-locked only up to core.PINNED_RESULT_BUDGET; past it the copy goes to pageable
    memory (here: budget 0, so a large vector must not ask for pinned memory -- which a CPU-only torch could not give)."""
    import torch

    from mcframework_b200 import core

    monkeypatch.setattr(core, "PINNED_RESULT_BUDGET", 0)
    before = core._PINNED_LIVE[0]
    t = torch.arange(float(1 << 20), dtype=torch.float64)
    out = core.MonteCarloSimulation._to_host(t)
    assert isinstance(out, np.ndarray) and out.shape == (1 << 20,) and out[-1] == float((1 << 20) - 1)
    assert core._PINNED_LIVE[0] == before


def test_bootstrap_batch_size_follows_free_memory(monkeypatch):
    """The binned bootstrap bins 64 resamples per batch when its two queue sets fit a third of the free device memory,
    else 32, else 16; never more than there are resamples."""
    import torch

    from mcframework_b200 import _stats_device as SD

    class Lib:
        @staticmethod
        def mcf_bootstrap_binned_workspace_bytes(n, rpb, shift):
            return 8 * n * (rpb + 2)  # two sets of 4-byte queues, window = rpb + 2

    class X:
        is_cuda, device = True, "cuda:0"

    n = 100_000_000
    for free, B, want in [(180e9, 10_000, 64), (100e9, 10_000, 32), (60e9, 10_000, 16), (180e9, 40, 40), (180e9, 5, 5), (1e9, 5, 5)]:
        monkeypatch.setattr(torch.cuda, "mem_get_info", lambda dev=None, free=free: (int(free), int(180e9)))
        assert SD._default_resamples_per_batch(Lib, X, n, B, 0) == want, (free, B)
