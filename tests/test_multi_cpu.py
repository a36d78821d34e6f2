"""World-size-2 gloo tests (CPU): the host side of the N>1 path -- block sharding, result gathering, and the
combination of per-rank statistical partials.  Kernels are not involved (those are covered by ``-m gpu``)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _summary_of(a):
    from mcframework_b200._stats_device import MomentSummary

    m = a.mean() if a.size else 0.0
    return MomentSummary(a.size, 0, 0, 0, m, ((a - m) ** 2).sum(), ((a - m) ** 3).sum(), ((a - m) ** 4).sum(),
                         ((a - 2.0) ** 2).sum(), a.sum())


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mcframework_b200 import _multi, _streams as S
        from mcframework_b200._stats_device import MomentSummary, merge_summaries, quantile_plan

        # (1) every rank derives the SAME block layout from the seed and takes a contiguous range of blocks
        n = 20_011
        full = S.parallel_layout(np.random.SeedSequence(7), n, 8)
        mine = full.shard(rank, world)
        lo = full.n_streams * rank // world
        assert np.array_equal(mine.records["s"], full.records["s"][lo:lo + mine.n_streams])
        assert int(mine.records["first_sim"][0]) == 0 and mine.n_total == int(mine.records["n_sims"].sum())

        # (2) results come back in reference order on every rank (and on the root only)
        first = int(full.records["first_sim"][lo])
        local = torch.arange(first, first + mine.n_total, dtype=torch.float64)
        everyone = _multi.gather_results(local, dist.group.WORLD)
        assert torch.equal(everyone, torch.arange(n, dtype=torch.float64))
        root = _multi.gather_results_to_root(local, dist.group.WORLD)
        assert (root is None) == (rank != 0)
        if rank == 0:
            assert torch.equal(root, torch.arange(n, dtype=torch.float64))

        # (2b) shared-memory gather: every rank writes its own shard into rank 0's segment; the segment is recycled once
        #      the result array is gone, and never while somebody still holds it
        import gc

        shared = _multi.gather_results_shared(local, dist.group.WORLD)
        assert (shared is None) == (rank != 0)
        if rank == 0:
            assert shared.dtype == np.float64 and np.array_equal(shared, np.arange(n, dtype=np.float64))
            pool = _multi._pool()
            assert len(pool._all) == 1 and not pool._free
        del shared
        gc.collect()
        again = _multi.gather_results_shared(local * 2.0, dist.group.WORLD)
        if rank == 0:
            assert len(pool._all) == 1, "the released segment must be reused"
            assert np.array_equal(again, 2.0 * np.arange(n))
        third = _multi.gather_results_shared(local + 1.0, dist.group.WORLD)  # `again` is still alive on rank 0
        if rank == 0:
            assert len(pool._all) == 2 and np.array_equal(again, 2.0 * np.arange(n)) and np.array_equal(third, np.arange(n) + 1.0)
        else:
            assert again is None and third is None and len(_multi._pool()._attached) == 2
        view = third[5:9] if rank == 0 else None
        del third, again
        gc.collect()
        if rank == 0:
            assert len(pool._free) == 1, "a live view keeps its segment out of the pool"
            assert view.tolist() == [6.0, 7.0, 8.0, 9.0]
        del view

        # (3) moment partials: all-gather 12 doubles per rank, merge with the pairwise update formulas
        x = np.random.default_rng(0).lognormal(0, 1, n)
        part = torch.from_numpy(_summary_of(x[first:first + mine.n_total]).to_array())
        parts = [torch.empty_like(part) for _ in range(world)]
        dist.all_gather(parts, part)
        merged = merge_summaries(MomentSummary.from_array(p.numpy()) for p in parts)
        whole = _summary_of(x)
        for f in ("n", "mean", "m2", "m3", "m4", "sq_target", "total"):
            np.testing.assert_allclose(getattr(merged, f), getattr(whole, f), rtol=1e-11)

        # (4) radix-select combine: summed per-rank digit histograms select the global order statistic
        keys = np.sort(x).view(np.uint64)  # positive doubles: bit patterns are already order-preserving
        ranks, _, _ = quantile_plan(n, [50])
        want = np.sort(x)[ranks[0]]
        mykeys = x[first:first + mine.n_total].view(np.uint64)
        prefix, rem = np.uint64(0), int(ranks[0])
        for p in range(8):
            shift = np.uint64(64 - 8 * (p + 1))
            sel = mykeys if p == 0 else mykeys[(mykeys >> (shift + np.uint64(8))) == prefix]
            hist = torch.from_numpy(np.bincount(((sel >> shift) & np.uint64(255)).astype(np.int64), minlength=256))
            dist.all_reduce(hist)
            c = np.cumsum(hist.numpy())
            d = int(np.searchsorted(c, rem, side="right"))
            rem -= int(c[d - 1]) if d else 0
            prefix = (prefix << np.uint64(8)) | np.uint64(d)
        assert np.array([prefix], dtype=np.uint64).view(np.float64)[0] == want and keys.size == n

        # (5) bootstrap slot-range sharding: ordinal bases from an all-gather of accepted counts
        from oracle import c_oracle as co

        npop, B = 1000, 7
        g = co.Gen.pcg64(5)
        slots = g.raw32(2 * 6000)
        thr = (1 << 32) % npop
        acc = ((slots.astype(np.uint64) * np.uint64(npop)) & np.uint64(0xFFFFFFFF)) >= thr
        total = slots.size
        a, b = total * rank // world, total * (rank + 1) // world
        cnt = torch.tensor([int(acc[a:b].sum())])
        cnts = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(cnts, cnt)
        base = sum(int(c) for c in cnts[:rank])
        idx = ((slots[a:b].astype(np.uint64) * np.uint64(npop)) >> np.uint64(32))[acc[a:b]]
        ords = base + np.arange(idx.size)
        keep = ords < B * npop
        xs = np.random.default_rng(1).integers(0, 1000, npop).astype(float)
        sums = torch.zeros(B, dtype=torch.float64)
        sums.index_add_(0, torch.from_numpy(ords[keep] // npop), torch.from_numpy(xs[idx[keep].astype(np.int64)]))
        dist.all_reduce(sums)
        ref = np.random.default_rng(5)
        want = np.array([xs[ref.integers(0, npop, size=npop)].sum() for _ in range(B)])
        assert np.array_equal(sums.numpy(), want)
        q.put((rank, "ok"))
    except Exception as exc:  # noqa: BLE001
        import traceback

        q.put((rank, traceback.format_exc() + repr(exc)))
    finally:
        dist.destroy_process_group()


def test_two_rank_host_logic_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(60)
    assert out == {0: "ok", 1: "ok"}, out


def _run_worker(rank, world, port, q):
    """Block-parallel public-API runs on 2 ranks (stand-in device of tests/refsuite, gloo): every gather mode returns the
    reference-ordered vector where it should, equal to the single-process run."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import logging
        import sys

        here = os.path.dirname(os.path.abspath(__file__))
        for p_ in (os.path.dirname(here), here):
            if p_ not in sys.path:
                sys.path.insert(0, p_)
        from refsuite import fakedev

        fakedev.install()
        import mcframework_b200 as m
        from oracle import np_port

        logging.getLogger("mcframework_b200.core").setLevel(logging.WARNING)
        n, workers = 20_003, 3  # 24 blocks of 833 + a ragged one: 12 / 13 blocks per rank
        want = np_port.Runner("pi", seed=9).run(n, parallel=True, n_workers=workers, n_points=40)
        first = {0: 0, 1: 12 * 833}[rank]
        for mode in ("all", "root", "shm"):
            sim = m.PiEstimationSimulation()
            sim.shard_group = "world"  # sharding is opt-in
            sim.gather_mode = mode
            sim.set_seed(9)
            got = sim.run(n, parallel=True, n_workers=workers, n_points=40, compute_stats=False).results
            if mode == "all" or rank == 0:
                assert got.shape == (n,) and np.array_equal(got, want), (mode, rank)
            else:  # the other ranks keep their own shard
                assert np.array_equal(got, want[first:first + got.size]) and got.size in (12 * 833, n - 12 * 833), (mode, rank)
            assert sim.seed_seq.n_children_spawned == 25
        # sharded statistics: moments, percentiles, z/t interval, seeded BCa bootstrap, distribution-free bounds --
        # every rank reports what the single-process run (the oracle's port of the reference engine) reports
        ctx_kw = dict(target=3.1, eps=0.01, rng=4, n_bootstrap=100, bootstrap="bca")
        sim = m.PiEstimationSimulation()
        sim.shard_group = dist.group.WORLD
        sim.set_seed(9)
        res = sim.run(n, parallel=True, n_workers=workers, n_points=40, percentiles=(1, 99), extra_context=ctx_kw)
        # without the opt-in an initialised process group is ignored: every process runs the whole job
        solo = m.PiEstimationSimulation()
        solo.set_seed(9)
        assert np.array_equal(solo.run(n, parallel=True, n_workers=workers, n_points=40, compute_stats=False).results, want)
        # more ranks than blocks cannot happen with 2 ranks; an EMPTY shard is exercised through the layout instead
        from mcframework_b200 import _streams as S_
        assert S_.parallel_layout(np.random.SeedSequence(1), 20_000, 1).shard(7, 16).n_total >= 0
        ref = np_port.default_metrics(want, n=n, percentiles=(5, 25, 50, 75, 95), **ctx_kw)
        for key in ("mean", "std", "skew", "kurtosis", "bias_to_target", "mse_to_target", "markov_error_prob",
                    "chebyshev_required_n"):
            np.testing.assert_allclose(res.stats[key], ref[key], rtol=1e-9, err_msg=key)
        for key in ("ci_mean", "ci_mean_bootstrap", "ci_mean_chebyshev"):
            for f in ("low", "high"):
                np.testing.assert_allclose(res.stats[key][f], ref[key][f], rtol=1e-9, err_msg=key)
        assert res.percentiles == {int(p): float(np.percentile(want, p)) for p in (1, 5, 25, 50, 75, 95, 99)}
        assert (res.mean, res.n_simulations) == (res.stats["mean"], n)
        q.put((rank, "ok"))
    except Exception as exc:  # noqa: BLE001
        import traceback

        q.put((rank, traceback.format_exc() + repr(exc)))
    finally:
        dist.destroy_process_group()


def test_two_rank_public_api_runs_and_gather_modes_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(60)
    assert out == {0: "ok", 1: "ok"}, out
