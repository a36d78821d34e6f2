#!/usr/bin/env python
"""N-rank parity check of the sharded paths (one process per GPU, NCCL).  Launch with

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py

Every rank runs the SAME public-API calls (SPMD); results must equal the single-GPU / reference fixtures at every
shard count: replications shard by spawned block, statistics combine partials with small collectives.
"""
import json
import logging
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import mcframework_b200 as m
    from mcframework_b200 import _stats_device as SD
    from mcframework_b200.stats_engine import DeviceSample, build_default_engine

    logging.getLogger("mcframework_b200.core").setLevel(logging.WARNING)
    a = dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz")))
    j = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_golden.json")))
    ok = []

    def check(name, cond):
        ok.append(bool(cond))
        if rank == 0:
            print(("PASS " if cond else "FAIL ") + name, flush=True)

    # replications sharded by block: full result vector on every rank, in reference order
    sim = m.PiEstimationSimulation()
    sim.shard_group = "world"  # sharding is opt-in: every rank makes the same calls
    sim.set_seed(123)
    r = sim.run(20_000, n_points=100, parallel=True, n_workers=8, compute_stats=False)
    check("pi blocks sharded == reference (bit-exact)", np.array_equal(r.results, a["G2_pi_parallel_results"]))
    bs = m.BlackScholesSimulation()
    bs.shard_group = "world"
    bs.set_seed(11)
    r = bs.run(20_000, parallel=True, n_workers=4, n_steps=64)
    check("black-scholes blocks sharded == reference", np.allclose(r.results, a["G4_bs_call_par"], rtol=1e-12, atol=1e-11))
    ref = a["G4_bs_call_par"]
    good = (abs(r.mean - ref.mean()) <= 1e-12 * ref.mean() and abs(r.std - ref.std(ddof=1)) <= 1e-11 * ref.std(ddof=1)
            and abs(r.percentiles[50] - np.percentile(r.results, 50)) <= 1e-12 and r.percentiles[95] == np.percentile(r.results, 95))
    if not good and rank == 0:
        print("   detail:", r.mean, ref.mean(), r.std, ref.std(ddof=1), r.percentiles, np.percentile(r.results, [50, 95]))
    check("sharded stats: mean/std/percentiles", good)
    # the large-run path at fixture size: every rank pipelines its shard into ONE shared-memory result buffer
    shm = m.BlackScholesSimulation()
    shm.shard_group, shm.gather_mode, shm._PIPELINE_MIN_SIMS = "world", "shm", 1000
    shm.set_seed(11)
    rs = shm.run(20_000, parallel=True, n_workers=4, n_steps=64)
    if rank == 0:
        good = rs.results.shape == (20_000,) and np.allclose(rs.results, a["G4_bs_call_par"], rtol=1e-12, atol=1e-11)
    else:  # the other ranks see their own shard (a view into the shared buffer)
        lo_blk = 33 * rank // world  # 20 000 // (4*8) = 625-sized blocks: 32 blocks (no ragged one)
        good = rs.results.size > 0 and np.allclose(rs.results, a["G4_bs_call_par"][625 * (32 * rank // world):][:rs.results.size],
                                                   rtol=1e-12, atol=1e-11)
        _ = lo_blk
    good = good and abs(rs.mean - ref.mean()) <= 1e-12 * ref.mean()
    check("shared-memory pipelined gather (gather_mode='shm') == reference, sharded stats", good)
    g = bs.calculate_greeks(20_000, option_type="put", n_steps=12, parallel=True, n_workers=j["G4_greeks_put_par_cpu_count"])
    check("greeks sharded == reference", all(abs(g[k] - w) <= 1e-9 * abs(w) for k, w in j["G4_greeks_put_par"].items()))

    # statistics of a sharded sample (each rank holds a slice): engine output == reference G7
    x = np.random.default_rng(0).normal(5, 2, 100_000)
    lo, hi = x.size * rank // world, x.size * (rank + 1) // world
    sample = DeviceSample(torch.from_numpy(x[lo:hi].copy()).cuda(), None, dist.group.WORLD)
    ctx = m.StatsContext(n=x.size, percentiles=(1, 5, 50, 95, 99), rng=7, n_bootstrap=1000, bootstrap="bca", target=5.0, eps=0.01)
    out = build_default_engine().compute(sample, ctx).metrics
    want = j["G7"]
    good = all(abs(out[k] - want[k]) <= 1e-10 * abs(want[k]) + 1e-13 for k in ("mean", "std", "skew", "kurtosis", "mse_to_target"))
    good &= out["percentiles"] == {int(k): v for k, v in want["percentiles"].items()}
    good &= all(abs(out["ci_mean_bootstrap"][k] - want["ci_mean_bootstrap"][k]) < 1e-11 * 5 for k in ("low", "high"))
    check("sharded 12-metric engine (moments all-gather, select all-reduce, slot-sharded bootstrap) == reference", good)

    # bootstrap sharded by slot range: identical index stream, generator advanced identically
    xi = np.random.default_rng(2).integers(0, 1 << 20, 50_000).astype(float)
    g_ref, g_dev = np.random.default_rng(5), np.random.default_rng(5)
    want_sums = np.stack([xi[g_ref.integers(0, xi.size, size=xi.size)].sum() for _ in range(40)])
    got = SD.bootstrap_means(torch.from_numpy(xi).cuda(), 40, g_dev, group=dist.group.WORLD).cpu().numpy() * xi.size
    check("slot-sharded bootstrap: integer sums exact, generator state equal",
          np.array_equal(np.rint(got), want_sums) and g_dev.bit_generator.state == g_ref.bit_generator.state)

    # the same through the binned gather and its single-pass form (safe CTAs bin without exact ordinals)
    xb = np.random.default_rng(3).integers(0, 1 << 20, 400_000).astype(float)
    g_ref, g_dev = np.random.default_rng(6), np.random.default_rng(6)
    want_b = np.stack([xb[g_ref.integers(0, xb.size, size=xb.size)].sum() for _ in range(24)])
    before = SD.BOOT_FORMS["single_pass"]
    got_b = SD.bootstrap_means(torch.from_numpy(xb).cuda(), 24, g_dev, group=dist.group.WORLD, binned=True, region_shift=15,
                               chunk_slots=256, resamples_per_batch=5).cpu().numpy() * xb.size
    check("slot-sharded BINNED bootstrap, single-pass form: integer sums exact, generator state equal",
          np.array_equal(np.rint(got_b), want_b) and g_dev.bit_generator.state == g_ref.bit_generator.state
          and SD.BOOT_FORMS["single_pass"] == before + 1)

    # LSM with paths sharded by rank
    ps = m.BlackScholesPathSimulation()
    ps.set_seed(1)
    paths = ps.simulate_paths(4000, n_steps=50)
    lo, hi = 4000 * rank // world, 4000 * (rank + 1) // world
    tm = SD.transpose(torch.from_numpy(np.ascontiguousarray(paths[lo:hi])).cuda())
    price = float(SD.lsm_price(tm, 100.0, 0.05, 1.0 / 50, True, group=dist.group.WORLD).cpu()[0])
    check("LSM with sharded paths == reference (fused step kernel, sums exchanged over peer memory)",
          abs(price - j["G6"]["lsm_put"]) <= 1e-9 * j["G6"]["lsm_put"])
    used_peer = SD.LsmPeerExchange.get(dist.group.WORLD).ok
    staged = float(SD.lsm_price(tm, 100.0, 0.05, 1.0 / 50, True, group=dist.group.WORLD, fused=False).cpu()[0])
    again = float(SD.lsm_price(tm, 100.0, 0.05, 1.0 / 50, False, group=dist.group.WORLD).cpu()[0])  # a second sweep, a call
    check(f"LSM staged (NCCL all-reduce per step) == fused (peer exchange in use: {used_peer}); call price == reference",
          abs(staged - price) <= 1e-12 * price and abs(again - j["G6"]["lsm_call"]) <= 1e-9 * j["G6"]["lsm_call"])

    flag = torch.tensor([int(all(ok))], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"multi_gpu_check world={world}: {'ALL PASS' if int(flag) else 'FAILURES'} ({sum(ok)}/{len(ok)} on rank 0)", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
