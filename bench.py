#!/usr/bin/env python
"""Headline benchmark: Black-Scholes European call, N paths x 252 steps, payoffs + Greeks (BASELINE cfg 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--paths P] [--impl b200|reference]

One *step* = the job a reference user runs for "terminal payoff + Greeks" on one batch of paths:
  (1) sim.run(P, parallel=True, n_workers=8, compute_stats=False)      -> P discounted payoffs
  (2) sim.calculate_greeks(P, parallel=True)                           -> price, delta, gamma, vega, theta, rho
Units per step = P * n_steps GBM path-steps (the batch priced; the bumped revaluations of (2) re-use the
same normals and are NOT counted again), for both arms.

  value : device-resident (stream records already in HBM, nothing copied back): plan + replay kernels.
  e2e   : the same job through the public API, host buffers: stream records H2D, result vector D2H.
  --impl reference : the oracle's NumPy port of the reference's own CPU path (parallel=True on all host
                     cores), on a bounded sample of the same workload.

N > 1 (torchrun, one rank per GPU): weak scaling -- every rank replays P paths (its share of the
8N-worker block layout of an N*P-path run); no data-path collective; time = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BS = dict(S0=100.0, K=100.0, T=1.0, r=0.05, sigma=0.20)
N_STEPS = 252
WORKERS_PER_GPU = 8  # reference block layout: 8 workers x 8 chunks = 64 Philox streams per GPU share


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--paths", type=float, default=1e8, help="paths per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-paths", type=int, default=20_000, help="bounded sample for the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi samples during the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        threading.Thread(target=self._pump, daemon=True).start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                power.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, cell in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if cell.lower().startswith("active"):
                    reasons.add(name)
        busy = [c for c, p in zip(sm, power) if p > 300.0] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------ reference arm
def cpu_job(n_paths: int, n_workers: int):
    """The reference's own CPU path for one step, restated by oracle/np_port.py (kind "port")."""
    from oracle import np_port

    run = np_port.Runner("black_scholes", seed=42)
    payoffs = run.run(n_paths, parallel=True, n_workers=n_workers, n_steps=N_STEPS, **BS)
    greeks = np_port.greeks(n_paths, parallel=True, n_workers=n_workers, n_steps=N_STEPS, **BS)
    return payoffs, greeks


def time_cpu(n_paths: int, reps: int, warm: int):
    cores = os.cpu_count() or 1
    for _ in range(warm):
        cpu_job(max(20_000, n_paths // 4), cores)
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        cpu_job(n_paths, cores)
        ts.append(time.perf_counter() - t)
    return ts, cores


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = max(20_000, args.cpu_paths)  # >= _PARALLEL_THRESHOLD so the reference really takes its parallel path
    ts, cores = time_cpu(n, max(1, args.steps), min(1, args.warmup))
    per = sum(ts) / len(ts)
    val = n * N_STEPS / per
    sample = f"{n} paths x {N_STEPS} steps per step: run(parallel=True) + calculate_greeks (8 revaluations), {cores} threads"
    line = {
        "impl": "reference", "metric": "gbm_path_steps_per_s", "value": val, "unit": "path-steps/s", "n_gpus": args.gpus,
        "steps": len(ts), "warmup": min(1, args.warmup), "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "black_scholes_european_call_payoffs_plus_greeks", "paths_per_step": n, "n_steps": N_STEPS,
                   "seed": 42, **BS},
        "cpu_baseline": {"value": val, "unit": "path-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "path-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ B200 arm
def b200_arm(args):
    import numpy as np
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist  # noqa: PLC0415

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import mcframework_b200 as mcf
    from mcframework_b200 import _device as D, _native as N, _stats_device as SD, _streams as S

    N.lib()
    P = int(args.paths)
    total_paths = P * world
    n_workers = WORKERS_PER_GPU * world
    greek_sets = None

    def layout():
        """This rank's share of the reference block layout for a (world*P)-path parallel run, seed 42."""
        ss = np.random.SeedSequence(42)
        full = S.parallel_layout(ss, total_paths, n_workers)
        return full.shard(rank, world) if world > 1 else full

    def bump_sets():
        S0, K, T, r, sigma = BS["S0"], BS["K"], BS["T"], BS["r"], BS["sigma"]
        dS, dsig, dT, dr = S0 * 0.01, sigma * 0.01, 1.0 / 365.0, r * 0.01
        return [[S0, K, T, r, sigma], [S0 + dS, K, T, r, sigma], [S0 - dS, K, T, r, sigma], [S0, K, T, r, sigma + dsig],
                [S0, K, T, r, sigma - dsig], [S0, K, T, r + dr, sigma], [S0, K, T, r - dr, sigma], [S0, K, T - dT, r, sigma]]

    greek_sets = [(N.BS_EURO_CALL, p) for p in bump_sets()]
    dl = D.DeviceLayout.of(layout())  # stream records resident in HBM before the timed region
    phases = {"plan": [], "replay1": [], "replay8": [], "moments": [], "draws": 0}

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def device_step(record: bool):
        marks = [ev()]
        plan = D.normal_plan(dl, N_STEPS)
        marks.append(ev())
        payoffs = D.gbm_terminal(dl, N_STEPS, plan, greek_sets[:1])
        marks.append(ev())
        del plan
        plan = D.normal_plan(dl, N_STEPS)  # calculate_greeks re-seeds and re-plans (no caching across calls)
        marks.append(ev())
        vals = D.gbm_terminal(dl, N_STEPS, plan, greek_sets)
        marks.append(ev())
        sums = [SD.moments_raw(vals[k]) for k in range(len(greek_sets))]
        marks.append(ev())
        if record:
            torch.cuda.synchronize()
            phases["draws"] = plan.stream_end  # 64-bit draws the replications consume (exact); summed after the timing
            d = [marks[i].elapsed_time(marks[i + 1]) for i in range(5)]
            phases["plan"] += [d[0], d[2]]
            phases["replay1"].append(d[1])
            phases["replay8"].append(d[3])
            phases["moments"].append(d[4])
        return payoffs, sums

    sim = mcf.BlackScholesSimulation()
    fw = mcf.MonteCarloFramework()
    fw.register_simulation(sim)
    if world > 1:
        # "root": rank 0 receives every shard through NCCL and copies the whole vector out; "shm" (opt-in, A/B): every
        # rank copies its own shard into one shared-memory buffer (mcframework_b200/_multi.py)
        sim.gather_mode = os.environ.get("MCF_BENCH_GATHER", "root")

    def e2e_step():
        sim.set_seed(42)
        res = fw.run_simulation(sim.name, total_paths, parallel=True, n_workers=n_workers, compute_stats=False, n_steps=N_STEPS, **BS)
        greeks = sim.calculate_greeks(total_paths, parallel=True, n_steps=N_STEPS, n_workers=n_workers, **BS)
        return res, greeks

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        out = None
        for _ in range(k):
            out = fn()
        b.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = a.elapsed_time(b)
        if dist is not None:
            t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms, wall, out

    import logging

    logging.getLogger("mcframework_b200.core").setLevel(logging.WARNING)
    for _ in range(args.warmup):
        device_step(False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    before = dict(D.LAUNCHES)
    ms_dev, _, (payoffs, sums) = timed(lambda: device_step(True), args.steps)
    launches = sum(D.LAUNCHES.values()) - sum(before.values())
    clocks = sampler.stop() if rank == 0 else None

    for _ in range(args.warmup):  # also lets the pinned result buffers (two alternate) reach steady state
        e2e_step()
    ms_e2e_dev, wall_e2e, (res, greeks) = timed(e2e_step, args.steps)
    ms_e2e = max(ms_e2e_dev, wall_e2e * 1e3)  # host copies / host work are part of the end-to-end time

    units = total_paths * N_STEPS
    value = units / (ms_dev / args.steps / 1e3)
    e2e_val = units / (ms_e2e / args.steps / 1e3)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- sanity of what was timed (cheap; outside the timed regions)
    price = float(sums[0].cpu().numpy()[4])
    checks = {"price_device": price, "price_e2e": greeks["price"], "bs_analytic_call": 10.450583572185565,
              "delta_e2e": greeks["delta"], "results_len": int(res.results.size)}

    # ---- roofline of the dominant kernel group (measured live, CUDA events on the launch stream)
    draws = int(phases.pop("draws").sum().item())
    if os.environ.get("MCF_BENCH_DEBUG"):
        print({k: [round(x, 2) for x in v] for k, v in phases.items()}, file=sys.stderr)
    mean_ms = {k: (sum(v) / len(v) if v else 0.0) for k, v in phases.items()}
    per_step_ms = {"plan": 2 * mean_ms["plan"], "replay1": mean_ms["replay1"], "replay8": mean_ms["replay8"], "moments": mean_ms["moments"]}
    dominant = max(per_step_ms, key=per_step_ms.get)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    draws_per_gpu = float(draws) if draws else P * N_STEPS * 1.0157  # the ziggurat consumes 1.57 % extra 64-bit draws
    n_chunks = draws_per_gpu * 1.03 / 64
    alg_bytes = {  # algorithmic HBM bytes per launch group (DESIGN.md "bytes per unit")
        "plan": n_chunks * (8 + 8 + 1 + 1 + 1 + 32) + P * 8,  # masks, exit, gen, entry, 4 sub-chunk sums + sim_start
        "replay1": P * (8 + 8) + n_chunks * 32,               # sim_start, payoff, the stored sub-chunk sums read once
        "replay8": P * (8 + 8 * 8) + n_chunks * 32,
        "moments": 8 * P * 8,
    }
    # The planning pass is bound by the integer-multiply pipe: Philox4x64-10 is 18 64x64->128 multiplications per block
    # of 4 draws (20 minus the two with launch-constant operands) = 72 IMAD.WIDE.U32 per block and thread, and the
    # measured issue rate of that instruction (tools/probe_pipes.cu on this GPU model, profiles/r1_probe_pipes.json) is
    # ~0.30 per clock per SM sub-partition.  ALGORITHMIC work = blocks that hold draws the replications consume (the
    # look-ahead block per chunk, the 3 % window slack and the re-walks are overhead, not work).
    ri = {}
    try:
        with open(os.path.join(ROOT, "profiles", "r1_roofline_inputs.json")) as f:
            ri = json.load(f)
    except OSError:
        pass
    dom_ms = mean_ms["plan"] if dominant == "plan" else per_step_ms[dominant]
    sm_mhz = (clocks or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
    roof = {"kernel": dominant, "bound": "imad" if dominant != "moments" else "hbm"}
    hbm_ach = alg_bytes[dominant] / (dom_ms / 1e3) / 1e9 if dom_ms else None
    if roof["bound"] == "hbm":
        roof.update({"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak, "traffic": None,
                     "peak_source": peak_src})
    else:
        wide_peak = float(ri.get("imad_wide_warp_inst_per_s", 3.4755e11)) * (sm_mhz / float(ri.get("probe_sm_mhz", 1965.0)))
        blocks = draws / 4.0 if dominant == "plan" else P * 2.0  # replay: at most two blocks per replication boundary
        ach = blocks * 72.0 / 32.0 / (dom_ms / 1e3) if dom_ms else None
        roof.update({"achieved": ach / 1e9 if ach else None, "peak": wide_peak / 1e9, "unit": "Gwarp-IMAD.WIDE/s",
                     "frac": (ach / wide_peak) if ach else None,
                     "traffic": ri.get(dominant + "_dram_bytes_per_launch_1e7_paths"),
                     "algorithmic_work": "Philox blocks holding consumed draws x 72 IMAD.WIDE.U32 per block and thread",
                     "peak_source": "measured IMAD.WIDE.U32 issue rate (tools/probe_pipes.cu, profiles/r1_probe_pipes.json), "
                                    f"scaled to {sm_mhz:.0f} MHz",
                     "ncu_pipe_busy": ri.get(dominant + "_fmaheavy_busy_frac"),
                     "ncu_issue_active": ri.get(dominant + "_issue_active_frac"),
                     "ncu_source": ri.get("source"),
                     "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak if hbm_ach else None,
                             "algorithmic_bytes_per_launch": alg_bytes[dominant], "peak_source": peak_src,
                             "note": "register-resident kernel: HBM is not the binding resource"}})
    roof["ms_per_step_by_kernel_group"] = per_step_ms

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        ts, cores = time_cpu(max(20_000, args.cpu_paths), 1, 1)
        cpu = {"value": max(20_000, args.cpu_paths) * N_STEPS / ts[0], "unit": "path-steps/s", "cores": cores, "kind": "port",
               "sample": f"{max(20_000, args.cpu_paths)} paths x {N_STEPS} steps, one step (run + calculate_greeks), "
                         f"parallel=True on {cores} threads, {ts[0]:.1f} s"}

    rec_bytes = 64 * 80
    line = {
        "metric": "gbm_path_steps_per_s", "value": value, "unit": "path-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "black_scholes_european_call_payoffs_plus_greeks", "paths_per_gpu": P, "paths_total": total_paths,
                   "n_steps": N_STEPS, "seed": 42, "stream_layout": f"reference parallel blocks, n_workers={n_workers}",
                   "normals": "NumPy ziggurat on the reference Philox streams (stream-exact)",
                   "l2_policy": "working set per step (plan workspace + outputs, >8 GB) exceeds the 126 MB L2", **BS},
        "e2e": {"value": e2e_val, "unit": "path-steps/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": 2 * rec_bytes * world, "d2h_bytes_per_step": total_paths * 8 + 8 * MOMENT_BYTES},
        "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "clocks": clocks, "checks": checks,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


MOMENT_BYTES = 12 * 8

if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        b200_arm(a)
