#!/usr/bin/env python
"""Benchmark of the replication hot path on B200 (BASELINE.json: GBM path-steps/s, Pi samples/s, bootstrap resamples/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--paths P] [--scaling weak|strong]
                    [--configs all|none|pi,portfolio_slices,lsm,stats] [--impl b200|reference]

HEADLINE (top-level keys) = BASELINE config 2: Black-Scholes European call, P paths x 252 steps, payoffs + Greeks.
One *step* = the job a reference user runs for "terminal payoff + Greeks" on one batch of paths:
  (1) sim.run(P, parallel=True, n_workers=8, compute_stats=False)      -> P discounted payoffs
  (2) sim.calculate_greeks(P, parallel=True)                           -> price, delta, gamma, vega, theta, rho
Units per step = P * n_steps GBM path-steps (the batch priced; the bumped revaluations of (2) re-use the same normals and
are NOT counted again), for both arms.  Each call plans its streams itself: nothing is cached across calls or steps.

  value : device-resident (stream records already in HBM, nothing copied back): plan + replay kernels.
  e2e   : the same job through the public API, host buffers: stream records H2D, result vector D2H.
  --impl reference : the reference's own CPU path (parallel=True on all host cores) on a bounded sample of the same
                     workload -- the UNMODIFIED reference from baseline/_ref when that install exists (kind "reference"),
                     else the oracle's NumPy port (kind "port"); see bench_reference.py.

"configs" = the other BASELINE configs, one entry each with value / unit / roofline / cpu_baseline (N = 1 only):
  pi (cfg 1), portfolio_slices (cfg 3), lsm (cfg 4), stats (cfg 5: moments, percentiles, BCa bootstrap B = 10 000).

N > 1 (torchrun, one rank per GPU).  --scaling weak (default): every rank replays P paths, its share of the 8N-worker
block layout of an N*P-path run.  --scaling strong: the P-path, 8-worker job of N = 1 is sharded (64/N blocks per rank).
No data-path collective either way; time = max over ranks.  checks.sharded_parity: the reference fixtures G2 (Pi) and
G4 (Black-Scholes), run SHARDED over the N ranks through the public API, equal the single-process goldens.
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
MOMENT_BYTES = 12 * 8
ALL_CONFIGS = ("pi", "portfolio_slices", "lsm", "stats")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--paths", type=float, default=1e8, help="paths per GPU per step (weak) / in total (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--configs", default="all", help="all | none | comma list of " + ",".join(ALL_CONFIGS))
    ap.add_argument("--config-scale", type=float, default=1.0, help="shrinks the sizes of the 'configs' entries (development)")
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
def cpu_line(fn, *a, unit, scale=1.0):
    """One bounded CPU sample -> cpu_baseline object (value in `unit`; `scale` converts the sample's own units)."""
    import bench_reference as R

    dt, units, sample, check = fn(*a)
    kind = R.load()[0]
    single = fn.__name__ in ("lsm", "bootstrap", "percentiles") or fn.__name__ == "pi_as_is"
    return {"value": units * scale / dt, "unit": unit, "cores": 1 if single else R.cores(), "kind": kind,
            "sample": f"{sample}, {dt:.2f} s", "check": check}


def reference_arm(args):
    import bench_reference as R

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind = R.load()[0]
    cores = R.cores()
    n = max(20_000, args.cpu_paths)  # >= _PARALLEL_THRESHOLD so the reference really takes its parallel path
    for _ in range(min(1, args.warmup)):
        R.black_scholes_step(n, cores, N_STEPS)
    ts, check = [], None
    for _ in range(max(1, args.steps)):
        dt, _, _, check = R.black_scholes_step(n, cores, N_STEPS)
        ts.append(dt)
    per = sum(ts) / len(ts)
    val = n * N_STEPS / per
    sample = f"{n} paths x {N_STEPS} steps per step: run(parallel=True) + calculate_greeks (8 revaluations), {cores} threads"
    line = {
        "impl": "reference", "metric": "gbm_path_steps_per_s", "value": val, "unit": "path-steps/s", "n_gpus": args.gpus,
        "steps": len(ts), "warmup": min(1, args.warmup), "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "black_scholes_european_call_payoffs_plus_greeks", "paths_per_step": n, "n_steps": N_STEPS,
                   "seed": 42, **BS},
        "cpu_baseline": {"value": val, "unit": "path-steps/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "path-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "checks": check,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ B200 arm
class Ctx:
    """What every measurement needs: rank layout, peaks, timing helpers."""

    def __init__(self, args):
        import torch

        self.args = args
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        self.peaks = peaks
        self.hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        self.hbm_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
        ri = {}
        for name in ("r2_roofline_inputs.json", "r1_roofline_inputs.json"):
            try:
                with open(os.path.join(ROOT, "profiles", name)) as f:
                    ri = json.load(f)
                break
            except OSError:
                continue
        self.ri = ri
        self.sm_mhz = float(peaks.get("sm_max_mhz", 1965.0))
        self._imad_peak, self.imad_src = None, None

    @property
    def group(self):
        return self.dist.group.WORLD if self.dist is not None else None

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        if self.dist is None:
            return vals if len(vals) > 1 else vals[0]
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        out = [float(v) for v in t]
        return out if len(out) > 1 else out[0]

    def timed(self, fn, k):
        """k calls of fn bracketed by barrier + synchronize; -> (device ms, wall s, last result), max over ranks."""
        torch = self.torch
        self.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        out = None
        for _ in range(k):
            out = fn()
        b.record()
        self.barrier()
        wall = time.perf_counter() - t0
        ms, wall_ms = self.max_over_ranks(a.elapsed_time(b), wall * 1e3)
        return ms, wall_ms / 1e3, out

    def imad_peak(self):
        """Issue rate of IMAD.WIDE.U32 measured IN THIS RUN on this GPU (mcf_probe_issue_rate: the loop of
        tools/probe_pipes.cu); the committed round-1 probe (profiles/r1_probe_pipes.json) is only the fallback."""
        if self._imad_peak is None:
            import ctypes as C

            from mcframework_b200 import _native as N

            v = C.c_double(0.0)
            try:
                rc = N.lib().mcf_probe_issue_rate(0, C.byref(v), C.c_void_p(self.torch.cuda.current_stream().cuda_stream))
            except Exception:  # noqa: BLE001
                rc = -1
            if rc == 0 and v.value > 0:
                self._imad_peak, self.imad_src = float(v.value), "measured in this run (mcf_probe_issue_rate, 8 chains x 1024 threads per SM)"
            else:
                self._imad_peak = float(self.ri.get("imad_wide_warp_inst_per_s", 3.4755e11)) * (self.sm_mhz / float(self.ri.get("probe_sm_mhz", 1965.0)))
                self.imad_src = "committed probe profiles/r1_probe_pipes.json, scaled to the SM clock"
        return self._imad_peak

    def imad_roofline(self, philox_blocks, seconds, kernel):
        """Philox4x64-10 = 18 64x64->128 products per block of 4 draws after folding = 72 IMAD.WIDE.U32 per block."""
        ach = philox_blocks * 72.0 / 32.0 / seconds
        peak = self.imad_peak()
        return {"kernel": kernel, "bound": "imad", "achieved": ach / 1e9, "peak": peak / 1e9, "unit": "Gwarp-IMAD.WIDE/s",
                "frac": ach / peak, "traffic": None,
                "algorithmic_work": "Philox blocks holding consumed draws x 72 IMAD.WIDE.U32 per block and thread",
                "peak_source": "BUILDER-measured IMAD.WIDE.U32 issue rate, not a driver-measured peak (MEASURED_PEAKS.json has no "
                               "issue-slot figure): " + str(self.imad_src)}

    def hbm_roofline(self, nbytes, seconds, kernel, traffic=None):
        ach = nbytes / seconds / 1e9
        return {"kernel": kernel, "bound": "hbm", "achieved": ach, "peak": self.hbm_peak, "unit": "GB/s",
                "frac": ach / self.hbm_peak, "traffic": traffic, "algorithmic_bytes_per_launch_group": nbytes,
                "peak_source": self.hbm_src}


def ev(torch):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def sharded_parity(cx):
    """Reference fixtures G2 (Pi, bit-exact) and G4 (Black-Scholes, rtol 1e-12) run SHARDED over the ranks."""
    import numpy as np

    import mcframework_b200 as mcf

    a = dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz")))
    group = "world" if cx.world > 1 else None
    sim = mcf.PiEstimationSimulation()
    sim.shard_group = group
    sim.set_seed(123)
    r = sim.run(20_000, n_points=100, parallel=True, n_workers=8, compute_stats=False)
    ok_pi = bool(np.array_equal(r.results, a["G2_pi_parallel_results"]))
    bs = mcf.BlackScholesSimulation()
    bs.shard_group = group
    bs.set_seed(11)
    r = bs.run(20_000, parallel=True, n_workers=4, n_steps=64)
    ok_bs = bool(np.allclose(r.results, a["G4_bs_call_par"], rtol=1e-12, atol=1e-11))
    ref = a["G4_bs_call_par"]
    # statistics of the sharded sample: mean against the fixture, order statistics exact on the gathered vector
    ok_stats = bool(abs(r.mean - ref.mean()) <= 1e-12 * ref.mean() and abs(r.std - ref.std(ddof=1)) <= 1e-11 * ref.std(ddof=1)
                    and r.percentiles[95] == float(np.percentile(r.results, 95))) if ok_bs else False
    flags = cx.max_over_ranks(0.0 if ok_pi else 1.0, 0.0 if ok_bs else 1.0, 0.0 if ok_stats else 1.0)
    return {"sharded_parity": all(f == 0.0 for f in flags), "sharded_pi_g2_bit_exact": flags[0] == 0.0,
            "sharded_bs_g4_rtol_1e-12": flags[1] == 0.0, "sharded_stats_of_g4": flags[2] == 0.0, "shards": cx.world}


# ------------------------------------------------------------------------------------ headline (cfg 2)
def headline(cx):
    import numpy as np

    import mcframework_b200 as mcf
    from mcframework_b200 import _device as D, _native as N, _stats_device as SD, _streams as S

    args, torch, rank, world = cx.args, cx.torch, cx.rank, cx.world
    P = int(args.paths)
    strong = args.scaling == "strong"
    total_paths = P if strong else P * world
    n_workers = WORKERS_PER_GPU if strong else WORKERS_PER_GPU * world

    def layout():
        """This rank's share of the reference block layout of the whole job, seed 42."""
        full = S.parallel_layout(np.random.SeedSequence(42), total_paths, n_workers)
        return full.shard(rank, world) if world > 1 else full

    S0, K, T, r, sigma = BS["S0"], BS["K"], BS["T"], BS["r"], BS["sigma"]
    dS, dsig, dT, dr = S0 * 0.01, sigma * 0.01, 1.0 / 365.0, r * 0.01
    bumps = [[S0, K, T, r, sigma], [S0 + dS, K, T, r, sigma], [S0 - dS, K, T, r, sigma], [S0, K, T, r, sigma + dsig],
             [S0, K, T, r, sigma - dsig], [S0, K, T, r + dr, sigma], [S0, K, T, r - dr, sigma], [S0, K, T - dT, r, sigma]]
    greek_sets = [(N.BS_EURO_CALL, p) for p in bumps]
    lay = layout()
    dl = D.DeviceLayout.of(lay)  # stream records resident in HBM before the timed region
    my_paths = lay.n_total
    phases = {"plan": [], "replay1": [], "replay8": [], "moments": [], "draws": 0}

    # The device-resident step owns its HBM buffers (plan workspace, replication starts, outputs) across steps: what is
    # timed is the kernels, not PyTorch's caching allocator (an occasional 6 GB cudaMalloc inside one step cost 180 ms).
    # Every step still PLANS from scratch, twice -- the buffers hold no state between steps.
    bufs = {}
    out1 = torch.empty((1, my_paths), dtype=torch.float64, device="cuda")
    out8 = torch.empty((len(greek_sets), my_paths), dtype=torch.float64, device="cuda")

    def device_step(record: bool):
        marks = [ev(torch)]
        plan = D.normal_plan(dl, N_STEPS, buffers=bufs)
        marks.append(ev(torch))
        payoffs = D.gbm_terminal(dl, N_STEPS, plan, greek_sets[:1], out=out1)
        marks.append(ev(torch))
        del plan
        plan = D.normal_plan(dl, N_STEPS, buffers=bufs)  # calculate_greeks re-seeds and re-plans (no caching across calls)
        marks.append(ev(torch))
        vals = D.gbm_terminal(dl, N_STEPS, plan, greek_sets, out=out8)
        marks.append(ev(torch))
        sums = [SD.moments_raw(vals[k]) for k in range(len(greek_sets))]
        marks.append(ev(torch))
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
        sim.shard_group = cx.group
        # "shm": every rank copies its own shard into one shared-memory result buffer over its own PCIe link (default);
        # "root": rank 0 receives every shard through NCCL and copies the whole vector out (MCF_BENCH_GATHER=root, A/B)
        sim.gather_mode = os.environ.get("MCF_BENCH_GATHER", "shm")

    def e2e_step():
        sim.set_seed(42)
        res = fw.run_simulation(sim.name, total_paths, parallel=True, n_workers=n_workers, compute_stats=False, n_steps=N_STEPS, **BS)
        greeks = sim.calculate_greeks(total_paths, parallel=True, n_steps=N_STEPS, n_workers=n_workers, **BS)
        return res, greeks

    for _ in range(args.warmup):
        device_step(False)
    sampler = ClockSampler(cx.local)
    if rank == 0:
        sampler.start()
    before = dict(D.LAUNCHES)
    ms_dev, _, (payoffs, sums) = cx.timed(lambda: device_step(True), args.steps)
    launches = sum(D.LAUNCHES.values()) - sum(before.values())
    clocks = sampler.stop() if rank == 0 else None
    if clocks and clocks.get("sm_mhz"):
        cx.sm_mhz = float(clocks["sm_mhz"])

    for _ in range(args.warmup):  # also lets the pinned result buffers (two alternate) reach steady state
        e2e_step()
    ms_e2e_dev, wall_e2e, (res, greeks) = cx.timed(e2e_step, args.steps)
    ms_e2e = max(ms_e2e_dev, wall_e2e * 1e3)  # host copies / host work are part of the end-to-end time

    units = total_paths * N_STEPS
    value = units / (ms_dev / args.steps / 1e3)
    e2e_val = units / (ms_e2e / args.steps / 1e3)
    draws = int(phases.pop("draws").sum().item())
    if os.environ.get("MCF_BENCH_DEBUG"):
        print({k: [round(x, 2) for x in v] for k, v in phases.items()}, file=sys.stderr)
    parity = sharded_parity(cx)
    if rank != 0:
        return None

    price = float(sums[0].cpu().numpy()[4])
    checks = {"price_device": price, "price_e2e": greeks["price"], "bs_analytic_call": 10.450583572185565,
              "delta_e2e": greeks["delta"], "results_len": int(res.results.size), **parity}

    # ---- roofline of the dominant kernel group (measured live, CUDA events on the launch stream)
    mean_ms = {k: (sum(v) / len(v) if v else 0.0) for k, v in phases.items()}
    per_step_ms = {"plan": 2 * mean_ms["plan"], "replay1": mean_ms["replay1"], "replay8": mean_ms["replay8"], "moments": mean_ms["moments"]}
    dominant = max(per_step_ms, key=per_step_ms.get)
    n_chunks = draws * 1.001 / 64
    alg_bytes = {  # algorithmic HBM bytes per launch group (DESIGN.md "bytes per unit")
        "plan": n_chunks * (8 + 8 + 1 + 1 + 1 + 32) + my_paths * 8,  # masks, exit, gen, entry, 4 sub-chunk sums + sim_start
        "replay1": my_paths * (8 + 8) + n_chunks * 32,               # sim_start, payoff, the stored sub-chunk sums read once
        "replay8": my_paths * (8 + 8 * 8) + n_chunks * 32,
        "moments": 8 * my_paths * 8,
    }
    dom_ms = mean_ms["plan"] if dominant == "plan" else per_step_ms[dominant]
    if dominant == "moments":
        roof = cx.hbm_roofline(alg_bytes["moments"], dom_ms / 1e3, "moments_kernel")
    else:
        # ALGORITHMIC work = blocks that hold draws the replications consume (look-ahead blocks, window slack and
        # re-walks are overhead, not work); the replay regenerates at most two blocks per replication boundary
        blocks = draws / 4.0 if dominant == "plan" else my_paths * 2.0
        roof = cx.imad_roofline(blocks, dom_ms / 1e3, "plan_scan_philox_stream_kernel" if dominant == "plan" else "gbm_terminal_affine_stream_kernel")
        hbm_ach = alg_bytes[dominant] / (dom_ms / 1e3) / 1e9
        roof.update({"traffic": cx.ri.get(dominant + "_dram_bytes_per_launch_1e7_paths"),
                     "ncu_pipe_busy": cx.ri.get(dominant + "_fmaheavy_busy_frac"),
                     "ncu_issue_active": cx.ri.get(dominant + "_issue_active_frac"), "ncu_source": cx.ri.get("source"),
                     "hbm": {"achieved": hbm_ach, "peak": cx.hbm_peak, "unit": "GB/s", "frac": hbm_ach / cx.hbm_peak,
                             "algorithmic_bytes_per_launch": alg_bytes[dominant], "peak_source": cx.hbm_src,
                             "note": "register-resident kernel: HBM is not the binding resource"}})
    roof["ms_per_step_by_kernel_group"] = per_step_ms

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        import bench_reference as R

        n = max(20_000, args.cpu_paths)
        R.black_scholes_step(max(20_000, n // 4), R.cores(), N_STEPS)  # warm-up
        cpu = cpu_line(R.black_scholes_step, n, R.cores(), N_STEPS, unit="path-steps/s")

    return {
        "metric": "gbm_path_steps_per_s", "value": value, "unit": "path-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "black_scholes_european_call_payoffs_plus_greeks", "paths_per_gpu": my_paths, "paths_total": total_paths,
                   "n_steps": N_STEPS, "seed": 42, "stream_layout": f"reference parallel blocks, n_workers={n_workers}",
                   "normals": "NumPy ziggurat on the reference Philox streams (stream-exact)",
                   "l2_policy": "working set per step (plan workspace + outputs, >8 GB) exceeds the 126 MB L2",
                   "multi_gpu_result_gather": getattr(sim, "gather_mode", None) if world > 1 else None, **BS},
        "e2e": {"value": e2e_val, "unit": "path-steps/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": 2 * 64 * 80 * world, "d2h_bytes_per_step": total_paths * 8 + 8 * MOMENT_BYTES},
        "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "clocks": clocks, "checks": checks,
    }


# ------------------------------------------------------------------------------------ the other BASELINE configs
def best_of(cx, fn, reps=2, warm=1):
    """Best device time (CUDA events, max over ranks per repetition) of fn after `warm` untimed calls; -> (s, last result)."""
    for _ in range(warm):
        fn()
    best, out = 1e30, None
    for _ in range(reps):
        ms, _, out = cx.timed(fn, 1)
        best = min(best, ms * 1e-3)
    return best, out


def cfg_pi(cx):
    """cfg 1.  (a) as the README runs it (10 000 x 5000, seed 123, parallel=True = the sequential PCG64 stream), end to
    end through run_simulation; (b) the scaled run the SURVEY asks for throughput: 2e6 replications x 5000 points per GPU
    on the reference's Philox block layout, device resident."""
    import numpy as np

    import mcframework_b200 as mcf
    from mcframework_b200 import _device as D, _streams as S

    sc = cx.args.config_scale
    out = {}
    sim = mcf.PiEstimationSimulation()
    fw = mcf.MonteCarloFramework()
    fw.register_simulation(sim)

    def as_is(stats):
        sim.set_seed(123)
        return fw.run_simulation("Pi Estimation", 10_000, n_points=5000, parallel=True, compute_stats=stats)

    for stats in (True, False):
        as_is(stats)
        cx.torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = as_is(stats)
        wall = time.perf_counter() - t0
        out["as_is_e2e" + ("" if stats else "_no_stats")] = {
            "wall_s": wall, "samples_per_s": 5e7 / wall, "sum_hits": int(round(float(res.results.sum()) * 5000 / 4)),
            "includes": ("default 12-metric engine (bootstrap B=10000 over n=10000) + " if stats else "") + "80 KB D2H"}
    n = int(2e6 * sc) * cx.world
    lay = S.parallel_layout(np.random.SeedSequence(123), n, 8 * cx.world)
    if cx.world > 1:
        lay = lay.shard(cx.rank, cx.world)
    dl = D.DeviceLayout.of(lay)
    t, (hits, _) = best_of(cx, lambda: D.pi_hits(dl, 5000), reps=3)
    samples = n * 5000
    out.update({"metric": "pi_samples_per_s", "value": samples / t, "unit": "samples/s", "ms": t * 1e3, "scaling": "weak",
                "config": {"workload": "pi_estimation_scaled", "sims_total": n, "n_points": 5000, "seed": 123,
                           "stream_layout": f"reference parallel blocks, n_workers={8 * cx.world}", "resident": "device"},
                "roofline": cx.imad_roofline(lay.n_total * 5000 / 2.0, t, "pi_philox_stream_kernel"),  # 2 points per block
                "checks": {"as_is_sum_hits_bit_exact": out["as_is_e2e"]["sum_hits"] == 39_268_872,
                           "mean_estimate_rank0": float(hits.double().mean().item()) * 4 / 5000}})
    if cx.world == 1 and not cx.args.no_cpu_baseline:
        import bench_reference as R

        out["cpu_baseline"] = cpu_line(R.pi_as_is, unit="samples/s")
        out["cpu_baseline_parallel"] = cpu_line(R.pi_parallel, 40_000, 5000, unit="samples/s")
    return out


def cfg_portfolio(cx):
    """cfg 3: 1e7 paths x 2520 steps per GPU, wealth stored every 21 steps (121 slices, time-major): plan + slices."""
    import numpy as np

    from mcframework_b200 import _device as D, _native as N, _streams as S

    torch = cx.torch
    n, steps, stride = int(1e7 * cx.args.config_scale), 2520, 21
    lay = S.parallel_layout(np.random.SeedSequence(42), n * cx.world, 8 * cx.world)
    if cx.world > 1:
        lay = lay.shard(cx.rank, cx.world)
    dl = D.DeviceLayout.of(lay)
    params = [1e4, 0.07, 0.2]
    stages = {}

    def job(dtype, tm=None):
        plan = D.normal_plan(dl, steps)
        mid = ev(torch)
        sl = D.gbm_slices(dl, steps, plan, N.PORTFOLIO_GBM, params, stride, True, dtype=dtype, timings=tm)
        return plan, mid, sl

    job(torch.float64)
    res = {}
    for name, dtype in (("f64", torch.float64), ("f32", torch.float32)):
        tm = {}
        cx.barrier()
        a = ev(torch)
        plan, mid, sl = job(dtype, tm)
        b = ev(torch)
        cx.barrier()
        t_all, t_plan = cx.max_over_ranks(a.elapsed_time(b) * 1e-3, a.elapsed_time(mid) * 1e-3)
        st = {"relocate_ms": tm["start"].elapsed_time(tm["relocated"]), "segment_sums_ms": tm["relocated"].elapsed_time(tm["segment_sums"]),
              "store_ms": tm["segment_sums"].elapsed_time(tm["stored"])}
        nbytes = sl.numel() * sl.element_size()
        res[name] = {"s": t_all, "plan_s": t_plan, **st, "slice_bytes": nbytes,
                     "store_kernel_GBps": (nbytes + lay.n_total * (steps // stride) * 8) / (st["store_ms"] * 1e-3) / 1e9}
        draws = int(plan.stream_end.sum().item())
        last_mean = float(sl[-1].double().mean().item())
        del plan, sl
    units = n * cx.world * steps
    r = res["f64"]
    store_bytes = r["slice_bytes"] + lay.n_total * (steps // stride) * 8  # slice values written + segment sums read
    out = {"metric": "gbm_path_steps_per_s", "value": units / r["s"], "unit": "path-steps/s", "ms": r["s"] * 1e3, "scaling": "weak",
           "value_f32_slices": units / res["f32"]["s"],
           "config": {"workload": "portfolio_gbm_stored_slices", "paths_per_gpu": lay.n_total, "n_steps": steps, "slice_stride": stride,
                      "n_slices": steps // stride + 1, "layout": "time-major (n_slices, n_paths)", "seed": 42,
                      "stream_layout": f"reference parallel blocks, n_workers={8 * cx.world}", "resident": "device"},
           "stages": res,
           "roofline": cx.imad_roofline(draws / 4.0, r["plan_s"], "plan_scan_philox_stream_kernel"),
           "store_roofline": cx.hbm_roofline(store_bytes, r["store_ms"] * 1e-3, "gbm_slices_from_sums_kernel<double>"),
           "bound_note": "RNG (integer-multiply) bound, not HBM-write bound: the planning scan is "
                         f"{r['plan_s'] / r['s']:.0%} of the job, the slice store kernel {r['store_ms'] * 1e-3 / r['s']:.0%}",
           "checks": {"mean_terminal_wealth_rank0": last_mean, "expected": 1e4 * float(np.exp(0.7))}}
    if cx.world == 1 and not cx.args.no_cpu_baseline:
        import bench_reference as R

        out["cpu_baseline"] = cpu_line(R.portfolio, 40_000, 10, unit="path-steps/s")
    return out


def cfg_lsm(cx):
    """cfg 4: American put by Longstaff-Schwartz on 1e7 paths x 252 steps (time-major paths from the GBM path kernel).
    N > 1: the SAME total path count sharded over the ranks (strong scaling; one 64-byte all-reduce per step)."""
    import mcframework_b200 as mcf
    from mcframework_b200 import _stats_device as SD

    torch = cx.torch
    n_total, steps = int(1e7 * cx.args.config_scale), 252
    n = n_total // cx.world
    sim = mcf.BlackScholesPathSimulation()
    sim.simulate_paths(2048, n_steps=steps, device=True, time_major=True)  # first use of the PCG64 kernels (module load)
    sim.set_seed(1 + 1000 * cx.rank)  # each rank draws its own shard of paths
    cx.barrier()
    a = ev(torch)
    paths = sim.simulate_paths(n, n_steps=steps, device=True, time_major=True)
    b = ev(torch)
    cx.barrier()
    t_paths = cx.max_over_ranks(a.elapsed_time(b) * 1e-3)
    t, price = best_of(cx, lambda: SD.lsm_price(paths, 100.0, 0.05, 1.0 / steps, True, group=cx.group), reps=3)
    alg = (steps - 1) * n * 24  # fused step kernel: S_t (decision) + S_{t-1} (regression) + discounted cash flow, 8 B each
    out = {"metric": "lsm_path_steps_per_s", "value": n * cx.world * steps / t, "unit": "path-steps/s", "ms": t * 1e3,
           "scaling": "strong" if cx.world > 1 else "weak",
           "config": {"workload": "american_put_longstaff_schwartz", "paths_total": n * cx.world, "paths_per_gpu": n, "n_steps": steps,
                      "K": 100.0, "r": 0.05, "sigma": 0.2, "resident": "device (time-major paths, 8 B x 253 x paths)"},
           "paths_generation": {"s": t_paths, "path_steps_per_s": n * cx.world * steps / t_paths,
                                "store_GBps_per_gpu": n * (steps + 1) * 8 / t_paths / 1e9,
                                "includes": "planning pass of the sequential PCG64 stream + path kernel"},
           "roofline": cx.hbm_roofline(alg, t, "lsm_fused_step_kernel" if cx.world == 1 else
                                       "lsm_fused_step_kernel (the ranks' sums meet inside the kernel, over peer memory)"),
           "tensor_pipe_note": "3-column regression (8 sums per path): not tensor-core shaped, FP64 FMA; HBM binds",
           "checks": {"american_put_price": float(price.cpu()[0]), "reference_1e6_path_price": 6.053948135228681}}
    if n * 8 <= 79 * (1 << 20):  # the cash flows fit the L2 set-aside (mcf_stats.cu, LsmL2Reservation)
        out["roofline"]["note"] = ("frac counts all 24 algorithmic bytes per path and step against the HBM peak; the 8 bytes of "
                                   "cash flows are kept L2-resident for the sweep (evict-last set-aside), so only the two path "
                                   "columns have to cross HBM and frac can exceed 1; hbm_frac_of_16_bytes is the same time "
                                   "against those 16 bytes")
        out["roofline"]["hbm_frac_of_16_bytes"] = (steps - 1) * n * 16 / t / 1e9 / cx.hbm_peak
    del paths
    if cx.world == 1 and not cx.args.no_cpu_baseline:
        import bench_reference as R

        out["cpu_baseline"] = cpu_line(R.lsm, 20_000, 252, unit="path-steps/s")
    return out


def cfg_stats(cx):
    """cfg 5: StatsEngine on 1e8 results -- moments, percentiles (1,5,50,95,99), BCa bootstrap with 10 000 resamples.
    N > 1: the sample is sharded over the ranks (moments all-gather, histogram all-reduce, slot-sharded bootstrap)."""
    import numpy as np

    import mcframework_b200 as mcf
    from mcframework_b200 import _stats_device as SD
    from mcframework_b200.stats_engine import DeviceSample

    torch = cx.torch
    sc = cx.args.config_scale
    n, B = int(1e8 * sc), int(max(100, 1e4 * sc))
    lo, hi = n * cx.rank // cx.world, n * (cx.rank + 1) // cx.world
    g = torch.Generator(device="cuda")
    g.manual_seed(1234 + cx.rank)
    x = torch.empty(hi - lo, dtype=torch.float64, device="cuda").normal_(5.0, 2.0, generator=g)  # synthetic results
    out = {}
    t, _ = best_of(cx, lambda: SD.moments(x, 5.0, group=cx.group) if cx.world > 1 else SD.moments_raw(x, 5.0), reps=3)
    out["moments"] = {"s": t, "roofline": cx.hbm_roofline((hi - lo) * 8, t, "moments_kernel")}
    t, p = best_of(cx, lambda: SD.percentiles(x, (1, 5, 50, 95, 99), n_valid=n, group=cx.group), reps=3)
    out["percentiles"] = {"s": t, "values": [float(v) for v in p],
                          "roofline": cx.hbm_roofline(8 * (hi - lo) * 8, t, "select_hist/compact (8 radix passes, algorithmic)")}
    # the bootstrap resamples the WHOLE sample: with N ranks every rank holds all of x (one all-gather of the shards, as
    # the engine does) and generates/gathers 1/N of the single index stream's slots; the B partial sums are all-reduced
    full = DeviceSample(x, None, cx.group).full() if cx.world > 1 else x
    gen = np.random.default_rng(7)
    cx.barrier()
    t0 = time.perf_counter()
    means = SD.bootstrap_means(full, B, gen, group=cx.group)
    cx.barrier()
    tb = cx.max_over_ranks(time.perf_counter() - t0)
    del full
    # the whole 12-metric engine, as run() calls it (ci_mean_bootstrap = BCa with its own B resamples)
    eng = mcf.stats_engine.build_default_engine()
    B_eng = max(100, B // 10)  # (the full-B bootstrap was timed above; the engine pass shows the other 11 metrics ride along)
    ctx = mcf.StatsContext(n=n, percentiles=(1, 5, 50, 95, 99), bootstrap="bca", n_bootstrap=B_eng, rng=7, target=5.0, eps=0.01)
    cx.barrier()
    t0 = time.perf_counter()
    res = eng.compute(DeviceSample(x, None, cx.group), ctx)
    cx.barrier()
    te = cx.max_over_ranks(time.perf_counter() - t0)
    ci = res.metrics.get("ci_mean_bootstrap") or {}
    out.update({"metric": "bootstrap_resamples_per_s", "value": B / tb, "unit": "resamples/s (n = %d)" % n, "ms": tb * 1e3,
                "scaling": "strong" if cx.world > 1 else "weak",
                "config": {"workload": "stats_engine_on_1e8_results", "n": n, "n_bootstrap": B, "percentiles": [1, 5, 50, 95, 99],
                           "bootstrap": "bca", "sharded_over_ranks": cx.world, "resident": "device"},
                "gathers_per_s": B * n / tb,
                "roofline": {**cx.hbm_roofline(B * n * 8.0 / cx.world, tb, "boot_count/boot_bin/boot_gather"),
                             "note": "algorithmic bytes = 8 B per gather; the binding resources are the SM->L2 request rate of the "
                                     "gather and the integer-multiply pipe of the PCG64 index stream (DESIGN.md section 3.5)"},
                "engine_all_12_metrics_s": te, "engine_n_bootstrap": B_eng,
                "checks": {"bca_ci": [ci.get("low"), ci.get("high")], "mean_of_means": float(SD.moments(means).mean),
                           "sample_mean": float(res.metrics.get("mean", float("nan")))}})
    if cx.world == 1 and not cx.args.no_cpu_baseline:
        import bench_reference as R

        # the reference's bootstrap costs n x B gathers; normalise the reduced-size rate to resamples of n = 1e8
        cb = cpu_line(R.bootstrap, 100_000, 1000, unit="resamples/s (n = %d)" % n, scale=100_000.0 / n)
        out["cpu_baseline"] = cb
        out["cpu_baseline_percentiles"] = cpu_line(R.percentiles, 10_000_000, unit="values/s")
        out["percentiles"]["values_per_s"] = n / out["percentiles"]["s"]
    return out


def b200_arm(args):
    import logging

    cx = Ctx(args)
    from mcframework_b200 import _native as N

    N.lib()
    logging.getLogger("mcframework_b200.core").setLevel(logging.WARNING)
    line = headline(cx)
    wanted = [] if args.configs == "none" else (list(ALL_CONFIGS) if args.configs == "all" else args.configs.split(","))
    table = {"pi": cfg_pi, "portfolio_slices": cfg_portfolio, "lsm": cfg_lsm, "stats": cfg_stats}
    configs = {}
    for name in wanted:
        cx.torch.cuda.empty_cache()
        t0 = time.perf_counter()
        try:
            entry = table[name](cx)
        except Exception as exc:  # noqa: BLE001 - one failing config must not lose the headline line
            import traceback

            entry = {"error": repr(exc), "trace": traceback.format_exc()[-1500:]}
            if cx.dist is not None:
                raise
        if entry is not None:
            entry["n_gpus"] = cx.world
            entry["bench_wall_s"] = time.perf_counter() - t0
        configs[name] = entry
    if cx.rank == 0:
        line["configs"] = configs
        print(json.dumps(line), flush=True)
    if cx.dist is not None:
        cx.dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        b200_arm(a)
