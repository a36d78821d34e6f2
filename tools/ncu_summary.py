#!/usr/bin/env python
"""Summarise ncu output into profiles/ (run in the build container; reads gpurun_out/).

    python tools/ncu_summary.py --tag r1 --rep gpurun_out/prof_r1.ncu-rep --launches gpurun_out/launches_r1.csv \
        --paths 1e7 --steps 252

Writes profiles/<tag>_launches.md (per-kernel share of the step), profiles/<tag>_kernels.json (selected counters of
every captured launch) and profiles/inst_per_unit.json (warp instructions per path-step, read by bench.py).
"""
import argparse
import collections
import csv
import json
import os
import subprocess

KEYS = [
    "gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.avg",
]
MB = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(name):
    return name.split("(")[0].replace("void ", "").strip()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", required=True)
    ap.add_argument("--rep")
    ap.add_argument("--launches")
    ap.add_argument("--paths", type=float, default=1e7)
    ap.add_argument("--steps", type=int, default=252)
    ap.add_argument("--cmd", default="python bench.py")
    ap.add_argument("--streams", type=int, default=64, help="plan scan = one launch per stream; captured launches are scaled")
    a = ap.parse_args()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out_dir = os.path.join(root, "profiles")
    os.makedirs(out_dir, exist_ok=True)

    if a.launches:
        lines = [ln for ln in open(a.launches) if ln.startswith('"')]
        tot = collections.OrderedDict()
        for row in csv.DictReader(lines):
            k = short(row["Kernel Name"])
            tot.setdefault(k, [0, 0.0])
            tot[k][0] += 1
            tot[k][1] += float(row["Metric Value"]) / 1e6
        total = sum(v[1] for v in tot.values())
        with open(os.path.join(out_dir, f"{a.tag}_launches.md"), "w") as f:
            f.write(f"# ncu launch list ({a.tag})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over `{a.cmd}`\n"
                    "(cold-cache, serialised: compare SHARES, not absolutes).\n\n| kernel | launches | total ms | avg ms | share |\n|---|---|---|---|---|\n")
            for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
                f.write(f"| `{k}` | {v[0]} | {v[1]:.2f} | {v[1] / v[0]:.3f} | {100 * v[1] / total:.1f} % |\n")
            f.write(f"\ntotal device time {total:.1f} ms over {sum(v[0] for v in tot.values())} launches\n")

    if a.rep:
        raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units, data = rows[0], rows[1], rows[2:]
        idx = {h: i for i, h in enumerate(hdr)}
        kernels = []
        for row in data:
            rec = {"kernel": short(row[idx["Kernel Name"]]), "grid": row[idx["Grid Size"]], "block": row[idx["Block Size"]]}
            for k in KEYS:
                if k in idx:
                    v = float(row[idx[k]].replace(",", ""))
                    if k.startswith("dram__bytes"):
                        v *= MB.get(units[idx[k]], 1.0)
                    rec[k] = v
            kernels.append(rec)
        units_total = a.paths * a.steps
        inst = {"source": f"profiles/{a.tag}_kernels.json (ncu --set full, {a.cmd} --paths {a.paths:g})", "paths": a.paths, "steps": a.steps}
        # the plan scan is one launch per stream: merge each run of consecutive launches of the same kernel
        runs = []
        for rec in kernels:
            if runs and runs[-1]["kernel"] == rec["kernel"] and "plan_scan" in rec["kernel"]:
                run = runs[-1]
                run["launches"] += 1
                for k in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"):
                    run[k] = run.get(k, 0) + rec.get(k, 0)
            else:
                runs.append(dict(rec, launches=1))
        for run in runs:
            name = run["kernel"]
            wi = run.get("smsp__inst_executed.sum")
            traffic = run.get("dram__bytes_read.sum", 0) + run.get("dram__bytes_write.sum", 0)
            if wi and "plan_scan" in name and "plan_warp_inst_per_path_step" not in inst:
                # `launches` of the a.streams equally sized per-stream launches were captured
                frac = run["launches"] / a.streams if "philox_stream" in name else 1.0
                inst["plan_warp_inst_per_path_step"] = wi / (units_total * frac)
                inst["plan_dram_bytes_per_path_step"] = traffic / (units_total * frac)
                inst["plan_scan_launches_captured"] = run["launches"]
            if wi and "gbm_terminal" in name:
                key = "replay8" if run.get("dram__bytes_write.sum", 0) > 3 * a.paths * 8 else "replay1"
                inst.setdefault(key + "_warp_inst_per_path_step", wi / units_total)
                inst.setdefault(key + "_dram_bytes_per_path_step", traffic / units_total)
        json.dump(kernels, open(os.path.join(out_dir, f"{a.tag}_kernels.json"), "w"), indent=1)
        json.dump(inst, open(os.path.join(out_dir, "inst_per_unit.json"), "w"), indent=1)
        print(json.dumps(inst, indent=1))


if __name__ == "__main__":
    main()
