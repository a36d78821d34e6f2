#!/bin/bash
# Multi-GPU run of one box: parity checks of the sharded paths, the sharded LSM A/B, and bench.py (weak scaling with every
# config, strong scaling, and the root-gather A/B of the headline).  Usage (from the repo root, under gpurun --gpus N):
#   bash tools/run_multi_gpu.sh N        -> gpurun_out/bench_r2_${N}gpu*.json
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR tools/multi_gpu_check.py 2>&1 | grep -v "INFO\|Warning\|warn" | tail -12
timeout 300 $TR tools/ab_lsm_multi.py 2>&1 | grep "^{" | tail -1
timeout 900 $TR bench.py --gpus $N > gpurun_out/bench_r2_${N}gpu.json 2> gpurun_out/bench_r2_${N}gpu.err; tail -3 gpurun_out/bench_r2_${N}gpu.err | cut -c1-300
timeout 600 $TR bench.py --gpus $N --scaling strong --configs none > gpurun_out/bench_r2_${N}gpu_strong.json 2>> gpurun_out/bench_r2_${N}gpu.err
MCF_BENCH_GATHER=root timeout 600 $TR bench.py --gpus $N --configs none > gpurun_out/bench_r2_${N}gpu_rootgather.json 2>> gpurun_out/bench_r2_${N}gpu.err
python - <<PY
import json
for f in ("bench_r2_${N}gpu.json","bench_r2_${N}gpu_strong.json","bench_r2_${N}gpu_rootgather.json"):
    try:
        j=json.loads([l for l in open("gpurun_out/"+f) if l.startswith("{")][-1])
    except Exception as e:
        print(f, "FAILED", e); continue
    print(f, "value %.4e ms %.1f e2e %.4e e2e_ms %.1f scaling %s parity %s"%(j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e"]["ms_per_step"], j["scaling"], j["checks"].get("sharded_parity")))
    for k,v in (j.get("configs") or {}).items():
        print("   ", k, "value %.4e %s ms %.2f frac %.3f"%(v.get("value",0), v.get("unit"), v.get("ms",0), (v.get("roofline") or {}).get("frac",0)) if "error" not in v else v["error"][:300])
PY
