#!/bin/bash
# Single-GPU round check: the whole `-m gpu` suite, bench.py with every config, smoke().  Under gpurun, from the repo root:
#   bash tools/run_single_gpu.sh   -> gpurun_out/final/{r2_gpu_tests.log,r2_bench.json}
mkdir -p gpurun_out/final
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/final/r2_gpu_tests.log 2>&1; tail -2 gpurun_out/final/r2_gpu_tests.log
python bench.py > gpurun_out/final/r2_bench.json 2> gpurun_out/final/r2_bench.err; python - <<'PY'
import json
j = json.loads([l for l in open("gpurun_out/final/r2_bench.json") if l.startswith("{")][-1])
print("headline %.4e ms %.1f e2e %.4e frac %.3f" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["roofline"]["frac"]))
for k, v in j["configs"].items():
    print("  ", k, ("value %.4e %s ms %.2f frac %.3f" % (v.get("value", 0), v.get("unit"), v.get("ms", 0), (v.get("roofline") or {}).get("frac", 0))) if "error" not in v else v["error"][:300])
PY
python -c "
import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for n in 5e6 1e7 2e7; do AB_PATHS=$n python tools/ab_lsm.py lsm_$n 2>/dev/null | grep "^{"; done
