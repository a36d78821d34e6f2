#!/bin/bash
# A/B runs of alternative builds of the C-ABI library (same ABI, different code generation) on the GPU box.
# usage: tools/ab_variants.sh <label>=<path-to-.so> ...     (label "base" = the in-tree default library)
# For each build: the device parity tests, then a short bench line -> gpurun_out/ab_<label>.{log,json}
mkdir -p gpurun_out
for spec in "$@"; do
  label=${spec%%=*}; lib=${spec#*=}
  if [ "$label" = base ]; then unset MCF_B200_LIB; else export MCF_B200_LIB=$(readlink -f "$lib"); fi
  timeout 120 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -q -x -p no:cacheprovider > gpurun_out/ab_$label.log 2>&1
  echo "[$label] pytest rc=$? $(tail -1 gpurun_out/ab_$label.log)"
  timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$label.json 2>> gpurun_out/ab_$label.log
  python - "$label" <<'PY'
import json, sys
try:
    j = json.load(open(f"gpurun_out/ab_{sys.argv[1]}.json"))
    print(f"[{sys.argv[1]}] value {j['value']:.4e} ms/step {j['ms_per_step']:.2f} e2e {j['e2e']['value']:.4e} groups {j['roofline']['ms_per_step_by_kernel_group']} price {j['checks']['price_device']!r}")
except Exception as e:
    print(f"[{sys.argv[1]}] bench failed: {e}")
PY
done
