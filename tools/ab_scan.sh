#!/bin/bash
# A/B of the scan code-generation variants built by tools/build_variants.sh: one ab_probe.py line per library.
# usage (GPU box): tools/ab_scan.sh name ...   (name "base" = the in-tree default)  -> gpurun_out/ab_scan.jsonl
mkdir -p gpurun_out
: > gpurun_out/ab_scan.jsonl
for name in "$@"; do
  if [ "$name" = base ]; then unset MCF_B200_LIB; else export MCF_B200_LIB=$(readlink -f mcframework_b200/_variants/libmcf_$name.so); fi
  timeout 300 python tools/ab_probe.py "$name" >> gpurun_out/ab_scan.jsonl 2>> gpurun_out/ab_scan.err || echo "{\"label\": \"$name\", \"failed\": true}" >> gpurun_out/ab_scan.jsonl
done
cat gpurun_out/ab_scan.jsonl
