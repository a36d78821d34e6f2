#!/bin/bash
# Builds alternative code-generation variants of the library next to the default one (for tools/ab_variants.sh).
# usage: tools/build_variants.sh name="-DFLAG=..." ...   -> mcframework_b200/_variants/libmcf_<name>.so
ROOT=$(cd "$(dirname "$0")/.." && pwd)
cd "$ROOT/mcframework_b200/csrc" || exit 1
mkdir -p ../_variants
for spec in "$@"; do
  name=${spec%%=*}; flags=${spec#*=}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -shared \
    $flags -o ../_variants/libmcf_$name.so mcf_kernels.cu mcf_stats.cu &
done
wait
ls -la ../_variants
