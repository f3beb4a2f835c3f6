#!/bin/bash
# Builds differently parameterised copies of libuwtrack.so into build/variants/ for A/B
# runs on the GPU box:  UWT_LIBRARY=build/variants/<name>.so python bench.py ...
# usage: tools/ab_variants.sh name1:"-DFLAG=.." name2:"..."
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
    -fmad=false -Xcompiler -fPIC -Xcompiler -ffp-contract=off -shared -cudart static $flags \
    -o build/variants/$name.so uw_slam_b200/csrc/*.cu &
done
wait
ls -la build/variants/
