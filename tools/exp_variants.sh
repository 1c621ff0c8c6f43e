#!/bin/bash
# What-if builds of the library (same sources + one -D switch each) under tools/exp_libs/; run on the GPU box with
# tools/gpu_exp.sh.  Usage: exp_variants.sh NAME:FLAGS ...   e.g.  exp_variants.sh base: nob:-DSMR_EXP_NO_B
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/exp_libs
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  /usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
    -Xcompiler -fPIC,-ffp-contract=off -ccbin /usr/bin/g++ -shared $flags -o tools/exp_libs/libsmr_$name.so \
    smelter_b200/csrc/kernels.cu smelter_b200/csrc/renderer.cpp smelter_b200/csrc/scene.cpp 2>&1 | grep -v "Wcomment\|within comment\|^ *[0-9]* |\|^ *|" &
done
wait
ls -la tools/exp_libs/
