#!/bin/bash
# Build tuning variants of libmjt_b200.so into build_variants/ (git-ignored, travels with gpurun).
# usage: scripts/build_variants.sh name "-DMJT_X=.. -DMJT_Y=.." [name2 "flags2" ...]
set -e
cd "$(dirname "$0")/.."
mkdir -p build_variants
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  ( cd mujoco_torch_b200/csrc && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
      -Xcompiler -fPIC -shared $flags mjt_api.cu mjt_solve_f32.cu mjt_solve_f64.cu mjt_peak.cu mjt_env.cu -o ../../build_variants/libmjt_$name.so ) &
done
wait
ls -la build_variants
