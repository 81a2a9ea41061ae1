#!/bin/bash
# Variants of the thread-per-env kernels only (k_smooth / k_collision / k_constraint / k_sensor live in mjt_api.cu): compile that
# one translation unit with the given flags and link it with the solver objects of the current in-tree build.
# usage: scripts/build_api_variants.sh name "-DMJT_X=.." [name2 "flags2" ...]
set -e
cd "$(dirname "$0")/.."
mkdir -p build_variants
B=mujoco_torch_b200/csrc/build
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  ( cd mujoco_torch_b200/csrc && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $flags \
      -c mjt_api.cu -o ../../build_variants/api_$name.o && \
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared ../../build_variants/api_$name.o build/mjt_solve_f32.o build/mjt_solve_f64.o build/mjt_peak.o -o ../../build_variants/libmjt_$name.so ) &
done
wait
ls -la build_variants/*.so
