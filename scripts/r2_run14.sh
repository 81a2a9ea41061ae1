#!/bin/bash
o=gpurun_out
for v in "" call t448 call448; do
  if [ -n "$v" ]; then export MJT_LIB=$PWD/build_variants/libmjt_$v.so; else unset MJT_LIB; fi
  echo "== variant: ${v:-base}"
  bash scripts/quick_bench.sh "--outputs state --no-other-mode --no-variants" "--outputs state --no-other-mode --no-variants --model halfcheetah"
done 2>&1 | tee $o/r2t_variants.txt
