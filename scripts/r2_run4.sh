#!/bin/bash
o=gpurun_out
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_golden.py tests/test_gpu_api.py -q -m gpu -x > $o/r2d_gpu_tests.log 2>&1; tail -2 $o/r2d_gpu_tests.log
bash scripts/quick_bench.sh "--outputs state" "--outputs full" "--outputs state --dtype f32" "--outputs state --model halfcheetah" "--outputs state --model ant --batch 32768" 2>&1 | tee $o/r2d_bench.txt
bash scripts/run_variants.sh "--outputs state" sm6 sm8 c6 c8 t256 t64 cmp32 2>&1 | tee $o/r2d_variants.txt
bash scripts/run_variants.sh "--outputs state --dtype f32" sm6 sm8 c8 cmp32 2>&1 | tee -a $o/r2d_variants.txt
