#!/bin/bash
o=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > $o/r2u_gpu_tests.log 2>&1; tail -2 $o/r2u_gpu_tests.log
bash scripts/quick_bench.sh "--outputs state --no-other-mode --no-variants" "--outputs state --no-other-mode --no-variants --model halfcheetah" "--outputs state --no-other-mode --no-variants --model hopper --batch 131072" "--outputs state --no-other-mode --no-variants --model ant --batch 32768" "--outputs state --no-other-mode --no-variants --model walker2d --batch 32768" 2>&1 | tee $o/r2u_bench.txt
