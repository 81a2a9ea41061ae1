#!/bin/bash
o=gpurun_out
timeout 600 python -m pytest tests/test_gpu_step.py tests/test_gpu_api.py tests/test_golden.py -q -x -m gpu -k "stack or live_tile" > $o/r2r_tests.log 2>&1; tail -2 $o/r2r_tests.log
bash scripts/quick_bench.sh "--outputs state --model stack --batch 65536 --steps 5 --dtype f32 --no-other-mode --no-variants" "--outputs state --model stack --batch 16384 --steps 5 --no-other-mode --no-variants" 2>&1 | tee $o/r2r_bench.txt
