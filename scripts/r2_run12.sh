#!/bin/bash
o=gpurun_out
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_api.py tests/test_golden.py -q -x -m gpu > $o/r2q_tests.log 2>&1; tail -2 $o/r2q_tests.log
bash scripts/quick_bench.sh "--outputs state --model stack --batch 65536 --steps 5 --dtype f32" "--outputs state --model stack --batch 16384 --steps 5" "--outputs state" 2>&1 | tee $o/r2q_bench.txt
