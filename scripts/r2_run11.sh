#!/bin/bash
o=gpurun_out
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_api.py tests/test_golden.py -q -x -m gpu > $o/r2p_tests.log 2>&1; tail -2 $o/r2p_tests.log
bash scripts/quick_bench.sh "--outputs state" "--outputs state --model stack --batch 65536 --steps 5 --dtype f32" "--outputs state --model stack --batch 16384 --steps 5" "--outputs state --model halfcheetah" "--outputs state --model ant --batch 32768" "--outputs state --model walker2d --batch 32768" "--outputs state --model hopper --batch 131072" 2>&1 | tee $o/r2p_bench.txt
