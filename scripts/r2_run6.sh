#!/bin/bash
o=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > $o/r2i_gpu_tests.log 2>&1; tail -3 $o/r2i_gpu_tests.log
bash scripts/quick_bench.sh "--outputs state" "--outputs state --dtype f32" "--outputs state --model stack --batch 16384 --steps 10 --dtype f32" "--outputs state --model stack --batch 16384 --steps 10" 2>&1 | tee $o/r2i_bench.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_solve|k_collision" -c 3 -f -o $o/r2i_stack_f32 \
  python bench.py --model stack --batch 4096 --dtype f32 --steps 1 --warmup 2 --no-cpu-baseline --no-e2e --no-other-mode --no-variants > $o/r2i_ncu.log 2>&1
ls -la $o | tail -4
