#!/bin/bash
o=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_solve" -c 1 -f -o $o/r2l_stack_f32_solve \
  python bench.py --model stack --batch 4096 --dtype f32 --steps 1 --warmup 2 --no-cpu-baseline --no-e2e --no-other-mode --no-variants > $o/r2l_ncu.log 2>&1
ls -la $o | tail -3
