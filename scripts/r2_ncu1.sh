#!/bin/bash
o=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ -c 4 -f -o $o/r2a_state_f64 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --outputs state > $o/r2a_ncu.log 2>&1
ls -la $o
