#!/bin/bash
o=gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
MJT_SANITIZE_ONLY=humanoid,capsules_topk,terrain,muscle,hopper,halfcheetah,stack timeout 1200 $CS --tool racecheck --print-limit 400 --error-exitcode 1 python scripts/sanitize.py > $o/r2_sanitize_racecheck.txt 2>&1; echo racecheck rc=$?; tail -3 $o/r2_sanitize_racecheck.txt
timeout 300 python -m pytest tests/test_gpu_api.py -q -x -k "bit_identical" > $o/r2l_tests.log 2>&1; tail -2 $o/r2l_tests.log
bash scripts/r2_ncu_stack.sh
