#!/bin/bash
o=gpurun_out
timeout 600 python -m pytest tests/test_gpu_api.py -q -x -k "graphed or sensor_slots" > $o/r2k_tests.log 2>&1; tail -15 $o/r2k_tests.log
bash scripts/small_batch_sweep.sh
