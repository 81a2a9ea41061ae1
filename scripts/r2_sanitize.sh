#!/bin/bash
# compute-sanitizer over the round-2 paths (scripts/sanitize.py); memcheck on everything, racecheck + synccheck on the solver-heavy subset
o=gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 $CS --tool memcheck --error-exitcode 1 python scripts/sanitize.py > $o/r2_sanitize_memcheck.txt 2>&1; echo memcheck rc=$? ; tail -3 $o/r2_sanitize_memcheck.txt
MJT_SANITIZE_ONLY=humanoid,capsules_topk,terrain,muscle,hopper timeout 1200 $CS --tool racecheck --error-exitcode 1 python scripts/sanitize.py > $o/r2_sanitize_racecheck.txt 2>&1; echo racecheck rc=$?; tail -3 $o/r2_sanitize_racecheck.txt
MJT_SANITIZE_ONLY=humanoid,capsules_topk,terrain,hopper timeout 900 $CS --tool synccheck --error-exitcode 1 python scripts/sanitize.py > $o/r2_sanitize_synccheck.txt 2>&1; echo synccheck rc=$?; tail -3 $o/r2_sanitize_synccheck.txt
