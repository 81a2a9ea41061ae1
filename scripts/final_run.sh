#!/bin/bash
# Round-end evidence on one B200: GPU test tier, smoke, the three bench lines, ncu launch list + one full capture of the
# live-tile solver, then the BASELINE config sweep.  Everything lands in gpurun_out/.
tag=${1:-r1_s3}
o=gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu > $o/gpu_tests_$tag.log 2>&1; tail -3 $o/gpu_tests_$tag.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 python bench.py 2>$o/bench_$tag.err | tail -1 > $o/bench_${tag}_f64.json
timeout 300 python bench.py --dtype f32 2>>$o/bench_$tag.err | tail -1 > $o/bench_${tag}_f32.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 2>>$o/bench_$tag.err | tail -1 > $o/bench_${tag}_reference.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_$tag.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_solve -c 2 -f -o $o/prof_${tag}_stack_tile \
  python bench.py --model stack --batch 4096 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 420 bash scripts/config_sweep.sh > $o/configs_$tag.txt 2>&1; cp $o/configs_r1.jsonl $o/configs_$tag.jsonl
head -c 600 $o/bench_${tag}_f64.json; echo; tail -16 $o/configs_$tag.txt
