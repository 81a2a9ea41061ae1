#!/bin/bash
o=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > $o/r2c_gpu_tests.log 2>&1; tail -3 $o/r2c_gpu_tests.log
rm -f $o/r2c_bench.jsonl
for a in "--outputs full" "--outputs state" "--outputs state --dtype f32" "--outputs state --integrator implicitfast" "--outputs state --model halfcheetah" "--outputs state --model stack --batch 16384 --steps 10" ; do
  timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e $a 2>>$o/r2c_bench.err | tail -1 >> $o/r2c_bench.jsonl
done
python - <<'PY'
import json
for ln in open('gpurun_out/r2c_bench.jsonl'):
    try:
        d = json.loads(ln)
        print(d['dtype'], d['config']['workload'][:70], f"{d['value']/1e6:8.3f} M/s {d['ms_per_step']:8.3f} ms", {k: round(v, 3) for k, v in d['roofline']['kernel_ms_per_step'].items()})
    except Exception as e:
        print('bad line', ln[:200])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_solve -c 1 -f -o $o/r2c_state_f64_solve \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --outputs state > $o/r2c_ncu.log 2>&1
ls -la $o | tail -5
