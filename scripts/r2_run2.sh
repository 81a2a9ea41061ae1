#!/bin/bash
# quick correctness subset + bench lines
o=gpurun_out
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_golden.py -q -m gpu -x > $o/r2b_gpu_tests.log 2>&1; tail -5 $o/r2b_gpu_tests.log
rm -f $o/r2b_bench.jsonl
for a in "--outputs full" "--outputs state" "--outputs full --dtype f32" "--outputs state --dtype f32" "--outputs state --integrator implicitfast" "--outputs state --model halfcheetah" "--outputs state --model ant --batch 32768" "--outputs state --model stack --batch 16384 --steps 10" ; do
  timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e $a 2>>$o/r2b_bench.err | tail -1 >> $o/r2b_bench.jsonl
done
python - <<'PY'
import json
for ln in open('gpurun_out/r2b_bench.jsonl'):
    try:
        d = json.loads(ln)
        print(d['dtype'], d['config']['workload'][:70], f"{d['value']/1e6:8.3f} M/s {d['ms_per_step']:8.3f} ms", {k: round(v, 3) for k, v in d['roofline']['kernel_ms_per_step'].items()})
    except Exception as e:
        print('bad line', ln[:200])
PY
