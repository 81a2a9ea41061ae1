#!/bin/bash
# round 2, first GPU pass: the whole GPU tier, then full-vs-state bench lines
o=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > $o/r2a_gpu_tests.log 2>&1; tail -40 $o/r2a_gpu_tests.log
for a in "--outputs full" "--outputs state" "--outputs full --dtype f32" "--outputs state --dtype f32" "--outputs state --integrator implicitfast" "--outputs state --integrator implicitfast --dtype f32"; do
  timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e $a 2>>$o/r2a_bench.err | tail -1 >> $o/r2a_bench.jsonl
done
python - <<'PY'
import json
for ln in open('gpurun_out/r2a_bench.jsonl'):
    try:
        d = json.loads(ln)
        print(d['dtype'], d['config']['workload'][:60], f"{d['value']/1e6:8.3f} M/s {d['ms_per_step']:8.3f} ms", {k: round(v, 3) for k, v in d['roofline']['kernel_ms_per_step'].items()})
    except Exception as e:
        print('bad line', ln[:200])
PY
