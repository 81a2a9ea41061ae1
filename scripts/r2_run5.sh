#!/bin/bash
o=gpurun_out
timeout 900 python -m pytest tests/test_gpu_api.py -q -m gpu -x > $o/r2e_gpu_tests.log 2>&1; tail -2 $o/r2e_gpu_tests.log
timeout 600 python bench.py 2>$o/r2e_bench.err | tail -1 > $o/r2e_bench_f64.json; tail -3 $o/r2e_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench_f64.json'))
print('value', d['value']/1e6, 'ms', d['ms_per_step'], 'e2e', d['e2e'], 'other', d['other_output_mode'])
print('variants', d['variants'])
print('roofline', {k:v for k,v in d['roofline'].items() if k not in ('hbm','whole_step')})
print('hbm', d['roofline']['hbm'])
print('cpu', d['cpu_baseline'], d['cpu_baseline_ref'])
PY
