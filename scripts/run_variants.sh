#!/bin/bash
# usage: scripts/run_variants.sh "<bench args>" v1 v2 ...   (prints per-kernel ms for each build variant)
args=$1; shift
for v in "$@"; do
  MJT_LIB=$PWD/build_variants/libmjt_$v.so python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e $args 2>&1 | tail -1 > /tmp/rv.json
  python - "$v" "$args" <<'PY'
import sys, json
try:
    d = json.loads(open('/tmp/rv.json').read())
    c = d.get('clocks') or {}
    print(sys.argv[1], sys.argv[2], '::', round(d['value'] / 1e6, 2), 'M/s', round(d['ms_per_step'], 3), 'ms',
          {k: round(v, 3) for k, v in d['roofline']['kernel_ms_per_step'].items()}, 'clk', c.get('sm_mhz'), 'E', d['config'].get('solve_envs_per_cta'), 'smem', d['config'].get('solve_smem_bytes'))
except Exception as e:
    print(sys.argv[1], 'failed:', open('/tmp/rv.json').read()[:300])
PY
done
