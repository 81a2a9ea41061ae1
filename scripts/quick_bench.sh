#!/bin/bash
# usage: scripts/quick_bench.sh "<args1>" "<args2>" ...   one line per run: value, ms/step, kernel split
for a in "$@"; do
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e $a 2>&1 | tail -1 > /tmp/qb.json
  python - "$a" <<'PY'
import sys, json
try:
    d = json.loads(open('/tmp/qb.json').read())
    print(f"{sys.argv[1]:45s} {d['value']/1e6:8.3f} M/s {d['ms_per_step']:8.3f} ms", {k: round(v, 3) for k, v in d['roofline']['kernel_ms_per_step'].items()}, 'E', d['config'].get('solve_envs_per_cta'), 'smem', d['config'].get('solve_smem_bytes'), d['config'].get('solve_live_tile', ''))
except Exception as e:
    print(sys.argv[1], 'failed:', open('/tmp/qb.json').read()[-300:])
PY
done
