#!/bin/bash
for c in "$@"; do
  MJT_E2E_CHUNKS=$c python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > /tmp/e.json
  python -c "import json; d=json.load(open('/tmp/e.json')); print('chunks', $c, 'value', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2))"
done
