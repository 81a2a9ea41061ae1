#!/bin/bash
o=gpurun_out
for c in 1 2 3 4; do
  MJT_E2E_CHUNKS=$c timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-mode --no-variants 2>>$o/r2v.err | tail -1 > /tmp/e.json
  python - $c <<'PY'
import json, sys
d = json.load(open('/tmp/e.json'))
print("chunks", sys.argv[1], "value", round(d["value"] / 1e6, 2), "e2e", round(d["e2e"]["value"] / 1e6, 2), "host-state", round(d["e2e_host_state"]["value"] / 1e6, 2))
PY
done 2>&1 | tee $o/r2v_chunks.txt
