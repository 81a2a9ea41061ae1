#!/bin/bash
# BASELINE.json configs on one B200 (device-resident value + e2e), one JSON line each into gpurun_out/configs_r1.jsonl
out=gpurun_out/configs_r1.jsonl; : > $out
run() { python bench.py --steps 50 --warmup 5 --no-cpu-baseline "$@" 2>/dev/null | tail -1 >> $out; }
run --model halfcheetah --batch 4096                      # config 2
run --model halfcheetah --batch 65536
run --model ant_zoo --batch 32768 --dtype f32 --cone elliptic --solver cg   # config 3
run --model ant --batch 32768                             # README "Ant" (raw ant.xml)
run --model humanoid --batch 65536                        # config 4 (Euler = published setup)
run --model humanoid --batch 65536 --dtype f32
run --model humanoid --batch 65536 --integrator implicitfast
run --model humanoid --batch 65536 --integrator implicitfast --dtype f32
run --model humanoid --batch 131072
run --model humanoid --batch 4096
for b in 1024 16384 65536; do run --model stack --batch $b; done   # config 5 (RK4)
run --model walker2d --batch 32768
run --model hopper --batch 131072
python - <<'PY'
import json
for l in open('gpurun_out/configs_r1.jsonl'):
    try: d=json.loads(l)
    except Exception: print('bad line', l[:100]); continue
    e=d.get('e2e') or {}
    print(f"{d['config']['workload']:95s} {d['dtype']} value {d['value']/1e6:8.3f} M/s  e2e {e.get('value',0)/1e6:8.3f} M/s  {d['ms_per_step']:.3f} ms  {({k:round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()})}")
PY
