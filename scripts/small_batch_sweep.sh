#!/bin/bash
# Humanoid f64 at the batch sizes the reference publishes (README.md:187-191): step() per step vs GraphedStep replay
o=gpurun_out
rm -f $o/r2k_small_batch.jsonl
for B in 1 128 1024 4096 32768 65536; do
  for G in "" "--graph"; do
    timeout 300 python bench.py --batch $B --steps 200 --warmup 20 $G --no-cpu-baseline --no-e2e --no-other-mode --no-variants 2>>$o/r2k.err | tail -1 >> $o/r2k_small_batch.jsonl
  done
done
python - <<PY
import json
for ln in open("gpurun_out/r2k_small_batch.jsonl"):
    try:
        d = json.loads(ln)
        print(d["config"]["batch_per_gpu"], d["config"]["call"][:12], round(d["value"]), "env-steps/s", round(d["ms_per_step"], 4), "ms", "vs", d["vs_baseline"] and round(d["vs_baseline"], 1))
    except Exception as e:
        print("bad", ln[:200])
PY
