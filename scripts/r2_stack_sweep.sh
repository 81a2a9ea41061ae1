#!/bin/bash
# BASELINE config 5: box stack, RK4, f32, batch sweep; usage: scripts/r2_stack_sweep.sh <ngpus> "<batch sizes per GPU>"
N=$1; SIZES=$2
o=gpurun_out
for B in $SIZES; do
  steps=5; [ $B -ge 262144 ] && steps=3
  if [ $N -eq 1 ]; then
    timeout 600 python bench.py --model stack --dtype f32 --batch $B --steps $steps --warmup 3 --no-cpu-baseline --no-e2e --no-other-mode --no-variants 2>>$o/r2_sweep.err | tail -1 >> $o/r2_stack_sweep_n$N.jsonl
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --model stack --dtype f32 --batch $B --steps $steps --warmup 3 --no-cpu-baseline --no-e2e --no-other-mode --no-variants --no-gather-obs 2>>$o/r2_sweep.err | tail -1 >> $o/r2_stack_sweep_n$N.jsonl
  fi
done
python - <<PY
import json
for ln in open("gpurun_out/r2_stack_sweep_n$N.jsonl"):
    try:
        d = json.loads(ln); print(d["n_gpus"], d["config"]["batch_per_gpu"], round(d["value"]/1e6, 4), "M env-steps/s", round(d["ms_per_step"], 2), "ms", {k: round(v, 2) for k, v in d["roofline"]["kernel_ms_per_step"].items()})
    except Exception as e:
        print("bad", ln[:200])
PY
