#!/bin/bash
# final multi-rank check with the default flags the driver uses (usage: r2_final2.sh <ngpus>)
N=$1; o=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N 2>$o/r2o_n$N.err | tail -1 > $o/r2o_bench_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --impl reference --steps 2 --warmup 1 2>>$o/r2o_n$N.err | tail -1 > $o/r2o_bench_n${N}_reference.json
python - <<PY
import json
d = json.load(open("gpurun_out/r2o_bench_n$N.json"))
print(d["n_gpus"], round(d["value"] / 1e6, 2), "M", round(d["ms_per_step"], 3), "ms e2e", d["e2e"] and round(d["e2e"]["value"] / 1e6, 2), "host-state", d["e2e_host_state"] and round(d["e2e_host_state"]["value"] / 1e6, 2),
      "other", d["other_output_mode"] and round(d["other_output_mode"]["value"] / 1e6, 2), "gather", d["obs_gather"])
print(open("gpurun_out/r2o_bench_n${N}_reference.json").read()[:300])
PY
tail -3 $o/r2o_n$N.err
