#!/bin/bash
o=gpurun_out
(nvidia-smi topo -m; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; nproc; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null) > $o/r2g_topo.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-other-mode --no-gather-obs 2>$o/r2g_n8.err | tail -1 > $o/r2g_bench_n8_bound.json
MJT_NO_CPU_BIND=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --no-other-mode --no-gather-obs 2>>$o/r2g_n8.err | tail -1 > $o/r2g_bench_n8_unbound.json
MJT_E2E_CHUNKS=4 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 --no-other-mode --no-gather-obs 2>>$o/r2g_n8.err | tail -1 > $o/r2g_bench_n8_c4.json
python - <<'PY'
import json
for f in ("bound","unbound","c4"):
    try:
        d=json.load(open(f"gpurun_out/r2g_bench_n8_{f}.json")); print(f, d["value"]/1e6, d["e2e"])
    except Exception as e: print(f, "failed", e)
PY
cat $o/r2g_topo.txt | head -40
