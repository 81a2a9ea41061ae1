#!/bin/bash
# last validation of the round: whole GPU suite, smoke, default / f32 / reference bench lines
o=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > $o/r2s_gpu_tests.log 2>&1; tail -3 $o/r2s_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $o/r2s_smoke.log 2>&1; tail -1 $o/r2s_smoke.log
timeout 900 python bench.py 2>$o/r2s_bench.err | tail -1 > $o/r2s_bench_f64.json
timeout 600 python bench.py --dtype f32 --no-cpu-baseline 2>>$o/r2s_bench.err | tail -1 > $o/r2s_bench_f32.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>>$o/r2s_bench.err | tail -1 > $o/r2s_bench_reference.json
python - <<PY
import json
for f in ("r2s_bench_f64", "r2s_bench_f32", "r2s_bench_reference"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"] / 1e6, 3), "M", round(d["ms_per_step"], 3), "ms", d.get("roofline") and {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()},
              "e2e", d["e2e"] and round(d["e2e"]["value"] / 1e6, 2))
    except Exception as e:
        print(f, "bad", e)
PY
