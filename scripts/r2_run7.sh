#!/bin/bash
# validation pass: whole GPU suite, then the default bench line and the f32 one
o=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > $o/r2j_gpu_tests.log 2>&1; tail -5 $o/r2j_gpu_tests.log
timeout 900 python bench.py 2>$o/r2j_bench.err | tail -1 > $o/r2j_bench_f64.json
timeout 600 python bench.py --dtype f32 --no-cpu-baseline 2>>$o/r2j_bench.err | tail -1 > $o/r2j_bench_f32.json
python - <<PY
import json
for f in ("r2j_bench_f64", "r2j_bench_f32"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"] / 1e6, 2), "M", round(d["ms_per_step"], 3), "ms", {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()},
              "e2e", d["e2e"] and round(d["e2e"]["value"] / 1e6, 2), "other", d["other_output_mode"] and round(d["other_output_mode"]["value"] / 1e6, 2))
    except Exception as e:
        print(f, "bad", e)
PY
