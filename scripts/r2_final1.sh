#!/bin/bash
# final single-GPU evidence: smoke, the default bench line, the reference arm, the ncu launch list and one full capture of a step
o=gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $o/r2n_smoke.log 2>&1; tail -2 $o/r2n_smoke.log
timeout 900 python bench.py 2>$o/r2n_bench.err | tail -1 > $o/r2n_bench_f64.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>>$o/r2n_bench.err | tail -1 > $o/r2n_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/r2n_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-other-mode --no-variants > $o/r2n_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ -c 4 -f -o $o/r2n_state_f64 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-other-mode --no-variants > $o/r2n_ncu.log 2>&1
ls -la $o | grep r2n
python - <<PY
import json
for f in ("r2n_bench_f64", "r2n_bench_reference"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d.get("impl"), round(d["value"], 1), d["unit"], round(d["ms_per_step"], 3), "ms", d.get("e2e") and round(d["e2e"]["value"], 1))
    except Exception as e:
        print(f, "bad", e)
PY
