"""gpurun_out/configs_*.jsonl (scripts/config_sweep.sh) -> markdown table for profiles/."""
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.strip().startswith("{")]
print("| workload | batch | dtype | value M env-steps/s | e2e M/s | ms/step | kernels ms (smooth / collision / constraint / solve) | roofline.fma frac | live tile (rows, overflow) | vs published |")
print("|---|---:|---|---:|---:|---:|---|---:|---|---:|")
for d in rows:
    c, r = d["config"], d["roofline"]
    k = r["kernel_ms_per_step"]
    pub = d.get("published_baseline")
    t = c.get("solve_live_tile")
    print(f"| {c['workload'].split(', batch')[0]} | {c['batch_per_gpu']} | {d['dtype']} | {d['value']/1e6:.2f} | {(d.get('e2e') or {}).get('value', 0)/1e6:.2f} | "
          f"{d['ms_per_step']:.3f} | {k.get('k_smooth', 0):.3f} / {k.get('k_collision', 0):.3f} / {k.get('k_constraint', 0):.3f} / {k.get('k_solve', 0):.3f} | "
          f"{(r.get('fma') or {}).get('frac', 0):.3f} | {str(t['rows']) + ', ' + format(t['overflow_frac_last_step'], '.3f') if t else '-'} | "
          f"{format(d['vs_baseline'], '.1f') + ' x (' + format(int(pub['value']), ',') + ')' if pub else '-'} |")
