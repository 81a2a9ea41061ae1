"""Attribute ncu SASS-level samples/instruction counts to CUDA source lines.
usage: ncu_lines.py <rep> <kernel-regex> <mangled-function-substring> [topN]
Joins `ncu --page source --csv` (per SASS instruction, in order) with `nvdisasm -g` line markers (same order)."""
import csv, re, subprocess, sys, os, glob, tempfile
rep, kre, fn = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mujoco_torch_b200", "csrc", "libmjt_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
sass, start = None, None
for cubin in sorted(glob.glob(tmp + "/*.cubin")):   # one cubin per translation unit: find the one that holds the function
    cand = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    hit = [i for i, l in enumerate(cand) if l.startswith(".text.") and fn in l and l.rstrip().endswith(":")]
    if hit:
        sass, start = cand, hit[0]
        break
assert sass is not None, f"no function matching {fn!r} in {so}"
lines, cur = [], None
for l in sass[start + 1:]:
    if l.startswith("//---------------------") or (l.startswith(".text.") and l.rstrip().endswith(":")):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(cur)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
agg = {}
n = 0
active, h, ii, wi = False, None, None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        if active and n:
            break
        active = re.search(kre, r[1]) is not None
        continue
    if not active:
        continue
    if "Source" in r and "Address" in r:
        h = r
        ii, wi = h.index("Instructions Executed"), h.index("# Samples")
        continue
    if h is None or len(r) <= max(ii, wi) or not r[ii].isdigit():
        continue
    key = lines[n] if n < len(lines) else None
    n += 1
    a = agg.setdefault(key, [0, 0])
    a[0] += int(r[ii]); a[1] += int(r[wi])
print("sass instrs in csv", n, "in nvdisasm", len(lines))
ti, ts = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
src_cache = {}
def src(key):
    if key is None: return "?"
    f, ln = key
    p = os.path.join(os.path.dirname(so), f)
    if p not in src_cache:
        src_cache[p] = open(p).read().splitlines() if os.path.exists(p) else []
    s = src_cache[p]
    return s[ln - 1].strip()[:100] if 0 < ln <= len(s) else ""
print(f"total warp-instr {ti}  samples {ts}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{a[0]/ti*100:5.1f}% inst {a[1]/ts*100:5.1f}% samp  {key}  {src(key)}")
