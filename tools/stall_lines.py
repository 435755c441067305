#!/usr/bin/env python3
"""Join an `ncu --page source --csv` dump (SASS rows) with nvdisasm line info: per source line
warp-instructions executed, stall samples and the dominant stall reasons.
usage: tools/stall_lines.py <source.csv> <mangled-substring> [min-samples]"""
import collections, csv, os, re, subprocess, sys, tempfile
so = os.environ.get("PLANTOS_LIB") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rl_env_b200/csrc/libplantos_b200.so")
src, pat = sys.argv[1], sys.argv[2]
mins = int(sys.argv[3]) if len(sys.argv) > 3 else 0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
inside = False; cur = None; line_of = {}
for l in txt:
    if l.startswith("//---") and ".text." in l:
        inside = pat in l; continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+', l)
    if m: line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
a0 = int(data[0][0], 16)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for r in data:
    off = int(r[0], 16) - a0
    k = line_of.get(off)
    a = agg[k]
    a["inst"] += int(r[ix["Instructions Executed"]] or 0)
    a["samples"] += int(r[ix["# Samples"]] or 0)
    for s in stall_cols:
        v = int(r[ix[s]] or 0); a[s] += v; tot[s] += v
    tot["inst"] += int(r[ix["Instructions Executed"]] or 0); tot["samples"] += int(r[ix["# Samples"]] or 0)
print("total inst", tot["inst"], "samples", tot["samples"], {s: tot[s] for s in stall_cols if tot[s] * 50 > tot["samples"]})
for k in sorted(agg, key=lambda k: (k or ("", 0))):
    a = agg[k]
    if a["samples"] < mins: continue
    top = sorted(((a[s], s[6:]) for s in stall_cols if a[s]), reverse=True)[:3]
    print(f"{k[0] if k else '?'}:{k[1] if k else 0}\tinst {a['inst']:>8}\tsamples {a['samples']:>6} ({100*a['samples']/tot['samples']:.1f}%)\t" + " ".join(f"{n}:{v}" for v, n in top))
