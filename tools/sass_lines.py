#!/usr/bin/env python3
"""Per-source-line SASS instruction counts of one kernel in libplantos_b200.so (needs -lineinfo).
usage: tools/sass_lines.py <mangled-substring> [file-substring]"""
import collections, os, re, subprocess, sys, tempfile
so = os.environ.get("PLANTOS_LIB") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rl_env_b200/csrc/libplantos_b200.so")
pat = sys.argv[1]; fsub = sys.argv[2] if len(sys.argv) > 2 else "plantos_fast"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
inside = False; cur = None; counts = collections.Counter(); ops = collections.Counter(); total = 0
for l in txt:
    if l.startswith("//---") and ".text." in l:
        inside = pat in l; continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)', l)
    if m:
        total += 1; counts[cur] += 1; ops[m.group(2).split('.')[0]] += 1
print("total static instructions:", total)
for k, v in sorted(counts.items(), key=lambda kv: (kv[0] or ("", 0))):
    if k and fsub in k[0]: print(f"{k[0]}:{k[1]}\t{v}")
print(ops.most_common(25))
