#!/usr/bin/env python3
"""Annotated SASS (source line per instruction) of one kernel. usage: sass_dump.py <mangled-substring>"""
import os, re, subprocess, sys, tempfile
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rl_env_b200/csrc/libplantos_b200.so")
pat = sys.argv[1]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
inside = False; cur = None
for l in txt:
    if l.startswith("//---") and ".text." in l:
        inside = pat in l; continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)).replace("plantos_", "").replace(".cuh", ""), int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: print(m.group(1), f"{cur[0]}:{cur[1]}" if cur else "", m.group(2).strip())
    elif re.match(r'\.L_x_\d+:', l): print(l.strip())
