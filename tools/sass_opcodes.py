#!/usr/bin/env python3
"""Opcode histogram per kernel of the shipped library (static SASS counts) -> profiles/<tag>_sass_opcodes.txt.
The memory-path opcodes that matter for the claims in DESIGN.md are listed first for every kernel:
UBLKCP (cp.async.bulk, the TMA bulk copy), SYNCS (mbarrier), LDGSTS (cp.async), LDG/STG .ENL2.256 (256-bit record access),
STG.E.EF.128 (streaming float4 observation stores), LDL/STL (spills), ACQBULK/ERRBAR/MEMBAR (fences).
usage: tools/sass_opcodes.py <tag>"""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "rl_env_b200/csrc/libplantos_b200.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout.split("\n")
kern, hist = None, collections.OrderedDict()
for l in txt:
    m = re.search(r"Function : (\S+)", l)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = ("UBLKCP", "SYNCS", "LDGSTS", "ENL2.256", "STG.E.EF.128", "LDL", "STL", "MEMBAR", "ERRBAR", "ACQBULK", "FLO", "PRMT", "NANOSLEEP",
       "REDG", "ATOMG", "MATCH", "VOTE", "SHFL", "LDS", "STS", "LDG", "STG")
out = [f"# static SASS opcode counts per kernel of rl_env_b200/csrc/libplantos_b200.so (cuobjdump -sass, CUDA 12.9, sm_100a); tools/sass_opcodes.py", ""]
for k, h in hist.items():
    tot = sum(h.values())
    out.append(f"== {k}   ({tot} instructions)")
    keyed = []
    for key in KEY:
        n = sum(v for op, v in h.items() if key in op.split(".")[0] or (("." in key) and key in op))
        if n:
            keyed.append(f"{key}={n}")
    out.append("   memory / sync path: " + "  ".join(keyed))
    fam = collections.Counter()
    for op, v in h.items():
        fam[op.split(".")[0]] += v
    out.append("   families: " + "  ".join(f"{op}={v}" for op, v in fam.most_common(24)))
    full = [f"{op}={v}" for op, v in sorted(h.items()) if any(s in op for s in ("UBLKCP", "SYNCS", "LDGSTS", "ENL2", "STG.E.EF", "LDG.E", "LDL", "STL"))]
    out.append("   exact: " + "  ".join(full))
    out.append("")
path = os.path.join(root, "profiles", f"{tag}_sass_opcodes.txt")
open(path, "w").write("\n".join(out))
print(path, len(hist), "kernels")
