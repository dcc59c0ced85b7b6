#!/usr/bin/env python
"""Instruction mix of the Sinkhorn loop of pair_fused_kernel<false, 1> in a built binary (cuobjdump -sass).
usage: loop_sass.py <binary> [--dump]"""
import re, subprocess, sys, collections
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", out)
body = next(f for f in funcs if f.startswith("_ZN2vr17pair_fused_kernelILb0ELi1"))
ins = []
for m in re.finditer(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", body):
    ins.append((int(m.group(1), 16), m.group(2).strip()))
best = None
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA\S*\s+.*?(0x[0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        tgt = int(m.group(1), 16)
        span = [x for x in ins if tgt <= x[0] <= a]
        n_ldtm = sum("LDTM" in x[1] for x in span)
        if n_ldtm >= 13 and (best is None or len(span) < len(best)):
            best = span
print("loop instructions:", len(best))
h = collections.Counter()
for a, t in best:
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    h[t.split()[0].split(".")[0]] += 1
print(", ".join(f"{k} {v}" for k, v in h.most_common()))
if "--dump" in sys.argv:
    for a, t in best:
        print(f"{a:06x}  {t}")
