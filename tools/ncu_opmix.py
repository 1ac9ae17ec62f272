#!/usr/bin/env python
"""Opcode mix + stall summary from an `ncu --page source --csv` dump of ONE kernel.
    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv;  python tools/ncu_opmix.py src.csv [top]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops, stalls = defaultdict(float), defaultdict(float)
tot, samples = 0.0, 0.0
hot = []
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows[2:]:
    if len(r) < len(hdr) or not r[ix["Instructions Executed"]].replace(".", "").isdigit():
        continue
    src = r[ix["Source"]].strip()
    op = src.split()[0] if src else "?"
    if op.startswith("@"):
        op = src.split()[1]
    n = float(r[ix["Instructions Executed"]] or 0)
    s = float(r[ix["# Samples"]] or 0)
    ops[op.split(".")[0]] += n
    tot += n
    samples += s
    for c in stall_cols:
        stalls[c] += float(r[ix[c]] or 0)
    hot.append((s, n, r[ix["Address"]], src[:90], {c: float(r[ix[c]] or 0) for c in stall_cols}))
print(f"warp instructions executed: {tot:.0f}; samples {samples:.0f}")
for op, n in sorted(ops.items(), key=lambda kv: -kv[1])[:top]:
    print(f"  {op:14s} {n:14.0f} {100 * n / tot:5.1f}%")
print("stall samples:")
for c, n in sorted(stalls.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {c:28s} {n:10.0f} {100 * n / max(samples, 1):5.1f}%")
print("hottest instructions by samples:")
for s, n, a, src, st in sorted(hot, key=lambda t: -t[0])[:top]:
    why = max(st.items(), key=lambda kv: kv[1])
    print(f"  {s:8.0f} {n:12.0f} {a[-6:]} {src:90s} {why[0]}={why[1]:.0f}")
