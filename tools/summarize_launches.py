#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name:
launch count, total time, share of the profiled region.  usage: summarize_launches.py launches.csv out.md"""
import csv
import collections
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
rows = []
with open(src, newline="") as fh:
    lines = [ln for ln in fh if not ln.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = r["Kernel Name"]
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    name = name.split("(")[0]
    rows.append((name[:110], us))
agg = collections.OrderedDict()
for n, us in rows:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
with open(dst, "w") as out:
    out.write(f"# ncu launch list summary ({src})\n\n")
    out.write(f"{len(rows)} launches, {tot/1e3:.3f} ms summed device time (cold-cache, serialised: compare SHARES)\n\n")
    out.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write(f"| `{n}` | {c} | {us/1e3:.3f} | {100*us/tot:.1f}% |\n")
print(f"{len(rows)} launches, {tot/1e3:.2f} ms -> {dst}")
