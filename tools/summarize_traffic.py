#!/usr/bin/env python
"""Per-launch DRAM traffic of the dominant kernel from the metrics-only ncu pass of tools/profile_round.sh
(dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum per launch).

    python tools/summarize_traffic.py gpurun_out/r01d_top_traffic.csv r01d [steps in the profiled run = 7]

Writes profiles/<tag>_top_traffic.md and merges {kernel name -> mean bytes per launch} into
profiles/traffic.json (bench.py's roofline.traffic looks the dominant kernel up there by name).
Only the launches of the LAST bench step are used (the first launches are the warm-up steps)."""
import csv
import json
import re
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def short(name):
    m = re.search(r"(\w+<[^>]*>|\w+)\(", name.replace("<unnamed>::", ""))
    return m.group(1) if m else name


def main():
    path, tag = Path(sys.argv[1]), sys.argv[2]
    # bench --steps 1 --warmup 3 runs 7 identical steps: 3 warm-up + 1 timed, then the e2e leg's 2 + 1
    steps_in_run = int(sys.argv[3]) if len(sys.argv) > 3 else 7
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    per = defaultdict(dict)
    order = []
    for r in csv.DictReader(lines):
        i = int(r["ID"])
        if i not in per:
            order.append(i)
        per[i]["name"] = short(r["Kernel Name"])
        per[i]["grid"] = r["Grid Size"]
        per[i][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    rows = [per[i] for i in order]
    n = len(rows)
    assert n % steps_in_run == 0, f"{n} launches do not divide into {steps_in_run} identical steps"
    last = rows[n - n // steps_in_run:]
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for r in last:
        a = agg[r["name"]]
        a[0] += 1
        a[1] += r.get("dram__bytes_read.sum", 0.0)
        a[2] += r.get("dram__bytes_write.sum", 0.0)
        a[3] += r.get("gpu__time_duration.sum", 0.0)
    out = [f"# {tag}: DRAM traffic per launch of the dominant kernel family (ncu metrics pass, last bench step)\n",
           f"source: `{path.name}` ({n} launches profiled, {len(last)} in the last step); cold-cache serialised "
           "launches, clocks not locked\n",
           "| kernel | launches/step | DRAM read MB/launch | DRAM write MB/launch | total MB/launch | ncu µs/launch | DRAM GB/s under ncu |",
           "|---|---|---|---|---|---|---|"]
    tj_path = ROOT / "profiles" / "traffic.json"
    tj = json.loads(tj_path.read_text()) if tj_path.exists() else {}
    for name, (c, rd, wr, ns) in sorted(agg.items(), key=lambda kv: -kv[1][3]):
        tot = (rd + wr) / c
        out.append(f"| `{name}` | {c} | {rd / c / 1e6:.1f} | {wr / c / 1e6:.1f} | {tot / 1e6:.1f} | {ns / c / 1e3:.1f} | "
                   f"{(rd + wr) / ns:.0f} |")
        tj[name] = {"dram_bytes_per_launch": round(tot), "launches_per_step": c, "source": f"profiles/{tag}_top_traffic.md"}
    (ROOT / "profiles" / f"{tag}_top_traffic.md").write_text("\n".join(out) + "\n")
    tj_path.write_text(json.dumps(tj, indent=1) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
