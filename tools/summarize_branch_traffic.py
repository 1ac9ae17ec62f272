#!/usr/bin/env python
"""DRAM traffic of a whole fused BRANCH (all its launches summed) from a metrics-only ncu pass over tools/run_mbconv.py:
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/<tag>_k3_traffic.csv python tools/run_mbconv.py 0 1024
    python tools/summarize_branch_traffic.py gpurun_out/<tag>_k3_traffic.csv <tag> K3 1048576 64 <fwd launches> <bwd launches>
The last (fwd + bwd) launches of the run are the third iteration; its first `fwd launches` are the forward.
Writes profiles/<tag>_k3_traffic.md and merges "K3 fwd M=.. C=.." / "K3 bwd .." into profiles/traffic.json."""
import csv
import json
import re
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def short(name):
    m = re.search(r"(\w+<[^>]*>|\w+)\(", name.replace("<unnamed>::", ""))
    return (m.group(1) if m else name)[:60]


def main():
    path, tag, branch, M, C = Path(sys.argv[1]), sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    per, order = defaultdict(dict), []
    for r in csv.DictReader(lines):
        i = int(r["ID"])
        if i not in per:
            order.append(i)
        per[i]["name"] = short(r["Kernel Name"])
        per[i][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    rows = [per[i] for i in order]
    # the third iteration: forward = from after the previous iteration's last weight-gradient GEMM up to the last
    # bn_apply; backward = from there to the end, minus the three torch kernels of the script's final print
    names = [r["name"] for r in rows]
    while names and not (names[-1].startswith("gemm_tc_kernel") or names[-1].startswith("bn_") or names[-1].startswith("dw")):
        rows.pop()
        names.pop()
    fwd_end = max(i for i, nm in enumerate(names) if nm.startswith("bn_apply_kernel")) + 1
    fwd_start = max(i for i, nm in enumerate(names[:fwd_end]) if nm.startswith("gemm_tc_kernel<64, float")) + 1
    last = rows[fwd_start:]
    split = fwd_end - fwd_start
    elt = 2
    out = [f"# {tag}: DRAM traffic of branch {branch} (MBConv) at M={M}, C={C}, bf16 -- every launch of one forward and one backward\n",
           f"source: `{path.name}` (ncu metrics pass over `tools/run_mbconv.py`, third iteration; cold-cache serialised launches)\n"]
    tj_path = ROOT / "profiles" / "traffic.json"
    tj = json.loads(tj_path.read_text()) if tj_path.exists() else {}
    for direction, part, alg in (("fwd", last[:split], elt * 2 * M * C), ("bwd", last[split:], elt * 3 * M * C)):
        tot = sum(r.get("dram__bytes_read.sum", 0) + r.get("dram__bytes_write.sum", 0) for r in part)
        ns = sum(r.get("gpu__time_duration.sum", 0) for r in part)
        out += [f"\n## {direction}: {len(part)} launches, {tot / 1e9:.3f} GB of DRAM traffic = {tot / alg:.1f} x the algorithmic "
                f"{alg / 1e6:.0f} MB (SURVEY 8(d)), {ns / 1e3:.0f} us under ncu\n",
                "| kernel | DRAM read MB | DRAM write MB | ncu us |", "|---|---:|---:|---:|"]
        for r in part:
            out.append(f"| `{r['name']}` | {r.get('dram__bytes_read.sum', 0) / 1e6:.1f} | {r.get('dram__bytes_write.sum', 0) / 1e6:.1f} | "
                       f"{r.get('gpu__time_duration.sum', 0) / 1e3:.1f} |")
        tj[f"{branch} {direction} M={M} C={C}"] = {"dram_bytes_per_instance": round(tot), "launches": len(part),
                                                      "algorithmic_bytes": alg, "source": f"profiles/{tag}_k3_traffic.md"}
    (ROOT / "profiles" / f"{tag}_k3_traffic.md").write_text("\n".join(out) + "\n")
    tj_path.write_text(json.dumps(tj, indent=1) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
