#!/usr/bin/env python
"""Markdown table of the `ncu --set full` capture of the dominant kernel (raw page CSV from tools/profile_round.sh).

    python tools/summarize_top.py gpurun_out/r01e_top.raw.csv > profiles/r01e_top_gemm_tc.md
"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
H = {h: i for i, h in enumerate(hdr)}
cols = [("Grid Size", "grid"), ("gpu__time_duration.sum", "time µs"), ("dram__bytes_read.sum", "DRAM read MB"),
        ("dram__bytes_write.sum", "DRAM write MB"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe %"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "ALU pipe %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts")]
print("| # | kernel | " + " | ".join(c[1] for c in cols) + " | top stalls (warps per issue) |")
print("|---|---|" + "---:|" * len(cols) + "---|")
for n, r in enumerate(rows[2:]):
    nm = re.sub(r"\(.*", "", r[H["Kernel Name"]]).replace("void <unnamed>::", "").replace("__nv_bfloat16", "bf16")
    vals = []
    for c, _ in cols:
        v = r[H[c]]
        if c == "Grid Size":
            v = v.strip("()").split(",")[0]
        else:
            try:
                v = f"{float(v):.1f}" if "." in v else v
            except ValueError:
                pass
        vals.append(v)
    st = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(r[H[h]])
          for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")}
    top = ", ".join(f"{k} {v:.1f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"| {n} | `{nm}` | " + " | ".join(vals) + f" | {top} |")
