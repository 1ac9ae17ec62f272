#!/usr/bin/env python
"""Per-family summary of bench.py's per-kernel profile (gpurun_out/bench_kernels.json)."""
import collections
import json
import sys

d = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/bench_kernels.json"))
fam = collections.Counter()
for r in d:
    k = r["kernel"]
    f = "gemm fwd" if "gemm[fwd" in k else "gemm bwd" if "gemm[bwd" in k else "gemm wgrad" if "wgrad" in k else k
    fam[f] += r["ms_per_step"]
print(f"total {sum(fam.values()):.2f} ms/step")
for f, ms in fam.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 16):
    print(f"{f:32s} {ms:8.3f}")
