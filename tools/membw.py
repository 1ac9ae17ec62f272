#!/usr/bin/env python
"""Raw HBM ceilings for context: write-only (fill), read-only (sum) and copy streams over 1 GiB."""
import torch
n = 1 << 29  # bf16 elements = 1 GiB
x = torch.empty(n, device="cuda", dtype=torch.bfloat16)
y = torch.empty(n, device="cuda", dtype=torch.bfloat16)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
gb = n * 2 / 1e9
ms = t(lambda: x.fill_(1.0)); print(f"fill  (write only): {gb / ms * 1e3:8.1f} GB/s")
ms = t(lambda: x.view(torch.int32).sum()); print(f"sum   (read only) : {gb / ms * 1e3:8.1f} GB/s")
ms = t(lambda: y.copy_(x)); print(f"copy  (read+write): {2 * gb / ms * 1e3:8.1f} GB/s")
ms = t(lambda: torch.add(x, x, out=y)); print(f"add   (2R+1W... x twice = 1R+1W): {2 * gb / ms * 1e3:8.1f} GB/s")
