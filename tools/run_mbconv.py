#!/usr/bin/env python
"""One MBConv forward + backward at a cfg 2 stage shape, alone (every kernel launched belongs to branch K3), for the
ncu DRAM-traffic pass behind bench.py's roofline.traffic:   python tools/run_mbconv.py [stage=0] [batch=1024]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import outlook_grid_vision_transformer_b200 as og  # noqa: E402

STAGES = {0: (64, 32), 1: (128, 16), 2: (256, 8), 3: (384, 4)}
stage = int(sys.argv[1]) if len(sys.argv) > 1 else 0
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
C, H = STAGES[stage]
dev = "cuda:0"
torch.manual_seed(0)
m = og.MBConv(C, C, 1, og.MBConvConfig()).to(dev).train()
x = torch.randn(B, C, H, H, device=dev).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
g = torch.randn(B, C, H, H, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
for it in range(3):
    torch.cuda.synchronize()
    print(f"ITER {it} FWD", flush=True)
    y = m(x)
    torch.cuda.synchronize()
    print(f"ITER {it} BWD", flush=True)
    y.backward(g)
    torch.cuda.synchronize()
    for p in m.parameters():
        p.grad = None
    x.grad = None
print("ok", float(y.float().abs().mean()))
