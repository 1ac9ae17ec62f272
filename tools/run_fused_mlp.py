#!/usr/bin/env python
"""Smallest program that launches the fused MLP kernels at the cfg 2 stage-0 shape (for ncu captures)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from outlook_grid_vision_transformer_b200 import ops  # noqa: E402

M, C, Hd = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576, int(sys.argv[2]) if len(sys.argv) > 2 else 64, \
    int(sys.argv[3]) if len(sys.argv) > 3 else 256
dev, dt = "cuda:0", torch.bfloat16
x = torch.randn(M, C, device=dev).to(dt)
dy = torch.randn(M, C, device=dev).to(dt)
w1 = (torch.randn(Hd, C, device=dev) * 0.05).to(dt)
w2 = (torch.randn(C, Hd, device=dev) * 0.05).to(dt)
b1 = torch.zeros(Hd, device=dev)
b2 = torch.zeros(C, device=dev)
for _ in range(3):
    y = ops.mlp_fwd(x, w1, b1, w2, b2, act="gelu", residual=dy)
    out = ops.mlp_bwd(x, dy, w1, w2.t().contiguous(), w1.t().contiguous(), b1, act="gelu")
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()), float(out[0].float().abs().mean()))
