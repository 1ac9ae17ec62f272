"""Debug aid: gradients of engine.TrainStep (flat arenas, direct accumulation) vs plain autograd on the same batch, bf16."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import outlook_grid_vision_transformer_b200 as og  # noqa: E402
from outlook_grid_vision_transformer_b200.engine import FlatState, TrainStep  # noqa: E402

CFG = {"type": "model_a", "num_classes": 10, "stem_dim": 16, "dpr_max": 0.0,
       "stages": [dict(dim=16, depth=1, num_heads=2, grid_size=2, outlook_heads=2),
                  dict(dim=32, depth=2, num_heads=2, grid_size=2, outlook_heads=2)]}
dev = "cuda:0"
use_graph = len(sys.argv) > 1 and sys.argv[1] == "graph"
g = torch.Generator().manual_seed(50)
x, y = torch.randn(8, 3, 8, 8, generator=g).to(dev), torch.randint(0, 10, (8,), generator=g).to(dev)
torch.manual_seed(100)
model = og.build_model(CFG).to(dev).train()
state = {k: v.detach().clone() for k, v in model.state_dict().items()}
ref = og.build_model(CFG).to(dev).train()
wants = []
for rep in range(2):
    ref.load_state_dict(state)
    ref.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        lg = ref(x)
    F.cross_entropy(lg.float(), y).backward()
    wants.append({k: p.grad.detach().float().clone() for k, p in ref.named_parameters()})
flat = FlatState(model)
step = TrainStep(model, lambda lg, yy: F.cross_entropy(lg, yy), x, y, lr=0.0, weight_decay=0.0, autocast_bf16=True,
                 use_graph=use_graph, warmup=2, flat=flat)
for it in range(2):
    step()
    torch.cuda.synchronize()
    rows = []
    for k, p in model.named_parameters():
        w = wants[0][k]
        err = float((p.grad.float() - w).norm() / (w.norm() + 1e-6 * w.numel() ** 0.5))
        noise = float((wants[1][k] - w).norm() / (w.norm() + 1e-6 * w.numel() ** 0.5))
        rows.append((err, noise, k))
    rows.sort(reverse=True)
    print(f"step {it}: worst", [(k, round(e, 4), "autograd-vs-autograd", round(n, 4)) for e, n, k in rows[:6]])
