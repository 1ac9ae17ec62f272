import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch, torch.nn.functional as F
import outlook_grid_vision_transformer_b200 as og
from outlook_grid_vision_transformer_b200.engine import FlatState, TrainStep
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
from test_gpu_ddp_nccl import CFG, _shard
dev = torch.device("cuda", 0)
torch.manual_seed(100)
model = og.build_model(CFG).to(dev).train()
state = {k: v.detach().clone() for k, v in model.state_dict().items()}
ref = og.build_model(CFG).to(dev).train()
ref.load_state_dict(state)
x, y = (t.to(dev) for t in _shard(0))
for bf in (True, False):
    ref.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf):
        lg = ref(x)
    F.cross_entropy(lg.float(), y).backward()
    want = {k: p.grad.detach().float().clone() for k, p in ref.named_parameters()}
    m2 = og.build_model(CFG).to(dev).train()
    m2.load_state_dict(state)
    step = TrainStep(m2, lambda lg, yy: F.cross_entropy(lg, yy), x, y, lr=0.0, weight_decay=0.0, autocast_bf16=bf, use_graph=False)
    step(); step()
    torch.cuda.synchronize()
    bad = []
    for k, p in m2.named_parameters():
        e = float((p.grad.float() - want[k]).norm() / (want[k].norm() + 1e-9))
        if e > 2e-2: bad.append((k, round(e, 4), float(want[k].norm()), float(p.grad.float().norm())))
    print("bf16" if bf else "fp32", "bad:", bad)
