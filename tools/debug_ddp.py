import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import torch, torch.distributed as dist, torch.nn.functional as F
import outlook_grid_vision_transformer_b200 as og
from outlook_grid_vision_transformer_b200 import functional as OF, modules as M_, ops
from outlook_grid_vision_transformer_b200.ddp import ArenaGradAllReduce, broadcast_parameters
from outlook_grid_vision_transformer_b200.engine import FlatState, TrainStep
from test_gpu_ddp_nccl import CFG, _shard
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
def P(*a):
    if rank == 0: print(*a, flush=True)
# 0. plain NCCL on a slice view
t = torch.full((1000,), float(rank + 1), device=dev)
dist.all_reduce(t[100:200]); torch.cuda.synchronize()
P("slice allreduce:", float(t[150]), float(t[50]))
w = dist.all_reduce(t[300:400], async_op=True); w.wait(); torch.cuda.synchronize()
P("async slice allreduce:", float(t[350]))
torch.manual_seed(100)
model = og.build_model(CFG).to(dev).train()
broadcast_parameters(model)
flat = FlatState(model)
sync = ArenaGradAllReduce(flat, bucket_bytes=64 << 10, tail_bytes=4 << 10)
P("world", sync.world, "buckets", [(b["lo"], b["hi"], len(b["params"])) for b in sync.buckets])
x, y = (t.to(dev) for t in _shard(rank))
# manual body
flat.zero_grad()
OF.SCRATCH = flat.scratch
with torch.autocast("cuda", dtype=torch.bfloat16):
    lg = model(x)
loss = F.cross_entropy(lg.float(), y)
loss.backward()
OF.SCRATCH = None
P("after backward: launched", [b["work"] is not None for b in sync.buckets], "pending", [b["pending"] for b in sync.buckets])
torch.cuda.synchronize()
own = flat.G.clone()
sync.finish()
torch.cuda.synchronize()
g1 = flat.G.clone()
other = own.clone(); dist.all_reduce(other); torch.cuda.synchronize()
P("|own|", float(own.norm()), "|after finish|", float(g1.norm()), "|true sum|", float(other.norm()),
  "err(after finish, true sum)", float((g1 - other).norm() / other.norm()))
dist.barrier(); dist.destroy_process_group()
