"""Debug aid: the NCCL arena all-reduce test body with every bad parameter printed (run under gpurun --gpus 2)."""
import os
import socket
import sys

import torch
import torch.multiprocessing as mp

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import test_gpu_ddp_nccl as T  # noqa: E402


def worker(rank, world, port, use_graph, q):
    import torch.distributed as dist
    import torch.nn.functional as F

    import outlook_grid_vision_transformer_b200 as og
    from outlook_grid_vision_transformer_b200.ddp import ArenaGradAllReduce, broadcast_parameters
    from outlook_grid_vision_transformer_b200.engine import FlatState, TrainStep

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        torch.manual_seed(100 + rank)
        model = og.build_model(T.CFG).to(dev).train()
        broadcast_parameters(model)
        state = {k: v.detach().clone() for k, v in model.state_dict().items()}
        ref = og.build_model(T.CFG).to(dev).train()
        want = None
        per = []
        for r in range(world):
            ref.load_state_dict(state)
            ref.zero_grad(set_to_none=True)
            x, y = (t.to(dev) for t in T._shard(r))
            with torch.autocast("cuda", dtype=torch.bfloat16):
                lg = ref(x)
            F.cross_entropy(lg.float(), y).backward()
            g = {k: p.grad.detach().float().clone() for k, p in ref.named_parameters()}
            per.append(g)
            want = g if want is None else {k: want[k] + g[k] for k in g}
        want = {k: v / world for k, v in want.items()}
        flat = FlatState(model)
        sync = ArenaGradAllReduce(flat, bucket_bytes=32 << 10, tail_bytes=4 << 10)
        x, y = (t.to(dev) for t in T._shard(rank))
        step = TrainStep(model, lambda lg, yy: F.cross_entropy(lg, yy), x, y, lr=0.0, weight_decay=0.0, autocast_bf16=True,
                         grad_sync=sync, use_graph=use_graph, warmup=2, world=world, flat=flat)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        bad = []
        for k, p in model.named_parameters():
            got = p.grad.detach().float() / world
            err = float((got - want[k]).norm() / (want[k].norm() + 1e-6 * want[k].numel() ** 0.5))
            own = float((p.grad.detach().float() - per[rank][k]).norm() / (per[rank][k].norm() + 1e-9))
            if err > 2e-2:
                bad.append((k, round(err, 3), "vs own-shard-only", round(own, 3)))
        q.put((rank, bad, [(b["lo"], b["hi"], len(b["params"])) for b in sync.buckets]))
        sync.remove()
        step.graph = None
        del step
        torch.cuda.synchronize()
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    use_graph = len(sys.argv) > 1 and sys.argv[1] == "graph"
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, use_graph, q)) for r in range(2)]
    for p in procs:
        p.start()
    for _ in range(2):
        rank, bad, buckets = q.get(timeout=300)
        print("rank", rank, "bad:", bad[:12], "n_bad", len(bad), "buckets", buckets)
    for p in procs:
        p.join(timeout=120)
