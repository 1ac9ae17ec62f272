"""World-size-2 test of the data-parallel exchange step (ddp.BucketedGradAllReduce) on CPU / gloo.

The product blocks are CUDA-only, so the replicas here carry a small plain-torch network; what is
tested is the host logic of the path's only collective: bucketing in reverse parameter order,
hook-driven launch, averaging, re-arming across steps, parameters that receive no gradient, and
`broadcast_parameters`.  The DDP contract: all-reduced grad == mean over ranks of the
per-shard gradients, each replica keeping its own BatchNorm statistics.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _net():
    torch.manual_seed(1234)
    net = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.BatchNorm2d(8), nn.SiLU(), nn.Conv2d(8, 8, 1),
                        nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(8, 5))
    net.unused = nn.Parameter(torch.ones(3))  # never receives a gradient
    return net


def _shard(rank: int):
    g = torch.Generator().manual_seed(100 + rank)
    return torch.randn(6, 3, 8, 8, generator=g), torch.randint(0, 5, (6,), generator=g)


def _local_grads(rank: int, steps: int):
    """Gradients a single replica computes on its own shard (the per-shard reference)."""
    net = _net()
    out = []
    for _ in range(steps):
        net.zero_grad(set_to_none=True)
        x, y = _shard(rank)
        nn.functional.cross_entropy(net(x), y).backward()
        out.append({k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None})
    return out, {k: v.clone() for k, v in net.state_dict().items() if "running" in k}


def _worker(rank: int, world: int, port: int, bucket_bytes: int, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from outlook_grid_vision_transformer_b200.ddp import BucketedGradAllReduce, broadcast_parameters

        net = _net()
        if rank == 1:  # perturb, then prove broadcast restores rank 0's values
            with torch.no_grad():
                for p in net.parameters():
                    p.add_(1.0)
        broadcast_parameters(net)
        ref = _net()
        for (k, a), (_, b) in zip(net.state_dict().items(), ref.state_dict().items()):
            assert torch.equal(a, b), f"broadcast_parameters: {k} differs on rank {rank}"
        sync = BucketedGradAllReduce(net.parameters(), bucket_bytes=bucket_bytes)
        got = []
        for _ in range(2):  # two steps: the hook counters must re-arm
            net.zero_grad(set_to_none=True)
            x, y = _shard(rank)
            nn.functional.cross_entropy(net(x), y).backward()
            sync.finish()
            got.append({k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None})
        bn = {k: v.clone() for k, v in net.state_dict().items() if "running" in k}
        # plain numpy payloads: torch tensors would travel as shared-memory fds that die with the worker
        q.put((rank, [{k: v.numpy() for k, v in g.items()} for g in got], {k: v.numpy() for k, v in bn.items()},
               len(sync.buckets)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("bucket_bytes", [256, 25 * 1024 * 1024], ids=["many_buckets", "one_bucket"])
def test_allreduced_grads_equal_mean_of_shard_grads(bucket_bytes):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, bucket_bytes, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    for _ in range(world):
        rank, got, bn, nb = q.get(timeout=180)
        results[rank] = ([{k: torch.from_numpy(v) for k, v in g.items()} for g in got],
                         {k: torch.from_numpy(v) for k, v in bn.items()}, nb)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    steps = 2
    local = {r: _local_grads(r, steps) for r in range(world)}
    if bucket_bytes == 256:
        assert results[0][2] > 1, "small bucket size should split the parameters into several buckets"
    for s in range(steps):
        for k in local[0][0][s]:
            want = sum(local[r][0][s][k] for r in range(world)) / world
            for r in range(world):
                torch.testing.assert_close(results[r][0][s][k], want, rtol=1e-5, atol=1e-6,
                                           msg=lambda m: f"step {s} rank {r} grad[{k}]: {m}")
    # the never-used parameter gets a zero gradient from finish() and stays zero after the reduce
    for r in range(world):
        assert torch.count_nonzero(results[r][0][0]["unused"]) == 0
    # BatchNorm running statistics stay per replica (no SyncBN in the reference)
    for r in range(world):
        for k, v in local[r][1].items():
            torch.testing.assert_close(results[r][1][k], v)
    assert not torch.allclose(results[0][1]["1.running_mean"], results[1][1]["1.running_mean"])


# ------------------------------------------------------------------------------------------------
# the arena form used by engine.TrainStep: all-reduce in place on slices of FlatState's gradient arena
# ------------------------------------------------------------------------------------------------
def _arena_worker(rank: int, world: int, port: int, bucket_bytes: int, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from outlook_grid_vision_transformer_b200.ddp import ArenaGradAllReduce
        from outlook_grid_vision_transformer_b200.engine import FlatState

        net = _net()
        flat = FlatState(net, scratch_floats=64)
        sync = ArenaGradAllReduce(flat, bucket_bytes=bucket_bytes, tail_bytes=64)
        got = []
        for _ in range(2):
            flat.zero_grad()
            x, y = _shard(rank)
            nn.functional.cross_entropy(net(x), y).backward()
            sync.finish()
            # the arena holds the SUM; the optimizer kernel applies 1/world (TrainStep(world=...))
            got.append({k: (p.grad / world).clone().numpy() for k, p in net.named_parameters()})
        covered = sorted((b["lo"], b["hi"]) for b in sync.buckets)
        q.put((rank, got, len(sync.buckets), covered, flat.n))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("bucket_bytes", [128, 8 << 20], ids=["many_buckets", "one_bucket"])
def test_arena_allreduce_equals_mean_of_shard_grads(bucket_bytes):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_arena_worker, args=(r, world, port, bucket_bytes, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    for _ in range(world):
        rank, got, nb, covered, n = q.get(timeout=180)
        results[rank] = (got, nb, covered, n)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    local = {r: _local_grads(r, 2) for r in range(world)}
    got, nb, covered, n = results[0]
    if bucket_bytes == 128:
        assert nb > 2
    # the buckets tile the arena exactly once
    assert covered[0][0] == 0 and covered[-1][1] == n and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    for s in range(2):
        for k in local[0][0][s]:
            want = sum(local[r][0][s][k] for r in range(world)) / world
            for r in range(world):
                torch.testing.assert_close(torch.from_numpy(results[r][0][s][k]), want, rtol=1e-5, atol=1e-6,
                                           msg=lambda m: f"step {s} rank {r} grad[{k}]: {m}")
        for r in range(world):  # the never-used parameter keeps a zero gradient
            assert not results[r][0][s]["unused"].any()
