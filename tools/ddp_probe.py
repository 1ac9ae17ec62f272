"""The multi-GPU tax of the train step, taken apart (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/ddp_probe.py

  * replayed step WITH the arena all-reduce vs. the same replicas WITHOUT any exchange (all N GPUs busy in both);
  * eager step with CUDA events around ArenaGradAllReduce.finish(): how long the main stream waits for the comm stream
    after backward's last kernel (the exposed tail of the all-reduce), and the bucket layout.
Rank 0 prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_MAX_CTAS", "8")
    dist.init_process_group("nccl", device_id=dev)
    sys.argv = sys.argv[:1]
    args = bench.parse_args()
    out = {"world": world}

    w = bench.Workload("cfg2_14m_32_bf16", args, rank, world, dev, use_graph=True, warm=3)
    for _ in range(5):
        w.step_resident()
    out["graph_with_allreduce_ms"] = bench.timed(w.step_resident, 20, world, dev)
    sync = w.sync
    out["buckets_mb"] = [round((b["hi"] - b["lo"]) * 4 / 2 ** 20, 2) for b in sync.buckets]
    w.close()
    del w

    # the same replicas without any exchange: world = 1 for the workload, all GPUs still busy
    w1 = bench.Workload("cfg2_14m_32_bf16", args, rank, 1, dev, use_graph=True, warm=3)
    for _ in range(5):
        w1.step_resident()
    out["graph_without_exchange_ms"] = bench.timed(w1.step_resident, 20, world, dev)
    w1.close()
    del w1

    # eager, events around finish()
    we = bench.Workload("cfg2_14m_32_bf16", args, rank, world, dev, use_graph=False, warm=2)
    waits = []
    orig = we.sync.finish

    def finish():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig()
        e1.record()
        waits.append((e0, e1))
    we.sync.finish = finish
    for _ in range(3):
        we.step_resident()
    waits.clear()
    out["eager_with_allreduce_ms"] = bench.timed(we.step_resident, 10, world, dev)
    torch.cuda.synchronize()
    ws = sorted(a.elapsed_time(b) for a, b in waits)
    out["exposed_allreduce_wait_ms"] = {"median": ws[len(ws) // 2], "min": ws[0], "max": ws[-1], "steps": len(ws)}
    t = torch.tensor([out["exposed_allreduce_wait_ms"]["median"]], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out["exposed_allreduce_wait_ms"]["median_max_over_ranks"] = float(t)
    we.close()
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    os._exit(0)


if __name__ == "__main__":
    main()
