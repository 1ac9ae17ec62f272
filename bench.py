#!/usr/bin/env python
"""Headline benchmark: OutGridViT Model A train-step images/s (BASELINE.json configs[1]:
14M, 32x32, batch 1024 per GPU, bf16 autocast), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--batch B]

`value`     : whole-job images/s with the batch already resident in HBM (forward + backward +
              gradient all-reduce (N>1) + AdamW step), CUDA-event timed, max over ranks.
`e2e`       : same step driven from pinned HOST buffers (H2D of images+labels and D2H of the loss
              inside the timed region).
`roofline`  : dominant kernel of the step (by device time, measured live with CUDA events on the
              launching stream in a separate profiled pass), algorithmic bytes or flops per launch
              over its mean duration, against MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference`: the CPU oracle port of the reference model (oracle/), timed on
              the host cores on a bounded sample of the same workload -- a reported baseline only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "train_step_images_per_sec"
UNIT = "img/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2_14m_32_bf16")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of replaying a CUDA graph")
    return ap.parse_args()


def workload_desc(name, wl, batch):
    return (f"OutGridViT {wl['yaml']} {wl['mode']} step, {wl['img']}x{wl['img']}, batch {batch}/GPU, {wl['dtype']}")


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=float(d["hbm_gbs"]), tc=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tc=1590.0, src="fallback")


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference model on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_train_step_rate(model_cfg, img, n_img, steps, warmup, seed=7):
    import torch

    from oracle import outgrid_oracle as O

    torch.manual_seed(seed)
    params = O.init_params(model_cfg, seed=seed)
    leaves = {k: v.requires_grad_(True) for k, v in params.items() if v.is_floating_point() and "running_" not in k}
    opt = torch.optim.AdamW(list(leaves.values()), lr=5e-4, weight_decay=0.05)
    x = torch.randn(n_img, 3, img, img)
    y = torch.randint(0, int(model_cfg.get("num_classes", 100)), (n_img,))
    cfg0 = dict(model_cfg, dpr_max=0.0)

    def step():
        opt.zero_grad(set_to_none=True)
        logits = O.model_forward(x, params, cfg0, True, {}, None)
        loss = torch.nn.functional.cross_entropy(logits, y, label_smoothing=0.1)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(leaves.values()), 1.0)
        opt.step()
        return float(loss.detach())

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return n_img / dt, dt


def run_reference(args):
    """Reference arm: the reference's own (CPU, PyTorch-eager) algorithm, via the oracle port --
    /root/reference is a Python package that cannot travel to the GPU box."""
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from outlook_grid_vision_transformer_b200.config import BASELINE_CONFIGS, CONFIG_DIR, load_yaml

    wl = BASELINE_CONFIGS[args.workload]
    batch = args.batch or wl["batch"]
    mcfg = load_yaml(CONFIG_DIR / wl["yaml"])["model"]
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    # bounded sample: size it so that (steps + warmup) steps take about two minutes
    rate0, dt0 = cpu_train_step_rate(mcfg, wl["img"], 2, 1, 1)
    budget = 120.0 / max(args.steps + args.warmup, 1)
    n_img = int(max(1, min(64, budget * rate0)))
    rate, dt = cpu_train_step_rate(mcfg, wl["img"], n_img, args.steps, args.warmup)
    sample = f"{n_img} images/step of the same workload, fp32, forward+backward+AdamW, {args.steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_desc(args.workload, wl, batch), "sample": sample},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = ROOT / "gpurun_out" / f"clocks_rank{gpu_index}.csv"

    def start(self):
        try:
            self.path.parent.mkdir(exist_ok=True)
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        for ln in self.path.read_text().splitlines():
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            top = sorted(sm)[len(sm) // 2:]  # upper half = samples under load
            out = {"sm_mhz": statistics.median(top), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def no_decay_groups(model, weight_decay):
    """Two AdamW groups keyed on names like the reference (src/training/warmup.py:4-26)."""
    decay, no_decay = [], []
    for n, p in model.named_parameters():
        if not p.requires_grad:
            continue
        ln = n.lower()
        if n.endswith(".bias") or any(t in ln for t in ("norm", "bn", "ln", "pos")):
            no_decay.append(p)
        else:
            decay.append(p)
    return [{"params": decay, "weight_decay": weight_decay}, {"params": no_decay, "weight_decay": 0.0}]


def run_ours(args):
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F

    import outlook_grid_vision_transformer_b200 as og
    from outlook_grid_vision_transformer_b200 import ops
    from outlook_grid_vision_transformer_b200.config import BASELINE_CONFIGS, CONFIG_DIR
    from outlook_grid_vision_transformer_b200.ddp import BucketedGradAllReduce, broadcast_parameters

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (sm_100a); there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    wl = BASELINE_CONFIGS[args.workload]
    batch = args.batch or wl["batch"]
    img = wl["img"]
    ycfg = og.load_yaml(CONFIG_DIR / wl["yaml"])
    mcfg = ycfg["model"]
    torch.manual_seed(7)
    torch.backends.cudnn.benchmark = True
    model = og.build_model(mcfg).to(dev).to(memory_format=torch.channels_last)
    training = wl["mode"] == "train"
    model.train(training)
    broadcast_parameters(model)
    use_bf16 = wl["dtype"] == "bf16"
    tcfg = ycfg.get("training", {})
    opt = torch.optim.AdamW(no_decay_groups(model, float(tcfg.get("weight_decay", 0.05))), lr=float(tcfg.get("lr", 5e-4)),
                            betas=(0.9, 0.999), eps=1e-8, fused=True) if training else None
    sync = BucketedGradAllReduce(model.parameters()) if (training and world > 1) else None
    ls = float(tcfg.get("label_smoothing", 0.1))

    g = torch.Generator().manual_seed(7 + rank)
    x_host = torch.randn(batch, 3, img, img, generator=g).contiguous(memory_format=torch.channels_last).pin_memory()
    y_host = torch.randint(0, int(mcfg.get("num_classes", 100)), (batch,), generator=g).pin_memory()
    x_dev = x_host.to(dev)
    y_dev = y_host.to(dev)

    from outlook_grid_vision_transformer_b200.engine import TrainStep

    warm = max(args.warmup, 3)
    if training:
        for grp in opt.param_groups:
            grp["capturable"] = True
        clip = tcfg.get("grad_clip_norm")
        runner = TrainStep(model, opt, lambda lg, yy: F.cross_entropy(lg, yy, label_smoothing=ls), x_dev, y_dev,
                           autocast_bf16=use_bf16, grad_sync=sync, use_graph=not args.no_graph, warmup=warm,
                           grad_clip_norm=float(clip) if clip else None)
        step_resident = lambda: runner()
        step_host = lambda: float(runner(x_host, y_host))  # H2D of the batch + D2H of the loss
        eager_body = runner._body
    else:
        def infer(x):
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=use_bf16):
                return model(x).float().logsumexp(1).mean()
        step_resident = lambda: infer(x_dev)
        step_host = lambda: float(infer(x_host.to(dev, non_blocking=True)))
        eager_body = lambda: infer(x_dev)

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    for _ in range(warm):
        step_resident()
    torch.cuda.synchronize()
    l0 = ops.LAUNCHES
    eager_body()  # one eager step: counts the library launches a step consists of
    launches_per_step = ops.LAUNCHES - l0
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(step_resident, args.steps)
    launches = launches_per_step * args.steps
    step_host()
    ms_e2e = timed(step_host, args.steps)
    clocks = sampler.stop()

    value = world * batch / (ms / 1e3)
    e2e_value = world * batch / (ms_e2e / 1e3)

    roof = None
    kernels = {}
    if not args.no_profile and rank != 0:
        # the eager step contains the gradient all-reduce: every rank has to run the profiled passes with rank 0
        # (rank 0 profiling alone blocks in its first collective while the others sit in the final barrier)
        for _ in range(3):
            eager_body()
        torch.cuda.synchronize()
    if not args.no_profile and rank == 0:
        peaks = load_peaks()
        # Eager pass with CUDA events around every library launch.  The stream is first parked on a
        # long sleep so the host can enqueue the whole pass: the events then see back-to-back device
        # execution (no host launch gaps inside the measured intervals).
        nprof = 2
        eager_body()
        torch.cuda.synchronize()
        torch.cuda._sleep(int(0.30 * 1.9e9))
        ops.PROFILER.start()
        for _ in range(nprof):
            eager_body()
        kernels = ops.PROFILER.stop()
        if kernels:
            tot = sum(k["ms"] for k in kernels.values())
            # dominance is judged per CUDA kernel (what an ncu launch list shows), not per call label: all the
            # shapes one GEMM instantiation serves are one kernel
            groups = {}
            for lbl, k in kernels.items():
                gk = groups.setdefault(k.get("cuda_kernel", lbl), dict(ms=0.0, calls=0, bytes=0, flops=0))
                for f in ("ms", "calls", "bytes", "flops"):
                    gk[f] += k[f]
            name, k = max(groups.items(), key=lambda kv: kv[1]["ms"])
            per_ms = k["ms"] / k["calls"]
            gbs = k["bytes"] / k["calls"] / (per_ms * 1e-3) / 1e9
            tfs = k["flops"] / k["calls"] / (per_ms * 1e-3) / 1e12
            if tfs / peaks["tc"] > gbs / peaks["hbm"]:
                roof = {"bound": "tensor", "achieved": tfs, "peak": peaks["tc"], "unit": "TFLOP/s", "frac": tfs / peaks["tc"]}
            else:
                roof = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"]}
            # DRAM traffic per launch of this kernel from the committed ncu --set full capture (profiles/), if any
            traffic = None
            tpath = ROOT / "profiles" / "traffic.json"
            if tpath.exists():
                try:
                    traffic = json.loads(tpath.read_text()).get(name, {}).get("dram_bytes_per_launch")
                except Exception:
                    traffic = None
            # context: a write-only stream on this part tops out well below the copy figure used as `peak`
            nfill = 1 << 28
            buf = torch.empty(nfill, device=dev, dtype=torch.bfloat16)
            for _ in range(3):
                buf.fill_(1.0)
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(10):
                buf.fill_(1.0)
            f1.record()
            torch.cuda.synchronize()
            fill_gbs = nfill * 2 * 10 / (f0.elapsed_time(f1) * 1e-3) / 1e9
            del buf
            roof.update({"traffic": traffic, "kernel": name, "avg_us": per_ms * 1e3, "share_of_ogv_time": k["ms"] / tot,
                         "algorithmic_bytes_per_launch": k["bytes"] / k["calls"],
                         "peak_source": peaks["src"], "ogv_kernel_ms_per_step": tot / nprof,
                         "launches_per_step": k["calls"] // nprof, "hbm_write_only_gbs_this_box": fill_gbs,
                         "by_cuda_kernel": {n: {"ms_per_step": round(v["ms"] / nprof, 3),
                                                "GBps": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1)}
                                            for n, v in sorted(groups.items(), key=lambda kv: -kv[1]["ms"])[:6]}})
            try:
                outp = ROOT / "gpurun_out"
                outp.mkdir(exist_ok=True)
                table = sorted(({"kernel": n, **v, "ms_per_step": v["ms"] / nprof,
                                 "GBps": v["bytes"] / max(v["ms"], 1e-9) / 1e6, "TFLOPs": v["flops"] / max(v["ms"], 1e-9) / 1e9}
                                for n, v in kernels.items()), key=lambda r: -r["ms"])
                (outp / "bench_kernels.json").write_text(json.dumps(table, indent=1))
                for r in table[:25]:
                    print(f"  {r['kernel']:52s} {r['ms_per_step']:8.3f} ms/step  {r['calls'] // nprof:4d} calls  "
                          f"{r['GBps']:8.1f} GB/s  {r['TFLOPs']:7.1f} TF/s", file=sys.stderr)
            except Exception as exc:  # pragma: no cover
                print(f"(could not write kernel table: {exc})", file=sys.stderr)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        n_img = 8
        rate, dt = cpu_train_step_rate(mcfg, img, n_img, 2, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{n_img} images/step of the same workload (oracle port, fp32, fwd+bwd+AdamW), 2 timed steps"}

    if rank == 0:
        xb = x_host.numel() * x_host.element_size() + y_host.numel() * y_host.element_size()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if use_bf16 else "f32", "data": "synthetic",
            "config": {"workload": workload_desc(args.workload, wl, batch), "global_batch": world * batch,
                       "parallelism": f"dp{world}", "optimizer": "grad-norm clip (device-side) + AdamW (torch fused) inside the timed region",
                       "executor": "eager" if args.no_graph else "CUDA graph replay of the whole step",
                       "l2": "per-step working set (saved activations, several GB) >> 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": xb, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear-down of a process group whose collectives were captured into a CUDA graph has been seen to
        # block in destroy_process_group(); every rank has its result out, so synchronise and leave hard.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
