"""Batch-sharded data parallelism for the train step: one process per GPU, per-replica BatchNorm
statistics (the reference has no SyncBN), and ONE exchange per iteration -- a sum all-reduce of
the parameter gradients in flat ~25 MB buckets, issued from post-accumulate-grad hooks in reverse
parameter order on a side stream so the collective overlaps the rest of backward.

The reference itself is single-process (one_epoch_train.py:31); `torch.distributed` (NCCL over
NVLink 5 / NVSwitch on the box, gloo in the CPU tests) is the plumbing.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


class _Bucket:
    __slots__ = ("params", "numel", "flat", "pending", "work", "offsets")

    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = params
        self.offsets = []
        n = 0
        for p in params:
            self.offsets.append(n)
            n += p.numel()
        self.numel = n
        self.flat: Optional[torch.Tensor] = None
        self.pending = len(params)
        self.work = None


class BucketedGradAllReduce:
    """Usage:  sync = BucketedGradAllReduce(model.parameters());  loss.backward();  sync.finish()."""

    def __init__(self, params, bucket_bytes: int = 25 * 1024 * 1024, process_group=None, average: bool = True):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.average = average
        plist = [p for p in params if p.requires_grad]
        self.buckets: List[_Bucket] = []
        cur, cur_bytes = [], 0
        for p in reversed(plist):  # gradients become ready roughly in reverse registration order
            cur.append(p)
            cur_bytes += p.numel() * 4
            if cur_bytes >= bucket_bytes:
                self.buckets.append(_Bucket(cur))
                cur, cur_bytes = [], 0
        if cur:
            self.buckets.append(_Bucket(cur))
        self._owner = {}
        for b in self.buckets:
            for p in b.params:
                self._owner[p] = b
        self._comm_stream = None
        self._hooks = []
        if self.world > 1:
            for p in plist:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad_ready))

    # ------------------------------------------------------------------------------------------
    def _on_grad_ready(self, p: torch.nn.Parameter) -> None:
        b = self._owner[p]
        b.pending -= 1
        if b.pending == 0:
            self._launch(b)

    def _launch(self, b: _Bucket) -> None:
        dev = b.params[0].device
        if b.flat is None or b.flat.device != dev:
            b.flat = torch.empty(b.numel, device=dev, dtype=torch.float32)
        views = [b.flat[o:o + p.numel()].view_as(p) for o, p in zip(b.offsets, b.params)]
        if dev.type == "cuda":
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device=dev)
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(dev))
            with torch.cuda.stream(self._comm_stream):
                self._comm_stream.wait_event(ready)
                torch._foreach_copy_(views, [p.grad for p in b.params])
                b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            torch._foreach_copy_(views, [p.grad for p in b.params])
            b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self) -> None:
        """Wait for every bucket, write the (averaged) gradients back, re-arm the hooks' counters."""
        if self.world > 1:
            for b in self.buckets:
                if b.pending != 0:  # a parameter received no gradient this step: reduce what exists
                    for p in b.params:
                        if p.grad is None:
                            p.grad = torch.zeros_like(p)
                    b.pending = 0
                    self._launch(b)
            scale = 1.0 / self.world if self.average else 1.0
            for b in self.buckets:
                dev = b.params[0].device
                views = [b.flat[o:o + p.numel()].view_as(p) for o, p in zip(b.offsets, b.params)]
                if dev.type == "cuda":
                    with torch.cuda.stream(self._comm_stream):
                        b.work.wait()
                        if scale != 1.0:
                            b.flat.mul_(scale)
                        torch._foreach_copy_([p.grad for p in b.params], views)
                else:
                    b.work.wait()
                    if scale != 1.0:
                        b.flat.mul_(scale)
                    torch._foreach_copy_([p.grad for p in b.params], views)
            if self._comm_stream is not None:
                torch.cuda.current_stream().wait_stream(self._comm_stream)
        for b in self.buckets:
            b.pending = len(b.params)
            b.work = None

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []


class ArenaGradAllReduce:
    """Gradient all-reduce IN PLACE on slices of engine.FlatState's flat gradient arena: no pack / unpack copies and
    no scaling pass (the optimizer kernel folds 1/world into its gradient scale, `TrainStep(world=...)`).

    Buckets are contiguous arena slices cut from the END (gradients become ready from the last layer backwards);
    the bucket that finishes last -- the front of the arena, i.e. the first layers -- is kept small (`tail_bytes`), so
    the all-reduce that cannot overlap with backward is short.  A bucket is launched on a side stream the moment its
    last parameter reports in: autograd post-accumulate hooks for the layers torch differentiates (stem, downsample,
    head), `functional.GRAD_READY` for the fused branches that accumulate into the arena themselves."""

    def __init__(self, flat, bucket_bytes: int = 8 << 20, tail_bytes: int = 1 << 20, process_group=None):
        from . import functional as _OF

        self.flat, self.group = flat, process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        params, offsets = flat.params, flat.offsets
        self.buckets = []  # dict(lo, hi, params, pending, work), in launch order (end of the arena first)
        cur, cur_hi, tail_cut = [], flat.n, False
        for i in range(len(params) - 1, -1, -1):
            cur.append(params[i])
            lo = offsets[i]
            cut = (cur_hi - lo) * 4 >= bucket_bytes
            if not tail_cut and 0 < lo * 4 <= tail_bytes:  # what is left in front of here is the short last bucket
                cut, tail_cut = True, True
            if cut and i > 0:
                self.buckets.append(dict(lo=lo, hi=cur_hi, params=cur, pending=len(cur), work=None))
                cur, cur_hi = [], lo
        if cur:
            self.buckets.append(dict(lo=0, hi=cur_hi, params=cur, pending=len(cur), work=None))
        self._owner = {p: b for b in self.buckets for p in b["params"]}
        self._seen = set()  # parameters already reported this step (both notification routes may fire for one)
        self._comm_stream = None
        self._hooks = []
        self._OF = _OF
        if self.world > 1:
            for p in params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad_ready))
            _OF.GRAD_READY = self._on_params_ready

    def _on_params_ready(self, plist) -> None:
        for p in plist:
            self._on_grad_ready(p)

    def _on_grad_ready(self, p) -> None:
        # A parameter can be reported twice in one backward: autograd runs the post-accumulate hook of a leaf even when
        # the fused branch returned None for it (it accumulated into the arena itself and reported through GRAD_READY).
        # Count every parameter once, or a bucket is launched while half of its gradients are still to be written.
        b = self._owner.get(p)
        if b is None or id(p) in self._seen:
            return
        self._seen.add(id(p))
        b["pending"] -= 1
        if b["pending"] == 0:
            self._launch(b)

    def _launch(self, b) -> None:
        buf = self.flat.G[b["lo"]:b["hi"]]
        if buf.is_cuda:
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device=buf.device)
            from . import ops as _ops

            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(buf.device))
            with torch.cuda.stream(self._comm_stream):
                self._comm_stream.wait_event(ready)
                if _ops.SIDE is not None:  # weight gradients of this bucket may still be running on the side stream
                    self._comm_stream.wait_stream(_ops.SIDE.stream)
                b["work"] = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            b["work"] = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self) -> None:
        """Launch what is still pending (parameters that received no gradient), wait for every bucket, re-arm."""
        if self.world > 1:
            for b in self.buckets:
                if b["work"] is None:
                    self._launch(b)
            for b in self.buckets:
                if self._comm_stream is not None:
                    with torch.cuda.stream(self._comm_stream):
                        b["work"].wait()
                else:
                    b["work"].wait()
            if self._comm_stream is not None:
                torch.cuda.current_stream().wait_stream(self._comm_stream)
        for b in self.buckets:
            b["pending"] = len(b["params"])
            b["work"] = None
        self._seen.clear()

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []
        if self._OF.GRAD_READY == self._on_params_ready:
            self._OF.GRAD_READY = None


def broadcast_parameters(module: torch.nn.Module, src: int = 0, process_group=None) -> None:
    """Make every replica start from rank `src`'s parameters and buffers."""
    if not dist.is_initialized() or dist.get_world_size(process_group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=process_group)
