"""Train-step executor: the reference's per-batch body (one_epoch_train.py:85-166 -- zero_grad, autocast forward,
loss, non-finite guard, backward, gradient-norm clip, AdamW step, LR schedule, loss / top-k bookkeeping) with every
host synchronisation removed, captured ONCE into a CUDA graph and replayed.

State lives in flat fp32 arenas (`FlatState`): parameters, gradients, Adam moments.  Every `p.data` / `p.grad` is a
view, so `state_dict()`, checkpoints and the reference's name-keyed weight-decay grouping (warmup.py:4-26) are
unchanged, while a step needs ONE memset (gradients + BatchNorm scratch + norm accumulator), ONE Bernoulli draw for
all DropPath masks, ONE weight-cast launch, ONE norm reduction and ONE clip+AdamW launch (`ogv_adamw_flat`), and the
gradient all-reduce (ddp.ArenaGradAllReduce) runs in place on arena slices.  Learning rate, bias corrections and
1/world are read by the optimizer kernel from a small DEVICE tensor the host rewrites before each replay, so
`WarmupCosineLR` (warmup.py:29-59) keeps working under graph replay.
"""
from __future__ import annotations

import math
import os
from typing import Callable, Dict, List, Optional

import torch

from . import functional as _OF
from . import modules as _modules
from . import ops

GRANULE = 8  # floats: every parameter starts on a 32-byte boundary of the arenas (TMA / v4 `red` alignment)


def default_no_decay(name: str) -> bool:
    """The reference's grouping (src/training/warmup.py:4-26): biases, norms, positional / class tokens."""
    ln = name.lower()
    return name.endswith(".bias") or any(t in ln for t in ("norm", "bn", "ln", "pos", "cls_token"))


class WarmupCosineLR:
    """warmup.py:29-59 as a pure function of the step number (host side; the value is fed to the device scalar)."""

    def __init__(self, base_lr: float, total_steps: int, warmup_steps: int, min_lr: float = 0.0):
        self.base_lr, self.total_steps, self.warmup_steps, self.min_lr = float(base_lr), int(total_steps), int(warmup_steps), float(min_lr)

    def lr_after(self, t: int) -> float:
        """learning rate in force after `t` calls of scheduler.step() (t = 0: the optimizer's base lr)."""
        if t <= 0:
            return self.base_lr
        if t <= self.warmup_steps and self.warmup_steps > 0:
            return self.base_lr * (t / self.warmup_steps)
        tt = min(t, self.total_steps)
        denom = max(1, self.total_steps - self.warmup_steps)
        progress = (tt - self.warmup_steps) / denom
        return self.min_lr + (self.base_lr - self.min_lr) * 0.5 * (1.0 + math.cos(math.pi * progress))


class _Scratch:
    def __init__(self, buf: torch.Tensor):
        self.buf, self.pos = buf, 0

    def reset(self) -> None:
        self.pos = 0

    def take(self, n: int) -> Optional[torch.Tensor]:
        n8 = (n + GRANULE - 1) // GRANULE * GRANULE
        if self.pos + n8 > self.buf.numel():
            return None
        t = self.buf[self.pos:self.pos + n]
        self.pos += n8
        return t


class FlatState:
    """Parameters, gradients and Adam moments of `model` in flat fp32 arenas (model must already be on its device).

    Arena order = registration (forward) order, so gradient readiness in backward runs from the end of the arena
    to its front -- except that each OutlookAttention2d's `v` and `attn` weights (and biases) are made adjacent and
    zero-padded to the fused v|logits GEMM's row count, so that GEMM's weight gradient lands in place."""

    def __init__(self, model: torch.nn.Module, no_decay: Callable[[str], bool] = default_no_decay,
                 scratch_floats: int = 1 << 18):
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        if not named:
            raise ValueError("model has no trainable parameters")
        dev = named[0][1].device
        if any(p.device != dev or p.dtype != torch.float32 for _, p in named):
            raise ValueError("FlatState needs every trainable parameter in float32 on one device")
        self.model = model
        import numpy as np

        by_param = {p: n for n, p in named}
        fused = {}  # first parameter of a fused pair -> (module, kind)
        for m in model.modules():
            if isinstance(m, _modules.OutlookAttention2d) and m.v.bias is not None and m.attn.bias is not None and \
                    all(q in by_param for q in (m.v.weight, m.attn.weight, m.v.bias, m.attn.bias)):
                if m.v.weight.shape[0] % GRANULE:
                    raise ValueError("outlook attention needs dim % 8 == 0")
                for q in (m.v.weight, m.attn.weight, m.v.bias, m.attn.bias):
                    fused[q] = m
        # items: runs of parameters laid out back to back, then padding up to `reserve` floats
        items, placed = [], set()
        for n, p in named:
            if p in placed:
                continue
            m = fused.get(p)
            if m is None:
                items.append(([p], (p.numel() + GRANULE - 1) // GRANULE * GRANULE))
                placed.add(p)
                continue
            C, nl = m.v.weight.shape[0], m.attn.weight.shape[0]
            npad = (C + nl + 7) // 8 * 8
            items.append(([m.v.weight, m.attn.weight], npad * C))   # rows of the fused v|logits GEMM: [v ; attn ; 0]
            items.append(([m.v.bias, m.attn.bias], npad))
            placed.update((m.v.weight, m.attn.weight, m.v.bias, m.attn.bias))
        offs, pos = {}, 0
        for plist, reserve in items:
            o = pos
            for p in plist:
                offs[p] = o
                o += p.numel()
            pos += reserve
        self.n = pos
        self.params = [p for plist, _ in items for p in plist]
        self.names = [by_param[p] for p in self.params]
        self.offsets = [offs[p] for p in self.params]
        self.P = torch.zeros(self.n, device=dev, dtype=torch.float32)
        self.M = torch.zeros_like(self.P)
        self.V = torch.zeros_like(self.P)
        # gradient arena + pre-zeroed scratch + [gnorm^2, pad...] in ONE allocation = one memset per step
        self._g_all = torch.zeros(self.n + scratch_floats + GRANULE, device=dev, dtype=torch.float32)
        self.G = self._g_all[:self.n]
        self.scratch = _Scratch(self._g_all[self.n:self.n + scratch_floats])
        self.gnorm_sq = self._g_all[self.n + scratch_floats:self.n + scratch_floats + 1]
        self.skipped = torch.zeros(1, device=dev, dtype=torch.float32)
        decay = np.zeros(((self.n // GRANULE + 31) // 32) * 32, dtype=np.uint8)
        for n, p, off in zip(self.names, self.params, self.offsets):
            with torch.no_grad():
                dense = _dense(p)
                view = self.P[off:off + p.numel()]
                view = view.as_strided(p.shape, p.stride()) if dense else view.view(p.shape)
                view.copy_(p.detach())
                p.data = view
                gview = self.G[off:off + p.numel()]
                p.grad = gview.as_strided(p.shape, p.stride()) if dense else gview.view(p.shape)
            if not no_decay(n):
                decay[off // GRANULE:(off + p.numel() + GRANULE - 1) // GRANULE] = 1
        words = np.packbits(decay.reshape(-1, 32), axis=1, bitorder="little").view("<u4").reshape(-1)
        self.decay_bits = torch.from_numpy(words.view(np.int32).copy()).to(dev)
        self.n_decay = sum(p.numel() for n, p in zip(self.names, self.params) if not no_decay(n))
        # module marks: branches accumulate parameter gradients in place; MBConv BatchNorm counters share one arena
        nbt = []
        from . import model as _model  # the callers around the blocks: conv -> BN -> act units and the head

        for m in model.modules():
            if isinstance(m, (_modules.MLP2d, _modules.MLP, _modules.OutlookAttention2d, _modules.MBConv,
                              _modules.MultiHeadSelfAttention, _model.ConvStem, _model.Downsample, _model._Backbone)):
                m.__dict__["_ogv_direct"] = True
            if isinstance(m, _modules.OutlookAttention2d) and fused.get(m.v.weight) is m:
                C, nl = m.v.weight.shape[0], m.attn.weight.shape[0]
                npad = (C + nl + 7) // 8 * 8
                ow, ob = offs[m.v.weight], offs[m.v.bias]
                m.__dict__["_ogv_flat"] = (self.G[ow:ow + npad * C].view(npad, C), self.G[ob:ob + npad])
            if isinstance(m, _modules.MBConv) and not isinstance(m.expand, torch.nn.Identity):
                bns = [b for b in (m.expand[1], m.depthwise[1], m.project[1]) if isinstance(b, torch.nn.BatchNorm2d)
                       and b.num_batches_tracked is not None]
                if len(bns) == 3:
                    nbt += bns
                    m.__dict__["_ogv_nbt_shared"] = True
        self.nbt = None
        if nbt:
            self.nbt = torch.stack([b.num_batches_tracked.detach().to(dev) for b in nbt])
            for i, b in enumerate(nbt):
                b._buffers["num_batches_tracked"] = self.nbt[i]

    # --------------------------------------------------------------------------------------------
    def zero_grad(self) -> None:
        """ONE memset: parameter gradients, the BatchNorm-statistics scratch and the norm accumulator."""
        self._g_all.zero_()
        self.scratch.reset()

    def grad_norm_sq(self) -> torch.Tensor:
        ops.sumsq(self.G, self.gnorm_sq)
        return self.gnorm_sq

    def release(self) -> None:
        """Detach the model from the arenas' special modes (parameters stay views of P, which is harmless)."""
        for m in self.model.modules():
            for k in ("_ogv_direct", "_ogv_flat", "_ogv_nbt_shared"):
                m.__dict__.pop(k, None)
        for p in self.params:
            p.grad = None


def _dense(p: torch.Tensor) -> bool:
    """non-overlapping and dense in SOME dimension order (e.g. channels_last conv weights): strides can be kept."""
    dims = sorted(range(p.dim()), key=lambda d: (p.stride(d), p.size(d)))
    expect = 1
    for d in dims:
        if p.size(d) != 1 and p.stride(d) != expect:
            return False
        expect *= p.size(d)
    return True


class TrainStep:
    """One training step, eager or as a replayed CUDA graph.

        step = TrainStep(model, loss_fn, example_x, example_y, lr=5e-4, weight_decay=0.05, grad_clip_norm=1.0,
                         scheduler=WarmupCosineLR(5e-4, total, warmup), grad_sync=ArenaGradAllReduce(...))
        loss = step(x, y)            # device tensor, no host sync
        step.metrics()               # {'loss', 'top1', 'top3', 'top5', 'samples', 'skipped_steps'} -- ONE sync, e.g. per epoch

    Construction runs `warmup` eager steps on the example batch (cuDNN autotune, weight-cast table) and captures the
    graph; parameters, Adam moments, BatchNorm buffers, the step counter and the CUDA RNG state are snapshotted
    before and restored after, so the first user step starts from the state the model was handed over in."""

    def __init__(self, model: torch.nn.Module, loss_fn: Callable, example_x: torch.Tensor, example_y: torch.Tensor, *,
                 lr: float = 5e-4, weight_decay: float = 0.05, betas=(0.9, 0.999), eps: float = 1e-8,
                 autocast_bf16: bool = True, grad_sync=None, use_graph: bool = True, warmup: int = 3,
                 grad_clip_norm: Optional[float] = None, scheduler: Optional[WarmupCosineLR] = None,
                 no_decay: Callable[[str], bool] = default_no_decay, fused_droppath: bool = True, world: int = 1,
                 flat: Optional[FlatState] = None, side_wgrad: Optional[bool] = None):
        self.model, self.loss_fn = model, loss_fn
        self.flat = flat if flat is not None else FlatState(model, no_decay)
        self.autocast_bf16 = autocast_bf16
        self.grad_sync = grad_sync
        self.grad_clip_norm = grad_clip_norm
        self.betas, self.eps, self.weight_decay = betas, float(eps), float(weight_decay)
        self.base_lr, self.scheduler, self.world = float(lr), scheduler, int(world)
        self._lr_override: Optional[float] = None
        self.step_num = 0
        dev = self.flat.P.device
        if side_wgrad is None:
            side_wgrad = os.environ.get("OGV_SIDE_WGRAD", "1") != "0"
        # weight-gradient GEMMs on a second stream (joined before the gradient exchange / the optimizer)
        self._side = ops.SideStream(dev) if (side_wgrad and dev.type == "cuda") else None
        self.x = example_x.clone()
        self.y = example_y.clone()
        self.hyper = torch.zeros(5, device=dev, dtype=torch.float32)
        # pinned staging slots for the per-step hyper-parameters: the host may run several replays ahead of the device,
        # so a slot is rewritten only after the asynchronous copy that read it has executed (event per slot)
        self._hyper_ring = [torch.zeros(5, dtype=torch.float32).pin_memory() if dev.type == "cuda" else torch.zeros(5)
                            for _ in range(4)]
        self._hyper_evt = [None] * 4
        self._hyper_i = 0
        self.acc = torch.zeros(5, device=dev, dtype=torch.float32)      # loss*B, top1, top3, top5, B
        self.loss: Optional[torch.Tensor] = None
        self._stage = None          # (x, y staging buffers, ready event, free event) of prefetch()
        self._copy_stream = None
        self._loss_host = None      # two pinned scalars + events of loss_to_host_async()
        self.logits: Optional[torch.Tensor] = None                       # static fp32 [B, K] buffer of the last step
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        # one Bernoulli draw for every DropPath of the model (call order), instead of 4 launches per DropPath
        self._dp_keep = None
        if fused_droppath:
            keeps = [1.0 - m.drop_prob for m in model.modules() if isinstance(m, _modules.DropPath) and m.drop_prob > 0.0]
            if keeps:
                B = example_x.shape[0]
                self._dp_keep = torch.tensor(keeps, device=dev, dtype=torch.float32).view(-1, 1).expand(-1, B).contiguous()
        if use_graph:
            self._capture(warmup)

    # ------------------------------------------------------------------------------------------ schedule
    def set_lr(self, lr: Optional[float]) -> None:
        """Override the learning rate of the following steps (None: back to base lr / scheduler)."""
        self._lr_override = None if lr is None else float(lr)

    def current_lr(self) -> float:
        if self._lr_override is not None:
            return self._lr_override
        if self.scheduler is not None:
            return self.scheduler.lr_after(self.step_num)  # scheduler.step() follows optimizer.step() in the reference
        return self.base_lr

    def _write_hyper(self) -> None:
        t = self.step_num + 1
        i = self._hyper_i
        self._hyper_i = (i + 1) % len(self._hyper_ring)
        if self._hyper_evt[i] is not None:
            self._hyper_evt[i].synchronize()  # the copy issued four steps ago: done long before, except far ahead of the device
        h = self._hyper_ring[i]
        h[0] = self.current_lr()
        h[1] = 1.0 - self.betas[0] ** t
        h[2] = 1.0 - self.betas[1] ** t
        h[3] = 1.0 / self.world
        h[4] = float(self.grad_clip_norm) if self.grad_clip_norm else 0.0
        self.hyper.copy_(h, non_blocking=True)
        if self.hyper.is_cuda:
            if self._hyper_evt[i] is None:
                self._hyper_evt[i] = torch.cuda.Event()
            self._hyper_evt[i].record(torch.cuda.current_stream(self.hyper.device))

    # ------------------------------------------------------------------------------------------ the step
    def _body(self):
        fl = self.flat
        fl.zero_grad()
        if fl.nbt is not None:
            fl.nbt.add_(1)
        table = None
        if self._dp_keep is not None and self.model.training:
            table = torch.bernoulli(self._dp_keep).div_(self._dp_keep)
        _modules.DROP_TABLE = iter(table.unbind(0)) if table is not None else None
        _OF.SCRATCH = fl.scratch
        ops.PROFILER.tag = None
        # every compute-dtype weight copy of the model in one launch (from the second step on; the first step
        # casts layer by layer and records what to refresh)
        _modules._BULK_FRESH = _modules.refresh_prepared(self.model)
        try:
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.autocast_bf16):
                logits = self.model(self.x)
            logits = logits.float()
            loss = self.loss_fn(logits, self.y)
            ops.SIDE = self._side
            loss.backward()
        finally:
            ops.SIDE = None
            if self._side is not None:
                self._side.join()
            _modules._BULK_FRESH = False
            _modules.DROP_TABLE = None
            _OF.SCRATCH = None
            ops.PROFILER.tag = None
        if self.grad_sync is not None:
            self.grad_sync.finish()
        loss = loss.detach().float().reshape(1)
        gn = fl.grad_norm_sq()
        ops.adamw_flat(fl.P, fl.G, fl.M, fl.V, fl.decay_bits, self.hyper, gn, loss, self.betas[0], self.betas[1],
                       self.eps, self.weight_decay, fl.skipped)
        if self.y.dtype == torch.int64 and self.y.dim() == 1:
            ops.train_metrics(logits.detach().contiguous(), self.y, loss, self.acc)
        return loss, logits.detach()

    def _snapshot(self):
        fl = self.flat
        bufs = {k: b.detach().clone() for k, b in self.model.named_buffers()}
        return dict(P=fl.P.clone(), M=fl.M.clone(), V=fl.V.clone(), bufs=bufs, nbt=None if fl.nbt is None else fl.nbt.clone(),
                    rng=torch.cuda.get_rng_state(fl.P.device), step=self.step_num, acc=self.acc.clone(),
                    skipped=fl.skipped.clone())

    def _restore(self, s) -> None:
        fl = self.flat
        with torch.no_grad():
            fl.P.copy_(s["P"]); fl.M.copy_(s["M"]); fl.V.copy_(s["V"])
            for k, b in self.model.named_buffers():
                b.copy_(s["bufs"][k])
            if fl.nbt is not None:
                fl.nbt.copy_(s["nbt"])
            self.acc.copy_(s["acc"]); fl.skipped.copy_(s["skipped"])
        torch.cuda.set_rng_state(s["rng"], fl.P.device)
        self.step_num = s["step"]
        _modules.note_parameters_updated()

    def _capture(self, warmup: int) -> None:
        snap = self._snapshot()
        self._write_hyper()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 2)):  # >= 2: the bulk weight-refresh table is built (H2D copy) in step 2
                self._body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        _modules.FORCE_PREP = True  # weight casts must be nodes of the graph, not cache hits
        try:
            with torch.cuda.graph(self.graph):
                self.loss, self.logits = self._body()
        finally:
            _modules.FORCE_PREP = False
        self._restore(snap)

    def __call__(self, x: Optional[torch.Tensor] = None, y: Optional[torch.Tensor] = None, *,
                 staged: bool = False) -> torch.Tensor:
        """Run one step; x / y (if given) are copied into the static input buffers first; `staged=True` takes the batch
        that `prefetch()` put into the staging buffers.  Returns the loss (device)."""
        if staged:
            if self._stage is None:
                raise RuntimeError("TrainStep(staged=True) needs a prefetch() first")
            sx, sy, ready, free = self._stage
            main = torch.cuda.current_stream(self.x.device)
            main.wait_event(ready)
            self.x.copy_(sx, non_blocking=True)   # device-to-device, a few microseconds
            self.y.copy_(sy, non_blocking=True)
            free.record(main)
        else:
            if x is not None:
                self.x.copy_(x, non_blocking=True)
            if y is not None:
                self.y.copy_(y, non_blocking=True)
        self._write_hyper()
        if self.graph is not None:
            self.graph.replay()
        else:
            self.loss, self.logits = self._body()
        self.step_num += 1
        _modules.note_parameters_updated()
        return self.loss

    # ------------------------------------------------------------------------------------------ input / loss pipelining
    def prefetch(self, x_host: torch.Tensor, y_host: torch.Tensor) -> None:
        """Start the host-to-device copy of the NEXT batch (pinned host tensors) on a copy stream while the current step
        runs -- what the reference's DataLoader(pin_memory=True) + `.to(device, non_blocking=True)` loop
        (one_epoch_train.py:85-93) does.  The next `step(staged=True)` consumes it."""
        dev = self.x.device
        if self._stage is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            free = torch.cuda.Event()
            free.record(torch.cuda.current_stream(dev))
            self._stage = (torch.empty_like(self.x), torch.empty_like(self.y), torch.cuda.Event(), free)
        sx, sy, ready, free = self._stage
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(free)  # the previous staged batch has been copied into the static buffers
            sx.copy_(x_host, non_blocking=True)
            sy.copy_(y_host, non_blocking=True)
            ready.record(self._copy_stream)

    def loss_to_host_async(self) -> None:
        """Queue a device-to-host copy of this step's loss into pinned memory (no host synchronisation)."""
        if self._loss_host is None:
            self._loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
            self._loss_evt = [torch.cuda.Event(), torch.cuda.Event()]
            self._loss_idx, self._loss_count = 0, 0
        i = self._loss_idx
        self._loss_count += 1
        self._loss_host[i].copy_(self.loss.detach().float().reshape(()), non_blocking=True)
        self._loss_evt[i].record(torch.cuda.current_stream(self.x.device))
        self._loss_idx = i ^ 1

    def previous_loss(self) -> Optional[float]:
        """The loss queued by the `loss_to_host_async()` call BEFORE the latest one (waits for that copy only): a one-step
        lag keeps the host a full step ahead of the device instead of stalling it every iteration."""
        if self._loss_host is None or self._loss_count < 2:
            return None
        i = self._loss_idx  # the slot the next call will overwrite = the older of the two
        if not self._loss_evt[i].query():
            self._loss_evt[i].synchronize()
        return float(self._loss_host[i])

    # ------------------------------------------------------------------------------------------ bookkeeping
    def metrics(self, reset: bool = True) -> Dict[str, float]:
        """Loss / top-k accumulated ON THE DEVICE since the last reset (one_epoch_train.py:155-166) -- one host sync."""
        a = self.acc.tolist()
        n = max(a[4], 1.0)
        out = {"loss": a[0] / n, "top1": 100.0 * a[1] / n, "top3": 100.0 * a[2] / n, "top5": 100.0 * a[3] / n,
               "samples": a[4], "skipped_steps": float(self.flat.skipped.item())}
        if reset:
            self.acc.zero_()
        return out

    def optimizer_state(self) -> Dict[str, object]:
        return {"step": self.step_num, "exp_avg": self.flat.M, "exp_avg_sq": self.flat.V, "offsets":
                dict(zip(self.flat.names, self.flat.offsets))}
