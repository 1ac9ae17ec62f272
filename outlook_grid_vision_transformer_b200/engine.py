"""Train-step executor: the reference's per-batch body (one_epoch_train.py:85-153 -- zero_grad,
autocast forward, loss, backward, optimizer step) captured ONCE into a CUDA graph and replayed, so
the ~3000 kernel launches of a step cost no host time (the reference loop is launch- and
host-sync-bound; SURVEY section 7 "hard parts").  Inputs live in static device buffers; the
compute-dtype weight copies are re-derived from the fp32 master weights inside the graph, so
optimizer updates are seen by the next replay.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import modules as _modules


class TrainStep:
    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, loss_fn: Callable,
                 example_x: torch.Tensor, example_y: torch.Tensor, *, autocast_bf16: bool = True,
                 grad_sync=None, use_graph: bool = True, warmup: int = 3, grad_clip_norm: Optional[float] = None):
        self.model, self.opt, self.loss_fn = model, optimizer, loss_fn
        self.autocast_bf16 = autocast_bf16
        self.grad_sync = grad_sync
        # global gradient-norm clipping like the reference loop (one_epoch_train.py:121-122,141-142), but with the
        # norm kept on the device: no float(gnorm) host sync, so the step stays graph-capturable
        self.grad_clip_norm = grad_clip_norm
        self._params = [p for p in model.parameters() if p.requires_grad]
        self.x = example_x.clone()
        self.y = example_y.clone()
        self.loss = None
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        if use_graph:
            self._capture(warmup)

    # the eager body; also what gets captured
    def _body(self):
        self.opt.zero_grad(set_to_none=True)
        # every compute-dtype weight copy of the model in one launch (from the second step on; the first step
        # casts layer by layer and records what to refresh)
        _modules._BULK_FRESH = _modules.refresh_prepared(self.model)
        try:
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.autocast_bf16):
                logits = self.model(self.x)
        finally:
            _modules._BULK_FRESH = False
        loss = self.loss_fn(logits.float(), self.y)
        loss.backward()
        if self.grad_sync is not None:
            self.grad_sync.finish()
        if self.grad_clip_norm is not None:
            torch.nn.utils.clip_grad_norm_(self._params, self.grad_clip_norm, foreach=True)
        self.opt.step()
        return loss.detach()

    def _capture(self, warmup: int) -> None:
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
            self.opt.zero_grad(set_to_none=True)
            with torch.cuda.graph(self.graph):
                self.loss = self._body()
        finally:
            _modules.FORCE_PREP = False

    def __call__(self, x: Optional[torch.Tensor] = None, y: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Run one step; x / y (if given) are copied into the static input buffers first."""
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        if y is not None:
            self.y.copy_(y, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
            return self.loss
        self.loss = self._body()
        return self.loss
