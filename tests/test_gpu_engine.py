"""-m gpu: the sync-free train step (engine.TrainStep over engine.FlatState) against the reference loop's semantics
(one_epoch_train.py:85-166, warmup.py:4-59, metrics.py:7-24), restated with torch.optim.AdamW / clip_grad_norm_ /
topk on the same model: flat clip+AdamW == torch AdamW with the two param groups, LR schedule under graph replay,
non-finite guard, device-side metrics, eval after replay sees fresh weights, capture leaves no trace."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

CFG = {"type": "model_a", "num_classes": 10, "stem_dim": 16, "dpr_max": 0.0,
       "stages": [dict(dim=16, depth=1, num_heads=2, grid_size=2, outlook_heads=2),
                  dict(dim=32, depth=1, num_heads=2, grid_size=2, outlook_heads=2)]}


def _data(B=4, seed=3):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 3, 8, 8, generator=g).to(DEV), torch.randint(0, 10, (B,), generator=g).to(DEV)


def _model(seed=11, cfg=CFG):
    import outlook_grid_vision_transformer_b200 as og
    torch.manual_seed(seed)
    return og.build_model(cfg).to(DEV).train()


def _torch_groups(model, wd):
    from outlook_grid_vision_transformer_b200.engine import default_no_decay
    decay = [p for n, p in model.named_parameters() if not default_no_decay(n)]
    nodecay = [p for n, p in model.named_parameters() if default_no_decay(n)]
    return [{"params": decay, "weight_decay": wd}, {"params": nodecay, "weight_decay": 0.0}]


def _ref_sched_lr(base, t, total, warm, min_lr):
    """warmup.py:44-53 for step_num = t"""
    if t <= warm and warm > 0:
        return base * (t / warm)
    tt = min(t, total)
    prog = (tt - warm) / max(1, total - warm)
    return min_lr + (base - min_lr) * 0.5 * (1.0 + math.cos(math.pi * prog))


@pytest.mark.parametrize("use_graph", [False, True], ids=["eager", "graph"])
def test_flat_step_equals_reference_loop_semantics(use_graph):
    """fp32 mode, 6 steps with clip 0.5, AdamW (wd 0.05 on weights only), warm-up 3 + cosine: parameters after every
    step equal the torch.optim.AdamW loop written like the reference's (optimizer.step(); scheduler.step())."""
    from outlook_grid_vision_transformer_b200.engine import TrainStep, WarmupCosineLR
    x, y = _data()
    steps, total, warm, base, min_lr, clip, wd = 6, 10, 3, 2e-3, 1e-5, 0.5, 0.05
    # Adam's update lr*g/(|g|+eps) is chaotic for |g| ~ eps: with the default 1e-8 an element whose gradient is pure
    # summation-order noise (1e-9) moves by a random fraction of lr in EITHER implementation.  1e-6 keeps the test about
    # the optimizer arithmetic (d update / d g <= lr / eps = 2e3, noise 1e-9 -> 2e-6, below the tolerance).
    eps = 1e-6
    # --- reference-loop restatement
    ref = _model()
    opt = torch.optim.AdamW(_torch_groups(ref, wd), lr=base, betas=(0.9, 0.999), eps=eps)
    ref_losses, ref_params = [], []
    for t in range(1, steps + 1):
        opt.zero_grad(set_to_none=True)
        loss = F.cross_entropy(ref(x).float(), y, label_smoothing=0.1)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), clip)
        opt.step()
        for g in opt.param_groups:
            g["lr"] = _ref_sched_lr(base, t, total, warm, min_lr)
        ref_losses.append(float(loss.detach()))
        ref_params.append({k: p.detach().clone() for k, p in ref.named_parameters()})
    ref_bufs = {k: b.detach().clone() for k, b in ref.named_buffers()}
    # --- the engine
    model = _model()
    step = TrainStep(model, lambda lg, yy: F.cross_entropy(lg, yy, label_smoothing=0.1), x, y, lr=base, weight_decay=wd,
                     autocast_bf16=False, use_graph=use_graph, warmup=2, grad_clip_norm=clip, eps=eps,
                     scheduler=WarmupCosineLR(base, total, warm, min_lr))
    for t in range(steps):
        loss = step()
        assert float(loss) == pytest.approx(ref_losses[t], rel=2e-4), f"step {t}"
        for k, p in model.named_parameters():
            a, b = p.detach(), ref_params[t][k]
            if k.endswith("mhsa.qkv.bias"):
                # the key bias has an exactly-zero gradient in exact arithmetic (softmax is shift invariant): what
                # reaches Adam is summation-order noise, which its normalisation turns into +-O(lr) either way
                C = a.numel() // 3
                a, b = torch.cat([a[:C], a[2 * C:]]), torch.cat([b[:C], b[2 * C:]])
            torch.testing.assert_close(a, b, rtol=2e-3, atol=2e-5, msg=lambda m: f"step {t} {k}: {m}")
    for k, b in model.named_buffers():
        torch.testing.assert_close(b.detach().float(), ref_bufs[k].float(), rtol=1e-3, atol=1e-5, msg=lambda m: f"{k}: {m}")
    assert step.metrics()["skipped_steps"] == 0


def test_graph_replay_equals_eager_bf16_and_trains():
    from outlook_grid_vision_transformer_b200.engine import TrainStep
    x, y = _data()
    losses = []
    for use_graph in (False, True):
        model = _model()
        step = TrainStep(model, lambda lg, yy: F.cross_entropy(lg, yy), x, y, lr=1e-3, autocast_bf16=True,
                         use_graph=use_graph, warmup=2)
        losses.append([float(step()) for _ in range(4)])
    assert losses[0] == pytest.approx(losses[1], rel=2e-2), f"{losses}"
    assert losses[1][-1] < losses[1][0]


def test_capture_leaves_no_trace():
    """Warm-up + capture run real steps on the example batch; parameters, Adam moments, BatchNorm buffers and the
    step counter are restored afterwards (the first user step starts from the state handed over)."""
    from outlook_grid_vision_transformer_b200.engine import TrainStep
    x, y = _data()
    model = _model()
    before = {k: v.detach().clone() for k, v in model.state_dict().items()}
    step = TrainStep(model, lambda lg, yy: F.cross_entropy(lg, yy), x, y, lr=1e-2, use_graph=True, warmup=3)
    for k, v in model.state_dict().items():
        assert torch.equal(v, before[k]), k
    assert step.step_num == 0 and float(step.flat.M.abs().sum()) == 0.0 and float(step.flat.V.abs().sum()) == 0.0


def test_non_finite_loss_skips_the_update():
    """one_epoch_train.py:99-109: a non-finite loss must not touch the weights (nor the Adam moments)."""
    from outlook_grid_vision_transformer_b200.engine import TrainStep
    x, y = _data()
    model = _model()
    step = TrainStep(model, lambda lg, yy: F.cross_entropy(lg, yy), x, y, lr=1e-2, use_graph=True, warmup=2,
                     grad_clip_norm=1.0)
    step()
    p0, m0 = step.flat.P.clone(), step.flat.M.clone()
    bad = x.clone()
    bad[0, 0, 0, 0] = float("inf")
    step(bad, y)
    assert torch.equal(step.flat.P, p0) and torch.equal(step.flat.M, m0)
    step(x, y)
    assert not torch.equal(step.flat.P, p0)
    assert step.metrics()["skipped_steps"] == 1.0


def test_device_side_metrics_match_topk():
    from outlook_grid_vision_transformer_b200.engine import TrainStep
    x, y = _data(B=16, seed=5)
    model = _model()
    step = TrainStep(model, lambda lg, yy: F.cross_entropy(lg, yy), x, y, lr=0.0, weight_decay=0.0, use_graph=True, warmup=2)
    want = {"loss": 0.0, 1: 0.0, 3: 0.0, 5: 0.0}
    n = 3
    for _ in range(n):
        loss = step()
        lg = step.logits
        _, pred = torch.topk(lg, 5, dim=1)
        hit = pred.eq(y.view(-1, 1))
        for k in (1, 3, 5):
            want[k] += float(hit[:, :k].any(1).float().sum())
        want["loss"] += float(loss) * 16
    got = step.metrics()
    assert got["samples"] == n * 16
    assert got["loss"] == pytest.approx(want["loss"] / (n * 16), rel=1e-5)
    for k in (1, 3, 5):
        assert got[f"top{k}"] == pytest.approx(100.0 * want[k] / (n * 16))


def test_eval_after_graph_replay_uses_fresh_weights():
    """the eval-mode weight-copy cache must not survive optimizer updates it cannot see."""
    import outlook_grid_vision_transformer_b200 as og
    from outlook_grid_vision_transformer_b200.engine import TrainStep
    x, y = _data()
    model = _model()
    step = TrainStep(model, lambda lg, yy: F.cross_entropy(lg, yy), x, y, lr=5e-2, use_graph=True, warmup=2)

    def eval_logits(m):
        m.eval()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            out = m(x).float()
        m.train()
        return out

    step()
    first = eval_logits(model)          # fills the eval cache
    for _ in range(3):
        step()                           # graph replays: no Python forward, no version bump
    second = eval_logits(model)
    fresh = og.build_model(CFG).to(DEV)
    fresh.load_state_dict(model.state_dict(), strict=True)
    want = eval_logits(fresh)
    torch.testing.assert_close(second, want, rtol=1e-3, atol=1e-3)
    assert not torch.allclose(first, second, rtol=1e-3, atol=1e-3)


def test_fused_droppath_table_feeds_every_droppath_once():
    from outlook_grid_vision_transformer_b200.engine import TrainStep
    cfg = dict(CFG, dpr_max=0.5)
    x, y = _data(B=8)
    model = _model(cfg=cfg)
    step = TrainStep(model, lambda lg, yy: F.cross_entropy(lg, yy), x, y, lr=1e-3, use_graph=True, warmup=2)
    assert step._dp_keep is not None and step._dp_keep.shape == (4, 8)  # block 0 has no DropPath; block 1 has 4
    l = [float(step()) for _ in range(3)]
    assert all(math.isfinite(v) for v in l)


def test_bulk_weight_refresh_equals_per_layer_casts():
    """TrainStep re-derives all bf16 weight copies in one launch per step (modules.refresh_prepared); a hand-written
    loop casts per layer.  Same losses step for step -- in particular the copies are never stale after AdamW."""
    from outlook_grid_vision_transformer_b200 import modules as M
    from outlook_grid_vision_transformer_b200.engine import TrainStep
    x, y = _data()
    losses = []
    for bulk in (False, True):
        model = _model()
        if bulk:
            step = TrainStep(model, lambda lg, yy: F.cross_entropy(lg, yy), x, y, lr=3e-3, weight_decay=0.05,
                             autocast_bf16=True, use_graph=False)
            losses.append([float(step()) for _ in range(5)])
            assert model.__dict__.get("_ogv_cast_batch") is not None and model.__dict__["_ogv_cast_batch"][1].n > 0
        else:
            opt = torch.optim.AdamW(_torch_groups(model, 0.05), lr=3e-3, fused=True)
            out = []
            for _ in range(5):
                opt.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    lg = model(x)
                loss = F.cross_entropy(lg.float(), y)
                loss.backward()
                opt.step()
                out.append(float(loss.detach()))
            losses.append(out)
        assert not M._BULK_FRESH
    assert losses[0] == pytest.approx(losses[1], rel=5e-3), f"{losses}"
    assert losses[1][-1] < losses[1][0]


def test_prefetched_batches_and_lagged_loss_equal_the_plain_loop():
    """TrainStep.prefetch / step(staged=True) / loss_to_host_async / previous_loss (the bench's e2e path, the reference's
    pinned-memory + non_blocking loader loop, one_epoch_train.py:85-93): every step trains on ITS batch and the loss
    read back with a one-step lag is the loss of the step before -- same numbers as copying each batch synchronously."""
    from outlook_grid_vision_transformer_b200.engine import TrainStep, WarmupCosineLR
    batches = [_data(seed=20 + i) for i in range(6)]
    hosts = [(x.cpu().pin_memory(), y.cpu().pin_memory()) for x, y in batches]
    mk = lambda: TrainStep(_model(), lambda lg, yy: F.cross_entropy(lg, yy, label_smoothing=0.1), *batches[0], lr=2e-3,
                           autocast_bf16=False, use_graph=True, warmup=2, grad_clip_norm=1.0, eps=1e-6,
                           scheduler=WarmupCosineLR(2e-3, 12, 3, 1e-5))  # noqa: E731
    plain = mk()
    want = [float(plain(xh, yh)) for xh, yh in hosts]
    piped = mk()
    got = []
    for xh, yh in hosts:
        piped.prefetch(xh, yh)
        piped(staged=True)
        piped.loss_to_host_async()
        got.append(piped.previous_loss())
    torch.cuda.synchronize()
    assert got[0] is None
    assert got[1:] == pytest.approx(want[:-1], rel=2e-4), f"{got} vs {want}"   # atomics reorder sums between runs
    for (k, p), (_, q) in zip(piped.model.named_parameters(), plain.model.named_parameters()):
        if k.endswith("mhsa.qkv.bias"):  # the key bias gradient is pure summation noise (softmax shift invariance)
            continue
        torch.testing.assert_close(p.detach(), q.detach(), rtol=2e-3, atol=2e-5, msg=lambda m: f"{k}: {m}")


def test_hyper_parameters_survive_a_host_running_ahead():
    """The host may queue many replays before the device starts the first one: every step must still see ITS learning rate
    and bias corrections (pinned staging slots in a ring guarded by events), i.e. 10 steps queued without a
    synchronisation produce the losses of 10 steps synchronised one by one (warm-up schedule: the LR changes 4x over the
    first steps, so a step that read a later step's slot shows up at once)."""
    from outlook_grid_vision_transformer_b200.engine import TrainStep, WarmupCosineLR
    x, y = _data()
    mk = lambda: TrainStep(_model(), lambda lg, yy: F.cross_entropy(lg, yy), x, y, lr=2e-3, autocast_bf16=False,
                           use_graph=True, warmup=2, eps=1e-6, grad_clip_norm=0.5,
                           scheduler=WarmupCosineLR(2e-3, 12, 4, 1e-5))  # noqa: E731
    a, b = mk(), mk()
    la, lb = [], []
    for _ in range(10):
        la.append(float(a()))          # synchronises every step
    for _ in range(10):
        lb.append(b().detach().clone())  # queued: no host synchronisation until the end
    torch.cuda.synchronize()
    lb = [float(t) for t in lb]
    assert lb == pytest.approx(la, rel=5e-4), f"{lb} vs {la}"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("smoothing", [0.0, 0.1])
@pytest.mark.parametrize("B,K", [(1024, 100), (7, 200), (33, 10), (4, 1000)])
def test_cross_entropy_kernel_matches_torch(B, K, smoothing, dtype):
    """og.cross_entropy (ogv_xent_fwd / ogv_xent_bwd) against F.cross_entropy: the criterion of the reference's loop
    (nn.CrossEntropyLoss(label_smoothing), train_full_model.py:52), loss and gradient, with an upstream gradient."""
    import outlook_grid_vision_transformer_b200 as og
    g = torch.Generator().manual_seed(B + K)
    x = (torch.randn(B, K, generator=g) * 3).to(DEV).to(dtype)
    y = torch.randint(0, K, (B,), generator=g).to(DEV)
    x1, x2 = x.clone().requires_grad_(True), x.float().clone().requires_grad_(True)
    l1 = og.cross_entropy(x1, y, label_smoothing=smoothing)
    l2 = F.cross_entropy(x2, y, label_smoothing=smoothing)
    (l1 * 1.7).backward()
    (l2 * 1.7).backward()
    assert l1.shape == l2.shape == ()
    rtol = 1e-5 if dtype == torch.float32 else 1e-4     # the loss itself is fp32 either way (bf16: same rounded inputs)
    assert float(l1) == pytest.approx(float(l2), rel=rtol)
    tol = dict(rtol=1e-4, atol=1e-8) if dtype == torch.float32 else dict(rtol=1e-2, atol=1e-6)   # bf16: the stored gradient
    torch.testing.assert_close(x1.grad.float(), x2.grad, **tol)
    # other signatures fall through to PyTorch
    w = og.cross_entropy(x.float().cpu(), y.cpu(), label_smoothing=smoothing)
    assert float(w) == pytest.approx(float(l2), rel=1e-4)
