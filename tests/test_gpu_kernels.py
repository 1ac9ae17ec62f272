"""-m gpu: each non-GEMM kernel of libogvit through its C-ABI entry point, against a float64 CPU
statement of the same formula (autograd supplies the backward answers).  fp32 tolerance 1e-4,
bf16 tolerance 2e-2 (relative to the tensor scale)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import outgrid_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
DT = [torch.float32, torch.bfloat16]
IDS = ["fp32", "bf16"]


def tol(dt):
    return 1e-4 if dt == torch.float32 else 2e-2


def close(got, want, t, what):
    got = got.double().cpu()
    want = want.double().cpu()
    assert got.shape == want.shape, f"{what}: {tuple(got.shape)} vs {tuple(want.shape)}"
    assert torch.isfinite(got).all(), f"{what}: non-finite"
    err = (got - want).abs().max() / (want.abs().max() + 1e-30)
    assert err < t, f"{what}: max rel err {err:.3e} >= {t}"


def dev(t, dt=None):
    return t.to(DEV, dt) if dt is not None else t.to(DEV)


# ------------------------------------------------------------------------------------------ layout
@pytest.mark.parametrize("dt", DT, ids=IDS)
def test_layout_roundtrip_bit_exact(dt):
    from outlook_grid_vision_transformer_b200 import ops
    x = torch.arange(3 * 24 * 5 * 7, dtype=torch.float32).reshape(3, 24, 5, 7).remainder(251).to(DEV, dt)
    rows = ops.nchw_to_rows(x)
    assert torch.equal(rows.cpu(), x.permute(0, 2, 3, 1).reshape(-1, 24).cpu())
    back = ops.rows_to_nchw(rows, 3, 24, 5, 7)
    assert torch.equal(back.cpu(), x.cpu())


@pytest.mark.parametrize("dt", DT, ids=IDS)
def test_rowscale_colsum(dt):
    from outlook_grid_vision_transformer_b200 import ops
    torch.manual_seed(9)
    B, P, C = 7, 36, 48
    x = torch.randn(B * P, C).to(dt)
    scale = (torch.rand(B) > 0.3).float() / 0.7
    out = torch.full((C,), 0.5, device=DEV)
    y = ops.rowscale_colsum(dev(x), dev(scale), P, out)
    want = x.double() * scale.double().repeat_interleave(P)[:, None]
    close(y, want, tol(dt), "rowscale_colsum y")
    close(out - 0.5, want.sum(0), tol(dt), "rowscale_colsum sum")


# --------------------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("dt", DT, ids=IDS)
@pytest.mark.parametrize("M,C", [(37, 16), (1000, 64), (513, 384), (64, 48)])
def test_layernorm_fwd_bwd(M, C, dt):
    from outlook_grid_vision_transformer_b200 import ops
    torch.manual_seed(M + C)
    x = (torch.randn(M, C) * 2 + 0.5).to(dt)
    w, b = torch.randn(C) * 0.3 + 1, torch.randn(C) * 0.1
    dy, dres = torch.randn(M, C).to(dt), torch.randn(M, C).to(dt)
    xr = x.double().requires_grad_(True)
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = O.layer_norm(xr, wr, br, 1e-5)
    (yr * dy.double()).sum().backward()
    y, mean, rstd = ops.layernorm_fwd(dev(x), dev(w), dev(b), 1e-5)
    close(y, yr.detach(), tol(dt), "ln fwd")
    dg, db = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    dx = ops.layernorm_bwd(dev(dy), dev(x), dev(w), mean, rstd, dev(dres), dg, db)
    close(dx, xr.grad + dres.double(), tol(dt), "ln dx")
    close(dg, wr.grad, tol(dt), "ln dgamma")
    close(db, br.grad, tol(dt), "ln dbeta")


# ------------------------------------------------------------------------------------ outlook core
def _outlook_core_ref(va, B, H, W, C, heads):
    hd = C // heads
    v = va[:, :C].reshape(B, H, W, C)
    A = torch.softmax(va[:, C:C + 9 * heads].reshape(B, H, W, heads, 9), dim=-1)
    y = torch.zeros(B, H, W, heads, hd, dtype=va.dtype)
    for t in range(9):
        ki, kj = divmod(t, 3)
        y = y + A[..., t].unsqueeze(-1) * O.shift2d(v, ki - 1, kj - 1).reshape(B, H, W, heads, hd)
    return y.reshape(B * H * W, C)


@pytest.mark.parametrize("dt", DT, ids=IDS)
@pytest.mark.parametrize("B,H,W,C,heads", [
    (2, 8, 8, 16, 4), (1, 6, 5, 48, 2), (2, 32, 32, 64, 2), (3, 4, 4, 384, 6),
    # the tiled bf16 kernels (W in {4, 8, 16, 32}, C % 64 == 0): every stage shape, ragged image groups (B not a multiple
    # of the images per tile), heights that are not a multiple of the tile rows, 16- / 32- / 64- / 128-wide heads
    (3, 16, 16, 128, 4), (5, 8, 8, 256, 8), (2, 12, 16, 64, 4), (9, 5, 8, 64, 1), (2, 20, 32, 128, 2), (11, 4, 4, 128, 1),
    (1, 3, 4, 64, 2), (2, 16, 16, 64, 2),
    # images wider than the 32-column tile (64 px stage 0 of cfg 3 / 4): column tiles with a real interior halo
    (2, 64, 64, 64, 2), (1, 7, 96, 128, 4)])
def test_outlook_core_fwd_bwd(B, H, W, C, heads, dt):
    from outlook_grid_vision_transformer_b200 import ops
    from outlook_grid_vision_transformer_b200.functional import outlook_npad
    torch.manual_seed(B * H + C)
    npad = outlook_npad(C, heads)
    M = B * H * W
    va = torch.zeros(M, npad)
    va[:, :C + 9 * heads] = torch.randn(M, C + 9 * heads)
    va = va.to(dt)
    dy = torch.randn(M, C).to(dt)
    var = va.double().requires_grad_(True)
    yr = _outlook_core_ref(var, B, H, W, C, heads)
    (yr * dy.double()).sum().backward()
    y = ops.outlook_core_fwd(dev(va), B, H, W, C, heads)
    close(y, yr.detach(), tol(dt), "outlook fwd")
    dva = ops.outlook_core_bwd(dev(va), dev(dy), B, H, W, C, heads)
    close(dva[:, :C], var.grad[:, :C], tol(dt), "outlook dv")
    close(dva[:, C:C + 9 * heads], var.grad[:, C:C + 9 * heads], tol(dt), "outlook dlogits")
    assert float(dva[:, C + 9 * heads:].abs().sum()) == 0.0, "padding columns must be zero"


@pytest.mark.parametrize("B,H,W,C,heads,dt", [(1, 5, 6, 8, 1, torch.float32), (3, 9, 8, 64, 2, torch.bfloat16),
                                              (2, 32, 32, 64, 2, torch.bfloat16), (9, 4, 4, 128, 2, torch.bfloat16)],
                         ids=["generic_fp32", "tiled_w8", "tiled_w32", "tiled_w4"])
def test_outlook_core_integer_indexing_bit_exact(B, H, W, C, heads, dt):
    """One-hot attention (huge logit on one tap) with integer v: the gather must pick exactly v[p + d_t]."""
    from outlook_grid_vision_transformer_b200 import ops
    from outlook_grid_vision_transformer_b200.functional import outlook_npad
    npad = outlook_npad(C, heads)
    M = B * H * W
    v = torch.arange(1, M * C + 1, dtype=torch.float32).reshape(M, C).remainder(97) + 1
    for t in range(9):
        va = torch.zeros(M, npad)
        va[:, :C] = v
        va[:, C:C + 9 * heads] = -1e4
        for hh in range(heads):
            va[:, C + 9 * hh + t] = 1e4
        y = ops.outlook_core_fwd(dev(va.to(dt)), B, H, W, C, heads).float().cpu()
        ki, kj = divmod(t, 3)
        want = O.shift2d(v.reshape(B, H, W, C), ki - 1, kj - 1).reshape(M, C)
        assert torch.equal(y, want), f"tap {t}"


# --------------------------------------------------------------------------------------- BatchNorm
@pytest.mark.parametrize("dt", DT, ids=IDS)
def test_batchnorm_pipeline(dt):
    from outlook_grid_vision_transformer_b200 import ops
    torch.manual_seed(0)
    M, C = 777, 48
    x = (torch.randn(M, C) * 1.5 + 0.3).to(dt)
    g, b = torch.rand(C) + 0.5, torch.randn(C) * 0.2
    rm, rv = torch.randn(C) * 0.1, torch.rand(C) + 0.5
    res, dy = torch.randn(M, C).to(dt), torch.randn(M, C).to(dt)
    xr = x.double().requires_grad_(True)
    gr, br = g.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = F.batch_norm(xr, rm.double().clone(), rv.double().clone(), gr, br, True, 0.1, 1e-5)
    (yr * dy.double()).sum().backward()
    st = torch.zeros(6, C, device=DEV)
    drm, drv = dev(rm.clone()), dev(rv.clone())
    ops.colstats(dev(x), st[0], st[1])
    ops.bn_finalize(st[0], st[1], dev(g), dev(b), drm, drv, st[2], st[3], st[4], st[5], M, 1e-5, 0.1, True)
    y = ops.bn_apply(dev(x), st[2], st[3], dev(res))
    close(y, yr.detach() + res.double(), tol(dt), "bn fwd")
    refm, refv = rm.double().clone(), rv.double().clone()
    F.batch_norm(x.double(), refm, refv, None, None, True, 0.1, 1e-5)
    close(drm, refm, 1e-4, "running_mean")
    close(drv, refv, 1e-4, "running_var")
    dg, db = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    ops.bn_bwd_reduce(dev(dy), dev(x), st[4], st[5], dg, db)
    dx = ops.bn_bwd_apply(dev(dy), dev(x), st[4], st[5], dev(g), dg, db)
    close(dg, gr.grad, tol(dt), "bn dgamma")
    close(db, br.grad, tol(dt), "bn dbeta")
    close(dx, xr.grad, tol(dt) * 2, "bn dx")


# ------------------------------------------------------------------------------ depthwise + SE path
def _dw_ref(e_pre, sc, sh, w, B, H, W):
    Cm = e_pre.shape[1]
    e = O.act_fn("silu", e_pre * sc + sh).reshape(B, H, W, Cm)
    d = torch.zeros_like(e)
    for t in range(9):
        ki, kj = divmod(t, 3)
        d = d + O.shift2d(e, ki - 1, kj - 1) * w[:, t]
    return d.reshape(B * H * W, Cm)


@pytest.mark.parametrize("dt", DT, ids=IDS)
@pytest.mark.parametrize("B,H,W,Cm", [(3, 8, 8, 32), (2, 32, 32, 64), (5, 4, 4, 96), (2, 16, 16, 40), (1, 5, 7, 96),
                                      (2, 64, 64, 32), (300, 32, 32, 32),
                                      # the TMA-tile walker forward (bf16, W in {4, 8, 16, 32}, Cm % 64 == 0): stage shapes,
                                      # ragged image groups, heights that are not a multiple of the tile rows, many tiles per CTA
                                      (3, 32, 32, 256), (2, 16, 16, 512), (5, 8, 8, 128), (11, 4, 4, 192), (2, 12, 16, 64),
                                      (9, 5, 8, 64), (1, 3, 4, 64), (700, 8, 8, 64),
                                      (2, 64, 64, 128), (1, 10, 96, 64)])
def test_dwconv_fwd_bwd(B, H, W, Cm, dt):
    from outlook_grid_vision_transformer_b200 import ops
    torch.manual_seed(B + H + Cm)
    M = B * H * W
    e_pre = torch.randn(M, Cm).to(dt)
    sc, sh = torch.rand(Cm) + 0.5, torch.randn(Cm) * 0.2
    mean1, rstd1 = torch.randn(Cm) * 0.1, torch.rand(Cm) + 0.5
    w = torch.randn(Cm, 9) * 0.3
    g = torch.randn(M, Cm).to(dt)
    er = e_pre.double().requires_grad_(True)
    wr = w.double().requires_grad_(True)
    dr = _dw_ref(er, sc.double(), sh.double(), wr, B, H, W)
    (dr * g.double()).sum().backward()
    s2, q2 = torch.zeros(Cm, device=DEV), torch.zeros(Cm, device=DEV)
    d = ops.dwconv_fwd(dev(e_pre), dev(sc), dev(sh), dev(w), s2, q2, B, H, W, "silu")
    close(d, dr.detach(), tol(dt), "dw fwd")
    # statistics describe the tensor as stored (rounded to the activation dtype)
    close(s2, d.double().sum(0), 1e-3, "dw stats sum")
    close(q2, (d.double() ** 2).sum(0), 1e-3, "dw stats sumsq")
    # backward: du1 = dL/d(u1) where u1 = sc*e_pre+sh  ->  dL/de_pre = du1 * sc
    dw_, dg1, db1 = torch.zeros(Cm, 9, device=DEV), torch.zeros(Cm, device=DEV), torch.zeros(Cm, device=DEV)
    du1 = ops.dwconv_bwd(dev(g), dev(e_pre), dev(sc), dev(sh), dev(mean1), dev(rstd1), dev(w), dw_, dg1, db1, B, H, W,
                         "silu")
    du_ref = er.grad / sc.double()
    close(du1, du_ref, tol(dt), "dw du1")
    close(dw_, wr.grad, tol(dt), "dw dfilter")
    close(db1, du_ref.sum(0), tol(dt), "dw dbeta1")
    xh = (e_pre.double() - mean1.double()) * rstd1.double()
    close(dg1, (du_ref * xh).sum(0), tol(dt), "dw dgamma1")


@pytest.mark.parametrize("dt", DT, ids=IDS)
@pytest.mark.parametrize("Cm", [40, 264])
def test_se_and_bn2_kernels(Cm, dt):
    from outlook_grid_vision_transformer_b200 import ops
    torch.manual_seed(4)
    B, HW, Cm = 3, 20, Cm
    M = B * HW
    d_pre = torch.randn(M, Cm).to(dt)
    sc, sh = torch.rand(Cm) + 0.5, torch.randn(Cm) * 0.2
    mean2, rstd2, gamma2 = torch.randn(Cm) * 0.1, torch.rand(Cm) + 0.5, torch.rand(Cm) + 0.5
    gate, dpool = torch.rand(B, Cm), torch.randn(B, Cm)
    dd_act = torch.randn(M, Cm).to(dt)
    dD = d_pre.double()
    act = O.act_fn("silu", dD * sc.double() + sh.double()).reshape(B, HW, Cm)
    pool = ops.se_pool(dev(d_pre), dev(sc), dev(sh), B, HW, "silu")
    close(pool, act.mean(1), tol(dt), "se_pool")
    d_act = ops.bn_act_gate(dev(d_pre), dev(sc), dev(sh), dev(gate), B, HW, "silu")
    close(d_act, (act * gate.double()[:, None, :]).reshape(M, Cm), tol(dt), "bn_act_gate")
    dgate = ops.se_bwd_reduce(dev(dd_act), dev(d_pre), dev(sc), dev(sh), B, HW, "silu")
    close(dgate, (dd_act.double().reshape(B, HW, Cm) * act).sum(1), tol(dt), "se_bwd_reduce")
    u = dD * sc.double() + sh.double()
    sig = torch.sigmoid(u)
    dsilu = sig * (1 + u * (1 - sig))
    dd = dd_act.double().reshape(B, HW, Cm) * gate.double()[:, None, :] + dpool.double()[:, None, :] / HW
    du = dd.reshape(M, Cm) * dsilu
    xh = (dD - mean2.double()) * rstd2.double()
    dg2, db2 = torch.zeros(Cm, device=DEV), torch.zeros(Cm, device=DEV)
    stats = ops.mbconv_bwd_stats(dev(dd_act), dev(d_pre), dev(sc), dev(sh), dev(mean2), dev(rstd2), B, HW, "silu")
    close(stats[0], (dd_act.double().reshape(B, HW, Cm) * act).sum(1), tol(dt), "bwd_stats dgate")
    ops.mbconv_bn2_finalize(stats, dev(gate), dev(dpool), dg2, db2, B, HW)
    close(db2, du.sum(0), tol(dt), "bn2 dbeta")
    close(dg2, (du * xh).sum(0), tol(dt), "bn2 dgamma")
    out = ops.dw_bn2_bwd_apply(dev(dd_act), dev(d_pre), dev(gate), dev(dpool), dev(sc), dev(sh), dev(mean2),
                               dev(rstd2), dev(gamma2), dg2, db2, B, HW, "silu")
    want = gamma2.double() * rstd2.double() * (du - du.sum(0) / M - xh * (du * xh).sum(0) / M)
    close(out, want, tol(dt) * 2, "bn2 dx")


# ---------------------------------------------------------------------------------- grid attention
def _ga_ref(qkv, B, H, W, C, heads, g):
    hd = C // heads
    idx = O.grid_index(B, H, W, g)
    Bg, N = idx.shape
    t = qkv[idx.reshape(-1)].reshape(Bg, N, 3, heads, hd)
    q, k, v = (t[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    attn = torch.softmax((q @ k.transpose(-2, -1)) * hd ** -0.5, dim=-1)
    o = (attn @ v).permute(0, 2, 1, 3).reshape(Bg * N, C)
    out = torch.zeros(B * H * W, C, dtype=qkv.dtype).index_copy(0, idx.reshape(-1), o)
    return out, attn


@pytest.mark.parametrize("dt", DT, ids=IDS)
@pytest.mark.parametrize("B,H,W,C,heads,g", [(2, 8, 8, 16, 4, 2), (2, 32, 32, 64, 2, 8), (3, 16, 16, 128, 4, 8),
                                             (2, 8, 8, 48, 2, 8), (2, 4, 4, 384, 6, 2), (1, 64, 64, 64, 2, 8),
                                             (2, 8, 16, 80, 2, 4), (1, 32, 64, 64, 2, 8), (3, 64, 64, 128, 2, 8),
                                             (1, 16, 32, 48, 2, 4)])
def test_grid_attention_fwd_bwd(B, H, W, C, heads, g, dt):
    from outlook_grid_vision_transformer_b200 import ops
    torch.manual_seed(C + g)
    M = B * H * W
    qkv = torch.randn(M, 3 * C).to(dt)
    do = torch.randn(M, C).to(dt)
    qr = qkv.double().requires_grad_(True)
    outr, attnr = _ga_ref(qr, B, H, W, C, heads, g)
    (outr * do.double()).sum().backward()
    out = ops.grid_attn_fwd(dev(qkv), B, H, W, C, heads, g)
    close(out, outr.detach(), tol(dt), "grid attn fwd")
    probs = ops.grid_attn_probs(dev(qkv), B, H, W, C, heads, g)
    close(probs, attnr.detach(), tol(dt), "grid attn probs")
    dqkv = ops.grid_attn_bwd(dev(qkv), dev(do), B, H, W, C, heads, g)
    close(dqkv, qr.grad, tol(dt) * 2, "grid attn dqkv")


def test_grid_attention_partition_is_bit_exact():
    """With q = k = 0 the softmax is uniform, so out = mean of v over the group: integer v and
    N a power of two make the expected value exact in fp32 -> checks the index arithmetic alone."""
    from outlook_grid_vision_transformer_b200 import ops
    B, H, W, C, heads, g = 2, 8, 8, 8, 2, 4
    M = B * H * W
    qkv = torch.zeros(M, 3 * C)
    qkv[:, 2 * C:] = (torch.arange(M * C).reshape(M, C) % 64).float() * 4
    out = ops.grid_attn_fwd(dev(qkv), B, H, W, C, heads, g).cpu()
    idx = O.grid_index(B, H, W, g)
    v = qkv[:, 2 * C:]
    want = torch.zeros(M, C)
    for grp in range(idx.shape[0]):
        want[idx[grp]] = v[idx[grp]].mean(0, keepdim=True)
    assert torch.equal(out, want)


# ------------------------------------------------------------------------------------ small kernels
def test_rowscale_colsum_cast_transpose_adamw():
    from outlook_grid_vision_transformer_b200 import ops
    torch.manual_seed(1)
    x = torch.randn(6 * 5, 16, device=DEV)
    s = torch.tensor([0., 1.25, 1.25, 0., 1.25, 1.25], device=DEV)
    y = ops.rowscale(x, s, 5)
    assert torch.equal(y, x * s.repeat_interleave(5)[:, None])
    out = torch.zeros(16, device=DEV)
    ops.colsum(x, out)
    close(out, x.double().sum(0), 1e-5, "colsum")
    w = torch.randn(37, 50, device=DEV)
    d = torch.zeros(40, 50, device=DEV, dtype=torch.bfloat16)
    dt_ = torch.zeros(50, 40, device=DEV, dtype=torch.bfloat16)
    ops.cast_transpose(w, d[:37], dt_[:, :37])
    assert torch.equal(d[:37], w.bfloat16()) and torch.equal(dt_[:, :37], w.bfloat16().t())
    assert float(d[37:].abs().sum()) == 0 and float(dt_[:, 37:].abs().sum()) == 0
    p = torch.randn(1000, device=DEV); g = torch.randn(1000, device=DEV)
    m = torch.zeros(1000, device=DEV); v = torch.zeros(1000, device=DEV)
    pr = p.clone().cpu().requires_grad_(True)
    opt = torch.optim.AdamW([pr], lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05)
    for step in (1, 2, 3):
        pr.grad = g.cpu().clone()
        opt.step()
        ops.adamw(p, g, m, v, 1e-2, 0.9, 0.999, 1e-8, 0.05, step)
    close(p, pr.detach(), 1e-5, "adamw")


@pytest.mark.parametrize("C", [64, 128, 384, 1536], ids=lambda c: f"C{c}")
@pytest.mark.parametrize("dt", DT, ids=IDS)
def test_column_reductions_and_streaming_passes_at_scale(dt, C):
    """Many CTAs, every fold path of the column-reduction skeleton (warp-shuffle pre-fold for C = 64 / 128,
    thread-row fold for C = 384 / 1536), rows that do not divide the per-CTA row count, and the kernels that keep
    several rows in flight per thread (LayerNorm, BatchNorm backward apply)."""
    from outlook_grid_vision_transformer_b200 import ops
    torch.manual_seed(C)
    M = 20011
    x = (torch.randn(M, C) * 1.3 + 0.4).to(dt)
    dy = torch.randn(M, C).to(dt)
    xd, dyd = x.double(), dy.double()
    f = lambda *s: torch.zeros(s, device=DEV)  # noqa: E731
    scale_t = 3e-4 if dt == torch.float32 else 2e-2  # sums of 2e4 terms: relative to the largest column result
    out = f(C)
    ops.colsum(dev(x), out)
    close(out, xd.sum(0), scale_t, "colsum")
    s, q = f(C), f(C)
    ops.colstats(dev(x), s, q)
    close(s, xd.sum(0), scale_t, "colstats sum")
    close(q, (xd * xd).sum(0), scale_t, "colstats sumsq")
    mean, var = xd.mean(0), xd.var(0, unbiased=False)
    rstd = (var + 1e-5).rsqrt()
    dg, db = f(C), f(C)
    ops.bn_bwd_reduce(dev(dy), dev(x), dev(mean.float()), dev(rstd.float()), dg, db)
    xh = (xd - mean) * rstd
    close(db, dyd.sum(0), scale_t, "bn_bwd_reduce dbeta")
    close(dg, (dyd * xh).sum(0), scale_t, "bn_bwd_reduce dgamma")
    g = torch.rand(C).double() + 0.5
    dx = ops.bn_bwd_apply(dev(dy), dev(x), dev(mean.float()), dev(rstd.float()), dev(g.float()), dg, db)
    want = g * rstd * (dyd - dyd.sum(0) / M - xh * (dyd * xh).sum(0) / M)
    close(dx, want, tol(dt) * 2, "bn_bwd_apply")
    if C <= 1024:
        w, b = torch.rand(C) + 0.5, torch.randn(C) * 0.1
        xr = xd.clone().requires_grad_(True)
        wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
        yr = F.layer_norm(xr, (C,), wr, br, 1e-5)
        (yr * dyd).sum().backward()
        y, mu, rs = ops.layernorm_fwd(dev(x), dev(w), dev(b), 1e-5)
        close(y, yr.detach(), tol(dt), "layernorm fwd")
        dgam, dbet = f(C), f(C)
        dxl = ops.layernorm_bwd(dev(dy), dev(x), dev(w), mu, rs, None, dgam, dbet)
        close(dxl, xr.grad, tol(dt) * 2, "layernorm dx")
        close(dgam, wr.grad, scale_t, "layernorm dgamma")
        close(dbet, br.grad, scale_t, "layernorm dbeta")


# ------------------------------------------------------------------------ fused squeeze-excite MLP
@pytest.mark.parametrize("B,Cm,Cs", [(37, 256, 64), (64, 512, 128), (9, 1024, 256), (8, 256, 16), (130, 768, 32)])
def test_se_mlp_fused_fwd_bwd(B, Cm, Cs):
    """ogv_se_mlp_fwd / _bwd (both 1x1 convs of SqueezeExcite, mbc_conv.py:17-27, in one kernel per direction) against
    the same math in fp64 on the bf16-rounded operands; bf16 outputs at 2e-2, fp32 outputs at 2e-3."""
    from outlook_grid_vision_transformer_b200 import _lib, ops
    assert _lib.lib().ogv_se_mlp_supported(Cm, Cs, _lib.BF16)
    g = torch.Generator().manual_seed(B + Cm)
    rb = lambda t: t.to(torch.bfloat16).double()  # noqa: E731
    pool = torch.randn(B, Cm, generator=g)
    w1 = torch.randn(Cs, Cm, generator=g) / Cm ** 0.5
    w2 = torch.randn(Cm, Cs, generator=g) / Cs ** 0.5
    b1, b2 = torch.randn(Cs, generator=g) * 0.1, torch.randn(Cm, generator=g) * 0.1
    w1b, w2b = w1.to(torch.bfloat16), w2.to(torch.bfloat16)
    pool_c, s1_pre, s1a, gate_pre, gate = ops.se_mlp_fwd(dev(pool), dev(w1b.t().contiguous()), dev(b1),
                                                          dev(w2b.t().contiguous()), dev(b2), "silu")
    z1 = rb(pool) @ w1b.double().t() + b1.double()
    a1 = torch.nn.functional.silu(z1)
    z2 = rb(a1) @ w2b.double().t() + b2.double()
    assert torch.equal(pool_c.cpu(), pool.to(torch.bfloat16))
    close(s1_pre, z1, 2e-2, "s1_pre")
    close(s1a, a1, 2e-2, "s1a")
    close(gate_pre, z2, 2e-3, "gate_pre")
    close(gate, torch.sigmoid(z2), 2e-3, "gate")
    # backward
    dgate = torch.randn(B, Cm, generator=g)
    dgate_c, ds1_pre, dpool = ops.se_mlp_bwd(dev(dgate), gate_pre, s1_pre, dev(w2b), dev(w1b), "silu")
    zz2 = gate_pre.double().cpu()
    sg = torch.sigmoid(zz2)
    dgp = dgate.double() * sg * (1 - sg)
    close(dgate_c, dgp, 2e-2, "dgate_c")
    zz1 = s1_pre.double().cpu()
    s1 = torch.sigmoid(zz1)
    dsilu = s1 * (1 + zz1 * (1 - s1))
    ds1 = (rb(dgp) @ w2b.double()) * dsilu
    close(ds1_pre, ds1, 2e-2, "ds1_pre")
    close(dpool, rb(ds1) @ w1b.double(), 5e-3, "dpool")
