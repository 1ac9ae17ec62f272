"""-m gpu: the pointwise-GEMM engines (tcgen05 bf16 and SIMT fp32) through ogv_gemm, against a
float64 CPU product of the same (bf16-rounded) operands.  bf16 products are exact in fp32, so the
only error is fp32 accumulation order: tolerance 2e-3 relative to the result scale (bf16 output
rounding dominates when the output dtype is bf16)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ref(A, B):
    return A.double().cpu() @ B.double().cpu().t()


def _check(D, want, tol, what):
    got = D.double().cpu()
    err = (got - want).abs().max() / (want.abs().max() + 1e-30)
    assert torch.isfinite(got).all(), f"{what}: non-finite"
    assert err < tol, f"{what}: max rel err {err:.3e} >= {tol}"


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 128, 64), (1000, 82, 64), (4096, 256, 64), (384, 64, 256),
                                   (777, 1536, 384), (2048, 96, 48), (512, 192, 128), (130, 16, 32)])
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32], ids=["obf16", "of32"])
def test_tc_gemm_kmajor(M, N, K, out_dtype):
    from outlook_grid_vision_transformer_b200 import ops
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device=DEV).bfloat16()
    B = torch.randn(N, K, device=DEV).bfloat16()
    D = torch.full((M, N), float("nan"), device=DEV, dtype=out_dtype)
    ops.gemm(A, B, D, engine=ops.ENGINE_TC)
    torch.cuda.synchronize()
    _check(D, _ref(A, B), 1e-2 if out_dtype == torch.bfloat16 else 2e-3, f"tc kmajor {M}x{N}x{K}")


@pytest.mark.parametrize("M,N,K", [(64, 64, 4096), (256, 64, 2048), (88, 64, 1000), (1536, 384, 512), (128, 256, 8192),
                                   (48, 192, 640)])
@pytest.mark.parametrize("split", [1, 4])
def test_tc_gemm_mnmajor_wgrad(M, N, K, split):
    """wgrad form: A(m,k) = dY[k,m], B(n,k) = X[k,n]; fp32 accumulate into a zeroed output."""
    from outlook_grid_vision_transformer_b200 import ops
    torch.manual_seed(M * 3 + N + K)
    dY = torch.randn(K, M, device=DEV).bfloat16()
    X = torch.randn(K, N, device=DEV).bfloat16()
    D = torch.zeros((M, N), device=DEV, dtype=torch.float32)
    ops.gemm(dY.t(), X.t(), D, accumulate=True, split_k=split, engine=ops.ENGINE_TC)
    torch.cuda.synchronize()
    _check(D, dY.double().cpu().t() @ X.double().cpu(), 2e-3, f"tc mnmajor {M}x{N}x{K} split={split}")


def test_tc_gemm_mixed_major():
    """dgrad form with an un-transposed weight: A K-major, B MN-major."""
    from outlook_grid_vision_transformer_b200 import ops
    torch.manual_seed(5)
    M, N, K = 512, 128, 256            # dX[M, N=in] = dY[M, K=out] @ W[K=out, N=in]
    dY = torch.randn(M, K, device=DEV).bfloat16()
    W = torch.randn(K, N, device=DEV).bfloat16()
    D = torch.empty((M, N), device=DEV, dtype=torch.float32)
    ops.gemm(dY, W.t(), D, engine=ops.ENGINE_TC)
    torch.cuda.synchronize()
    _check(D, dY.double().cpu() @ W.double().cpu(), 2e-3, "tc mixed major")


@pytest.mark.parametrize("engine_name", ["tc", "simt"])
def test_gemm_epilogue(engine_name):
    """bias, saved pre-activation, GELU, per-sample row scale, residual -- the fused forward epilogue."""
    from outlook_grid_vision_transformer_b200 import ops
    torch.manual_seed(11)
    M, N, K, P = 512, 128, 64, 64
    dt = torch.bfloat16 if engine_name == "tc" else torch.float32
    eng = ops.ENGINE_TC if engine_name == "tc" else ops.ENGINE_SIMT
    A = torch.randn(M, K, device=DEV).to(dt)
    B = (torch.randn(N, K, device=DEV) * 0.2).to(dt)
    bias = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV).to(dt)
    scale = (torch.rand(M // P, device=DEV) > 0.3).float() / 0.7
    D = torch.empty((M, N), device=DEV, dtype=dt)
    pre = torch.empty((M, N), device=DEV, dtype=dt)
    ops.gemm(A, B, D, bias=bias, pre_out=pre, act="gelu", row_scale=scale, rows_per_scale=P, residual=res, engine=eng)
    torch.cuda.synchronize()
    z = _ref(A, B) + bias.double().cpu()
    want = torch.nn.functional.gelu(z) * scale.double().cpu().repeat_interleave(P)[:, None] + res.double().cpu()
    tol = 1e-2 if dt == torch.bfloat16 else 1e-5
    _check(pre, z, tol, "pre_out")
    _check(D, want, tol, "epilogue output")
    # backward-style epilogue: multiply by act'(saved pre-activation)
    G = torch.randn(M, K, device=DEV).to(dt)
    D2 = torch.empty((M, N), device=DEV, dtype=dt)
    ops.gemm(G, B, D2, dact_src=pre, dact="gelu", engine=eng)
    torch.cuda.synchronize()
    zp = pre.double().cpu()
    dgelu = 0.5 * (1 + torch.erf(zp / 2 ** 0.5)) + zp * torch.exp(-0.5 * zp * zp) / (2 * torch.pi) ** 0.5
    _check(D2, _ref(G, B) * dgelu, tol, "dact epilogue")
    # saved-derivative form: forward stores act'(z), backward multiplies by it
    gp = torch.empty((M, N), device=DEV, dtype=dt)
    D3 = torch.empty((M, N), device=DEV, dtype=dt)
    ops.gemm(A, B, D3, bias=bias, pre_out=gp, act="gelu", pre_out_grad=True, engine=eng)
    torch.cuda.synchronize()
    dz = 0.5 * (1 + torch.erf(z / 2 ** 0.5)) + z * torch.exp(-0.5 * z * z) / (2 * torch.pi) ** 0.5
    _check(D3, torch.nn.functional.gelu(z), tol, "gelu output (saved-derivative form)")
    _check(gp, dz, tol, "saved gelu'")
    s1 = torch.zeros(N, device=DEV)
    ops.gemm(G, B, D2, dact_src=gp, dact="mul", col_sum=s1, engine=eng)
    torch.cuda.synchronize()
    _check(D2, _ref(G, B) * gp.double().cpu(), tol, "mul epilogue")
    _check(s1, D2.double().cpu().sum(0), 5e-3, "mul epilogue col_sum")


@pytest.mark.parametrize("M,N,K,ta,tb", [(100, 70, 33, False, False), (64, 64, 1000, True, True), (130, 20, 77, False, True)])
def test_simt_gemm_fp32(M, N, K, ta, tb):
    from outlook_grid_vision_transformer_b200 import ops
    torch.manual_seed(M + K)
    A = torch.randn(K, M, device=DEV).t() if ta else torch.randn(M, K, device=DEV)
    B = torch.randn(K, N, device=DEV).t() if tb else torch.randn(N, K, device=DEV)
    D = torch.zeros((M, N), device=DEV)
    ops.gemm(A, B, D, accumulate=True, split_k=3, engine=ops.ENGINE_SIMT)
    torch.cuda.synchronize()
    _check(D, _ref(A, B), 1e-5, "simt fp32")


def test_tc_gemm_large_multi_tile_persistent():
    """More tiles than SMs and several k-chunks: exercises the smem ring wrap-around and TMEM double buffer."""
    from outlook_grid_vision_transformer_b200 import ops
    torch.manual_seed(3)
    M, N, K = 148 * 128 * 3 + 77, 256, 320
    A = torch.randn(M, K, device=DEV).bfloat16()
    B = torch.randn(N, K, device=DEV).bfloat16()
    D = torch.empty((M, N), device=DEV, dtype=torch.bfloat16)
    ops.gemm(A, B, D, engine=ops.ENGINE_TC)
    torch.cuda.synchronize()
    want = (A.float() @ B.float().t())
    err = (D.float() - want).abs().max() / want.abs().max()
    assert err < 1e-2, f"rel err {err:.3e}"


@pytest.mark.parametrize("M,N,K", [(1000, 256, 64), (50000, 512, 128), (777, 82 * 0 + 88, 64), (300, 1536, 384)])
@pytest.mark.parametrize("engine_name", ["tc", "simt"])
def test_gemm_column_statistics(M, N, K, engine_name):
    """BatchNorm batch statistics / bias gradients out of the GEMM epilogue: col_sum, col_sumsq of D."""
    from outlook_grid_vision_transformer_b200 import ops
    torch.manual_seed(M + N)
    dt = torch.bfloat16 if engine_name == "tc" else torch.float32
    eng = ops.ENGINE_TC if engine_name == "tc" else ops.ENGINE_SIMT
    if engine_name == "simt" and M > 2000:
        pytest.skip("large case is for the persistent tcgen05 kernel")
    A = torch.randn(M, K, device=DEV).to(dt)
    B = (torch.randn(N, K, device=DEV) * 0.3).to(dt)
    bias = torch.randn(N, device=DEV)
    D = torch.empty((M, N), device=DEV, dtype=dt)
    s = torch.full((N,), 1.0, device=DEV)      # accumulated INTO (+=): starts at 1
    q = torch.full((N,), 2.0, device=DEV)
    ops.gemm(A, B, D, bias=bias, col_sum=s, col_sumsq=q, engine=eng)
    torch.cuda.synchronize()
    z = _ref(A, B) + bias.double().cpu()
    _check(D, z, 1e-2 if dt == torch.bfloat16 else 1e-5, "D")
    Dd = D.double().cpu()                      # statistics describe the tensor as stored
    _check(s - 1.0, Dd.sum(0), 2e-3, "col_sum")
    _check(q - 2.0, (Dd * Dd).sum(0), 2e-3, "col_sumsq")
    # sum only (bias gradient of a dgrad output)
    s2 = torch.zeros(N, device=DEV)
    ops.gemm(A, B, D, col_sum=s2, engine=eng)
    torch.cuda.synchronize()
    _check(s2, D.double().cpu().sum(0), 2e-3, "col_sum only")


@pytest.mark.parametrize("M,N,K", [(1000, 192, 48), (4096, 64, 256), (300, 88, 64)])
def test_fp32_gemm_on_tensor_cores_bf16x3(M, N, K):
    """fp32 operands, engine AUTO: three bf16 planes per operand on the tcgen05 engine (a*b ~= hi*hi + hi*lo + lo*hi).
    Error budget 2^-17 per product -> well inside the fp32 parity tolerance (1e-3); forward, dgrad-with-view and
    wgrad forms."""
    from outlook_grid_vision_transformer_b200 import ops
    torch.manual_seed(K + N)
    A = torch.randn(M, K, device=DEV)
    W = torch.randn(N, K, device=DEV) * 0.2
    bias = torch.randn(N, device=DEV)
    D = torch.empty(M, N, device=DEV)
    ops.gemm(A, W, D, bias=bias)
    _check(D, _ref(A, W) + bias.double().cpu(), 5e-5, "fp32 forward (bf16x3)")
    G = torch.randn(M, N, device=DEV)
    dA = torch.empty(M, K, device=DEV)
    ops.gemm(G, W.t(), dA)                      # dgrad with the un-transposed weight (MN-major view)
    _check(dA, G.double().cpu() @ W.double().cpu(), 5e-5, "fp32 dgrad (bf16x3)")
    dW = torch.zeros(N, K, device=DEV)
    ops.wgrad(G, A, dW)
    _check(dW, G.double().cpu().t() @ A.double().cpu(), 5e-5, "fp32 wgrad (bf16x3)")
    s = torch.zeros(N, device=DEV)
    ops.gemm(A, W, D, col_sum=s)
    _check(s, _ref(A, W).sum(0), 1e-4, "fp32 col_sum after bf16x3 product")


@pytest.mark.parametrize("M,N,K", [(1000, 256, 64), (4096 + 24, 88, 64), (2048, 192, 64), (8192, 512, 128), (3000, 384, 128),
                                   (2048, 1024, 256), (1024, 768, 256), (520, 1536, 384), (64 * 1024, 256, 64)])
def test_wgrad_bias_gradient_rides_with_the_weight_gradient(M, N, K):
    """ops.wgrad(dY, X, dW, bias_grad=db): dW += dY^T X and db += colsum(dY) in ONE launch (an extra N=16 MMA per K
    step against a tile of ones) == the separate column-sum pass, in bf16 (tcgen05) and fp32 (three-plane / FFMA)."""
    from outlook_grid_vision_transformer_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    for dt in (torch.bfloat16, torch.float32):
        dY = torch.randn(M, N, generator=g).to(DEV, dt)
        X = torch.randn(M, K, generator=g).to(DEV, dt)
        dW = torch.zeros(N, K, device=DEV)
        db = torch.full((N,), 0.5, device=DEV)  # accumulates on top of what is there
        ops.wgrad(dY, X, dW, bias_grad=db)
        torch.cuda.synchronize()
        want_w = dY.float().t() @ X.float()
        want_b = dY.float().sum(0) + 0.5
        tol = 2e-2 if dt == torch.bfloat16 else 1e-3
        torch.testing.assert_close(dW, want_w, rtol=tol, atol=tol * float(want_w.abs().max()))
        torch.testing.assert_close(db, want_b, rtol=1e-3, atol=1e-3 * float(want_b.abs().max()) + 1e-3)


@pytest.mark.parametrize("M,N,K", [(148 * 512, 64, 64), (148 * 512 + 300 * 128 + 37, 64, 256), (148 * 512 + 77, 40, 88),
                                   (1048576, 64, 64)])
def test_tc_gemm_two_subtile_work_items(M, N, K):
    """Narrow outputs with many rows run 256-row work items (two 128-row sub-tiles, all 16 epilogue warps busy): plain,
    bias + GELU + saved derivative, DropPath scale + residual, dgrad x saved derivative -- ragged M, N below the tile."""
    from outlook_grid_vision_transformer_b200 import ops
    g = torch.Generator().manual_seed(M % 1000 + N + K)
    A = torch.randn(M, K, generator=g).to(DEV, torch.bfloat16)
    B = (torch.randn(N, K, generator=g) * 0.2).to(DEV, torch.bfloat16)
    bias = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(M, N, generator=g).to(DEV, torch.bfloat16)
    P = 64
    scale = ((torch.arange((M + P - 1) // P, device=DEV) % 3 != 0).float() / 0.75).contiguous()
    acc = A.float() @ B.float().t()
    # plain
    D = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.gemm(A, B, D, engine=ops.ENGINE_TC)
    torch.testing.assert_close(D.float(), acc, rtol=1e-2, atol=1e-2 * float(acc.abs().max()))
    # bias + act + saved derivative
    pre = torch.empty_like(D)
    ops.gemm(A, B, D, bias=bias, act="gelu", pre_out=pre, pre_out_grad=True, engine=ops.ENGINE_TC)
    z = (acc + bias).requires_grad_(True)
    h = torch.nn.functional.gelu(z)
    (dh,) = torch.autograd.grad(h, z, torch.ones_like(h))
    torch.testing.assert_close(D.float(), h.detach(), rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(pre.float(), dh, rtol=2e-2, atol=2e-2)
    # bias + DropPath scale + residual
    ops.gemm(A, B, D, bias=bias, row_scale=scale, rows_per_scale=P, residual=res, engine=ops.ENGINE_TC)
    want = (acc + bias) * scale.repeat_interleave(P)[:M, None] + res.float()
    torch.testing.assert_close(D.float(), want, rtol=2e-2, atol=2e-2 * float(want.abs().max()))
    # dgrad x saved derivative
    ops.gemm(A, B, D, dact_src=res, dact="mul", engine=ops.ENGINE_TC)
    want = acc * res.float()
    torch.testing.assert_close(D.float(), want, rtol=2e-2, atol=2e-2 * float(want.abs().max()))
    torch.cuda.synchronize()


def test_fast_gelu_is_right_far_from_zero():
    """The bf16-mode GELU (tanh of an odd polynomial) must saturate correctly for large |x|: gelu(x) -> x, gelu(-x) -> 0,
    gelu' -> 1 / 0 (its polynomial fit is only valid for |x| <= 7 and is clamped there)."""
    from outlook_grid_vision_transformer_b200 import ops
    M = 4096
    x = torch.linspace(-60.0, 60.0, M * 64, device=DEV).reshape(M, 64).bfloat16()
    eye = torch.eye(64, device=DEV).bfloat16()
    D = torch.empty((M, 64), device=DEV, dtype=torch.bfloat16)
    pre = torch.empty_like(D)
    ops.gemm(x, eye, D, act="gelu", pre_out=pre, pre_out_grad=True, engine=ops.ENGINE_TC)
    z = x.float().requires_grad_(True)
    h = torch.nn.functional.gelu(z)
    (dh,) = torch.autograd.grad(h, z, torch.ones_like(h))
    torch.testing.assert_close(D.float(), h.detach(), rtol=1e-2, atol=2e-3)
    torch.testing.assert_close(pre.float(), dh, rtol=1e-2, atol=4e-3)
