#!/usr/bin/env python
"""Per-kernel micro-benchmark at the BASELINE cfg 2 stage shapes (14M, 32 px, batch 1024, bf16):
every C-ABI op timed alone with CUDA events, achieved GB/s = algorithmic bytes (each tensor argument
once) / mean launch time, against MEASURED_PEAKS.json.

    python tools/kbench.py [--only substr,substr] [--stages 0,1,2,3] [--batch 1024] [--reps 20]
"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from outlook_grid_vision_transformer_b200 import ops  # noqa: E402

STAGES = {  # C, H(=W), heads, grid, outlook heads
    0: dict(C=64, H=32, heads=2, g=8, oh=2),
    1: dict(C=128, H=16, heads=4, g=8, oh=4),
    2: dict(C=256, H=8, heads=8, g=4, oh=8),
    3: dict(C=384, H=4, heads=6, g=2, oh=6),
}


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def nbytes(*ts):
    return sum(t.numel() * t.element_size() for t in ts if t is not None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--stages", default="0,1,2,3")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "kbench.json"))
    a = ap.parse_args()
    only = [s for s in a.only.split(",") if s]
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    dev = "cuda:0"
    dt = torch.bfloat16
    B = a.batch
    rows = []

    def run(name, stage, fn, byts, flops=0):
        full = f"s{stage}:{name}"
        if only and not any(o in full for o in only):
            return
        try:
            ms = timeit(fn, a.reps)
        except Exception as exc:  # keep going: one broken kernel must not hide the others
            print(f"{full:34s} FAILED: {exc}")
            torch.cuda.synchronize()
            return
        gbs = byts / ms / 1e6
        rows.append(dict(kernel=full, ms=ms, GBps=gbs, frac=gbs / hbm, TFLOPs=flops / ms / 1e9))
        print(f"{full:34s} {ms * 1e3:9.1f} us  {gbs:8.1f} GB/s  {100 * gbs / hbm:5.1f}% HBM  {flops / ms / 1e9:7.1f} TF/s", flush=True)

    # box-to-box variation is +-20 %: print this box's raw stream rates next to the kernel numbers
    n_ref = 1 << 28
    xr = torch.empty(n_ref, device=dev, dtype=dt)
    yr = torch.empty(n_ref, device=dev, dtype=dt)
    ms = timeit(lambda: yr.copy_(xr), 10)
    print(f"[box] copy (read+write) {2 * n_ref * 2 / ms / 1e6:7.0f} GB/s", end="   ")
    ms = timeit(lambda: xr.fill_(1.0), 10)
    print(f"fill (write only) {n_ref * 2 / ms / 1e6:7.0f} GB/s   (MEASURED_PEAKS hbm {hbm:.0f})", flush=True)
    del xr, yr
    for s in [int(x) for x in a.stages.split(",")]:
        st = STAGES[s]
        C, H, W = st["C"], st["H"], st["H"]
        Cm, P, M = 4 * C, H * W, B * H * W
        f32 = dict(device=dev, dtype=torch.float32)
        x = torch.randn(M, C, device=dev).to(dt)
        dy = torch.randn(M, C, device=dev).to(dt)
        wide = torch.randn(M, Cm, device=dev).to(dt)
        wide2 = torch.randn(M, Cm, device=dev).to(dt)
        ones_c, zeros_c = torch.ones(C, **f32), torch.zeros(C, **f32)
        sc, sh = torch.rand(Cm, **f32) + 0.5, torch.randn(Cm, **f32) * 0.1
        mu, rs, gm = torch.randn(Cm, **f32) * 0.1, torch.rand(Cm, **f32) + 0.5, torch.rand(Cm, **f32) + 0.5
        gate, dpool = torch.rand(B, Cm, **f32), torch.randn(B, Cm, **f32)
        wdw = torch.randn(Cm, 9, **f32) * 0.3
        acc2 = torch.zeros(2, Cm, **f32)
        # ---- LayerNorm
        y, mean, rstd = ops.layernorm_fwd(x, ones_c, zeros_c, 1e-5)
        run("layernorm_fwd", s, lambda: ops.layernorm_fwd(x, ones_c, zeros_c, 1e-5), nbytes(x, x))
        dg, db = torch.zeros(C, **f32), torch.zeros(C, **f32)
        run("layernorm_bwd", s, lambda: ops.layernorm_bwd(dy, x, ones_c, mean, rstd, dy, dg, db), nbytes(x, x, x, x))
        # ---- reductions / elementwise on [M, C]
        run("colsum[M,C]", s, lambda: ops.colsum(dy, dg), nbytes(dy))
        run("colsum[M,4C]", s, lambda: ops.colsum(wide, acc2[0]), nbytes(wide))
        run("colstats[M,4C]", s, lambda: ops.colstats(wide, acc2[0], acc2[1]), nbytes(wide))
        scale_b = torch.rand(B, **f32)
        run("rowscale[M,C]", s, lambda: ops.rowscale(dy, scale_b, P), nbytes(dy, dy))
        run("rowscale_colsum[M,C]", s, lambda: ops.rowscale_colsum(dy, scale_b, P, dg), nbytes(dy, dy))
        run("rowscale_colsum[M,C] no store", s, lambda: ops.rowscale_colsum(dy, scale_b, P, dg, store=False), nbytes(dy))
        run("bn_apply[M,C]", s, lambda: ops.bn_apply(x, ones_c, zeros_c, dy), nbytes(x, x, x))
        run("bn_bwd_reduce[M,C]", s, lambda: ops.bn_bwd_reduce(dy, x, zeros_c, ones_c, dg, db), nbytes(x, x))
        run("colstats[M,C]", s, lambda: ops.colstats(x, dg, db), nbytes(x))
        run("bn_act_bwd_reduce[M,C]", s, lambda: ops.bn_act_bwd_reduce(dy, x, ones_c, zeros_c, zeros_c, ones_c, dg, db, "silu"), nbytes(x, x))
        run("bn_bwd_apply[M,C]", s, lambda: ops.bn_bwd_apply(dy, x, zeros_c, ones_c, ones_c, dg, db), nbytes(x, x, x))
        run("bn_bwd_apply[M,4C]", s, lambda: ops.bn_bwd_apply(wide, wide2, mu, rs, gm, acc2[0], acc2[1]), nbytes(wide, wide, wide))
        # ---- MBConv interior
        run("dwconv_fwd", s, lambda: ops.dwconv_fwd(wide, sc, sh, wdw, acc2[0], acc2[1], B, H, W, "silu"), nbytes(wide, wide))
        dw_ = torch.zeros(Cm, 9, **f32)
        run("dwconv_bwd", s, lambda: ops.dwconv_bwd(wide2, wide, sc, sh, mu, rs, wdw, dw_, acc2[0], acc2[1], B, H, W, "silu"),
            nbytes(wide, wide, wide))
        run("se_pool", s, lambda: ops.se_pool(wide, sc, sh, B, P, "silu"), nbytes(wide))
        run("bn_act_gate", s, lambda: ops.bn_act_gate(wide, sc, sh, gate, B, P, "silu"), nbytes(wide, wide))
        run("mbconv_bwd_stats", s, lambda: ops.mbconv_bwd_stats(wide2, wide, sc, sh, mu, rs, B, P, "silu"), nbytes(wide, wide))
        stats = ops.mbconv_bwd_stats(wide2, wide, sc, sh, mu, rs, B, P, "silu")
        run("mbconv_bn2_finalize", s, lambda: ops.mbconv_bn2_finalize(stats, gate, dpool, acc2[0], acc2[1], B, P), nbytes(stats, gate, dpool))
        run("dw_bn2_bwd_apply", s, lambda: ops.dw_bn2_bwd_apply(wide2, wide, gate, dpool, sc, sh, mu, rs, gm, acc2[0], acc2[1], B, P, "silu"),
            nbytes(wide, wide, wide))
        del wide2
        # ---- outlook core
        npad = (C + 9 * st["oh"] + 7) // 8 * 8
        va = torch.randn(M, npad, device=dev).to(dt)
        run("outlook_core_fwd", s, lambda: ops.outlook_core_fwd(va, B, H, W, C, st["oh"]), nbytes(va, x))
        run("outlook_core_bwd", s, lambda: ops.outlook_core_bwd(va, dy, B, H, W, C, st["oh"]), nbytes(va, x, va))
        del va
        # ---- grid attention
        qkv = torch.randn(M, 3 * C, device=dev).to(dt)
        run("grid_attn_fwd", s, lambda: ops.grid_attn_fwd(qkv, B, H, W, C, st["heads"], st["g"]), nbytes(qkv, x))
        run("grid_attn_bwd", s, lambda: ops.grid_attn_bwd(qkv, dy, B, H, W, C, st["heads"], st["g"]), nbytes(qkv, x, qkv))
        del qkv
        # ---- GEMMs (forward / dgrad / wgrad of the MLP and MBConv 1x1s)
        w1 = (torch.randn(Cm, C, device=dev) * 0.05).to(dt)
        w1t = w1.t().contiguous()
        b1 = torch.zeros(Cm, **f32)
        bC = torch.zeros(C, **f32)
        z, h = torch.empty(M, Cm, device=dev, dtype=dt), torch.empty(M, Cm, device=dev, dtype=dt)
        out_c = torch.empty(M, C, device=dev, dtype=dt)
        run("gemm fc1+gelu (z,h)", s, lambda: ops.gemm(x, w1, h, bias=b1, pre_out=z, act="gelu"), nbytes(x, z, h), 2 * M * Cm * C)
        run("gemm expand (plain)", s, lambda: ops.gemm(x, w1, h), nbytes(x, h), 2 * M * Cm * C)
        cs, cq = torch.zeros(Cm, **f32), torch.zeros(Cm, **f32)
        run("gemm expand +colstats", s, lambda: ops.gemm(x, w1, h, col_sum=cs, col_sumsq=cq), nbytes(x, h), 2 * M * Cm * C)
        run("gemm dgrad mul +colsum", s, lambda: ops.gemm(dy, w1, h, dact_src=z, dact="mul", col_sum=cs), nbytes(dy, z, h), 2 * M * Cm * C)
        run("gemm fc1+gelu (g',h)", s, lambda: ops.gemm(x, w1, h, bias=b1, pre_out=z, act="gelu", pre_out_grad=True), nbytes(x, z, h), 2 * M * Cm * C)
        run("gemm fc2+res", s, lambda: ops.gemm(wide, w1t, out_c, bias=bC, residual=x), nbytes(wide, x, out_c), 2 * M * Cm * C)
        run("gemm project (plain)", s, lambda: ops.gemm(wide, w1t, out_c), nbytes(wide, out_c), 2 * M * Cm * C)
        run("gemm dgrad fc2 (gelu')", s, lambda: ops.gemm(dy, w1, h, dact_src=z, dact="gelu"), nbytes(dy, z, h), 2 * M * Cm * C)
        run("gemm dgrad fc1", s, lambda: ops.gemm(wide, w1t, out_c), nbytes(wide, out_c), 2 * M * Cm * C)
        dW = torch.zeros(Cm, C, **f32)
        run("wgrad [4C,C]", s, lambda: ops.wgrad(wide, x, dW), nbytes(wide, x), 2 * M * Cm * C)
        dW2 = torch.zeros(C, Cm, **f32)
        run("wgrad [C,4C]", s, lambda: ops.wgrad(dy, wide, dW2), nbytes(wide, dy), 2 * M * Cm * C)
        wq = (torch.randn(3 * C, C, device=dev) * 0.05).to(dt)
        q3 = torch.empty(M, 3 * C, device=dev, dtype=dt)
        run("gemm qkv", s, lambda: ops.gemm(x, wq, q3, bias=torch.zeros(3 * C, **f32)), nbytes(x, q3), 2 * M * 3 * C * C)
        wp = (torch.randn(C, C, device=dev) * 0.05).to(dt)
        run("gemm proj+res", s, lambda: ops.gemm(x, wp, out_c, bias=bC, residual=dy), nbytes(x, dy, out_c), 2 * M * C * C)
        # ---- the squeeze-excite GEMMs on [B, Cm] rows: tcgen05 engine (bf16 operands) vs FFMA engine (fp32 operands)
        Cs = C
        pool32 = torch.randn(B, Cm, **f32)
        pool16 = pool32.to(dt)
        ws1_32 = torch.randn(Cs, Cm, **f32) * 0.05
        ws2_32 = torch.randn(Cm, Cs, **f32) * 0.05
        ws1_16, ws2_16 = ws1_32.to(dt), ws2_32.to(dt)
        s1_16, s1p_16 = torch.empty(B, Cs, device=dev, dtype=dt), torch.empty(B, Cs, device=dev, dtype=dt)
        s1_32, s1p_32 = torch.empty(B, Cs, **f32), torch.empty(B, Cs, **f32)
        g32, gp32 = torch.empty(B, Cm, **f32), torch.empty(B, Cm, **f32)
        bs1, bs2 = torch.zeros(Cs, **f32), torch.zeros(Cm, **f32)
        run("se fc1 tc", s, lambda: ops.gemm(pool16, ws1_16, s1_16, bias=bs1, pre_out=s1p_16, act="silu"), nbytes(pool16, ws1_16), 2 * B * Cm * Cs)
        run("se fc1 simt", s, lambda: ops.gemm(pool32, ws1_32, s1_32, bias=bs1, pre_out=s1p_32, act="silu", engine=ops.ENGINE_SIMT), nbytes(pool32, ws1_32), 2 * B * Cm * Cs)
        run("se fc2 tc", s, lambda: ops.gemm(s1_16, ws2_16, g32, bias=bs2, pre_out=gp32, act="sigmoid"), nbytes(g32, ws2_16), 2 * B * Cm * Cs)
        run("se fc2 simt", s, lambda: ops.gemm(s1_32, ws2_32, g32, bias=bs2, pre_out=gp32, act="sigmoid", engine=ops.ENGINE_SIMT), nbytes(g32, ws2_32), 2 * B * Cm * Cs)
        dWs = torch.zeros(Cm, Cs, **f32)
        g16 = g32.to(dt)
        run("se wgrad tc", s, lambda: ops.wgrad(g16, s1_16, dWs), nbytes(g16, s1_16), 2 * B * Cm * Cs)
        run("se wgrad simt", s, lambda: ops.wgrad(g32, s1_32, dWs, engine=ops.ENGINE_SIMT), nbytes(g32, s1_32), 2 * B * Cm * Cs)
        run("se cast", s, lambda: ops.cast(pool32, dt), nbytes(pool32, pool16))
        if ops.se_mlp_supported(Cm, Cs, dt):
            w1t16, w2t16 = ws1_16.t().contiguous(), ws2_16.t().contiguous()
            run("se mlp fused fwd", s, lambda: ops.se_mlp_fwd(pool32, w1t16, bs1, w2t16, bs2, "silu"), nbytes(pool32, g32, gp32), 4 * B * Cm * Cs)
            run("se mlp fused bwd", s, lambda: ops.se_mlp_bwd(g32, gp32, s1p_16, ws2_16, ws1_16, "silu"), nbytes(pool32, g32, gp32), 4 * B * Cm * Cs)
        # ---- fused MLP (hidden tile on chip): fc1 -> act -> fc2 -> +res in one kernel; backward recomputes it
        for tag, Hd in (("mlp 4C", 4 * C), ("mlp2d 2C", 2 * C)):
            if not ops.mlp_fused_supported(C, Hd, dt):
                continue
            wa = (torch.randn(Hd, C, device=dev) * 0.05).to(dt)
            wb = (torch.randn(C, Hd, device=dev) * 0.05).to(dt)
            wat, wbt = wa.t().contiguous(), wb.t().contiguous()
            ba = torch.zeros(Hd, **f32)
            run(f"fused {tag} fwd", s, lambda: ops.mlp_fwd(x, wa, ba, wb, bC, act="gelu", residual=dy), nbytes(x, dy, out_c), 4 * M * Hd * C)
            run(f"fused {tag} bwd", s, lambda: ops.mlp_bwd(x, dy, wa, wbt, wat, ba, act="gelu"),
                nbytes(x, dy, out_c) + 2 * M * Hd * 2, 6 * M * Hd * C)
        # ---- Downsample convolution C -> 2C, 3x3 stride 2 (downsampling.py:41-47): implicit GEMM / materialised patches /
        #      the library's convolution, all three directions
        if s < 3:
            Co, Ho = STAGES[s + 1]["C"], H // 2
            Mo = B * Ho * Ho
            xi = x.view(B, H, W, C).permute(0, 3, 1, 2)                      # channels_last image
            w4 = (torch.randn(Co, C, 3, 3, device=dev) * 0.05)
            w2 = w4.permute(0, 2, 3, 1).reshape(Co, 9 * C).to(dt).contiguous()
            w2t = w2.t().contiguous()
            dyo = torch.randn(Mo, Co, device=dev).to(dt)
            dw2 = torch.zeros(Co, 9 * C, **f32)
            cfl = 2 * Mo * Co * 9 * C
            if ops.conv3x3_supported(xi, Co, 2):
                run("conv implicit fwd", s, lambda: ops.conv3x3_fwd(xi, w2, 2), nbytes(x, dyo), cfl)
                run("conv implicit wgrad", s, lambda: ops.conv3x3_wgrad(xi, dyo, dw2, 2), nbytes(x, dyo), cfl)
            cols = ops.im2col3x3_vec(xi, 2)
            yo = torch.empty(Mo, Co, device=dev, dtype=dt)
            run("conv im2col", s, lambda: ops.im2col3x3_vec(xi, 2), nbytes(x, cols))
            run("conv patches fwd gemm", s, lambda: ops.gemm(cols, w2, yo), nbytes(cols, yo), cfl)
            run("conv patches wgrad", s, lambda: ops._wgrad(dyo, cols, dw2), nbytes(cols, dyo), cfl)
            dcols = torch.empty_like(cols)
            run("conv dgrad gemm", s, lambda: ops.gemm(dyo, w2t, dcols), nbytes(dcols, dyo), cfl)
            run("conv dgrad col2im", s, lambda: ops.col2im3x3_vec(dcols, B, H, W, C, 2), nbytes(dcols, x))
            wl = w4.to(dt).contiguous(memory_format=torch.channels_last)
            dyo4 = dyo.view(B, Ho, Ho, Co).permute(0, 3, 1, 2)
            conv = torch.ops.aten.convolution
            convb = torch.ops.aten.convolution_backward
            run("conv library fwd", s, lambda: conv(xi, wl, None, [2, 2], [1, 1], [1, 1], False, [0, 0], 1), nbytes(x, dyo), cfl)
            run("conv library dgrad", s, lambda: convb(dyo4, xi, wl, None, [2, 2], [1, 1], [1, 1], False, [0, 0], 1, [True, False, False]), nbytes(x, dyo), cfl)
            run("conv library wgrad", s, lambda: convb(dyo4, xi, wl, None, [2, 2], [1, 1], [1, 1], False, [0, 0], 1, [False, True, False]), nbytes(x, dyo), cfl)
            del cols, dcols, yo, dyo
        del x, dy, wide, z, h, out_c, q3
        torch.cuda.empty_cache()
    Path(a.out).parent.mkdir(exist_ok=True)
    Path(a.out).write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
