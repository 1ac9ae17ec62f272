#!/usr/bin/env python
"""Context number (NOT the bench's reference arm): the same model / parameters run through plain PyTorch
eager ops (cuDNN / cuBLAS / ATen kernels), i.e. what the reference's own code path does on a B200 --
conv2d 1x1, F.unfold, softmax, BatchNorm2d, LayerNorm, bmm -- under bf16 autocast with channels_last
and cudnn.benchmark, like scripts/train.py.  /root/reference cannot travel to the GPU box, so the
forward bodies below restate the reference modules' ATen op sequence on OUR parameter containers (the
reference adds a few more permute/contiguous copies; this is an upper bound on its eager speed).

    python tools/torch_eager_bench.py [--batch 1024] [--steps 5] [--workload cfg2_14m_32_bf16]
"""
import argparse
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import outlook_grid_vision_transformer_b200 as og  # noqa: E402
from outlook_grid_vision_transformer_b200 import modules as M  # noqa: E402
from outlook_grid_vision_transformer_b200.config import BASELINE_CONFIGS, CONFIG_DIR  # noqa: E402


def drop_path(dp, x):
    if not isinstance(dp, M.DropPath) or dp.drop_prob == 0.0 or not dp.training:
        return x
    keep = 1.0 - dp.drop_prob
    mask = torch.empty((x.shape[0],) + (1,) * (x.dim() - 1), device=x.device, dtype=x.dtype).bernoulli_(keep)
    return x * mask / keep


def ln2d(self, x):
    return self.ln(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2).contiguous()


def outlook_attn(self, x):
    B, C, H, W = x.shape
    h, hd = self.num_heads, C // self.num_heads
    a = self.attn(x).reshape(B, h, 9, H * W).softmax(dim=2)
    v = self.v(x)
    vu = F.unfold(v, kernel_size=3, padding=1).reshape(B, h, hd, 9, H * W)
    y = (vu * a.unsqueeze(2)).sum(dim=3).reshape(B, C, H, W)
    return self.proj(y)


def mlp2d(self, x):
    return self.fc2(self.act(self.fc1(x)))


def outlooker(self, x):
    x = x + drop_path(self.dp1, self.attn(self.norm1(x)))
    return x + drop_path(self.dp2, self.mlp(self.norm2(x)))


def se(self, x):
    s = self.pool(x)
    return x * self.gate(self.fc2(self.act(self.fc1(s))))


def mbconv(self, x):
    out = self.project(self.se(self.depthwise(self.expand(x))))
    return x + out if self.use_res else out


def mhsa(self, x):
    Bg, N, C = x.shape
    h = self.num_heads
    qkv = self.qkv(x).reshape(Bg, N, 3, h, C // h).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = ((q @ k.transpose(-2, -1)) * self.scale).softmax(dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(Bg, N, C)
    return self.proj(out)


def grid_attn(self, x):
    B, H, W, C = x.shape
    grids, meta = M.grid_partition(x, self.cfg.grid_size)
    Bg, Hg, Wg, _ = grids.shape
    y = self.mhsa(grids.reshape(Bg, Hg * Wg, C)).reshape(Bg, Hg, Wg, C)
    return M.grid_unpartition(y, meta)


def mlp(self, x):
    return self.fc2(self.act(self.fc1(x)))


def outgrid_block(self, x):
    x = self.outlook(x)
    x = self.mbconv(x)
    x = x.permute(0, 2, 3, 1).contiguous()
    x = x + drop_path(self.dp2, self.grid_attn(self.norm2(x)))
    x = x + drop_path(self.dp3, self.mlp(self.norm3(x)))
    return x.permute(0, 3, 1, 2).contiguous()


PATCH = {M.LayerNorm2d: ln2d, M.OutlookAttention2d: outlook_attn, M.MLP2d: mlp2d, M.OutlookerBlock2d: outlooker,
         M.SqueezeExcite: se, M.MBConv: mbconv, M.MultiHeadSelfAttention: mhsa, M.GridAttention2D: grid_attn,
         M.MLP: mlp, M.OutGridBlock: outgrid_block}


def to_eager(model):
    for m in model.modules():
        fn = PATCH.get(type(m))
        if fn is not None:
            m.forward = types.MethodType(fn, m)
    return model


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2_14m_32_bf16")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    wl = BASELINE_CONFIGS[a.workload]
    batch = a.batch or wl["batch"]
    cfg = og.load_yaml(CONFIG_DIR / wl["yaml"])["model"]
    torch.manual_seed(7)
    torch.backends.cudnn.benchmark = True
    dev = "cuda:0"
    model = to_eager(og.build_model(cfg)).to(dev).to(memory_format=torch.channels_last).train()
    opt = torch.optim.AdamW(model.parameters(), lr=5e-4, weight_decay=0.05, fused=True)
    x = torch.randn(batch, 3, wl["img"], wl["img"], device=dev).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, int(cfg.get("num_classes", 100)), (batch,), device=dev)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=wl["dtype"] == "bf16"):
            logits = model(x)
        loss = F.cross_entropy(logits.float(), y, label_smoothing=0.1)
        loss.backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(f"torch-eager restatement ({a.workload}, batch {batch}): {ms:.1f} ms/step = {batch / ms * 1e3:.0f} img/s, "
          f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")


if __name__ == "__main__":
    main()
