"""Measure the REFERENCE's own bf16-autocast deviation from its float64 result on the whole-model
golden case, per parameter gradient (normwise relative error).  Run in the authoring container only:

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_bf16_noise.py

Writes tests/golden/ref_bf16_noise.json.  The GPU parity test uses it as the noise floor of a deep
bf16 gradient: the CUDA path must be within max(2e-2, 1.25 x the reference's own bf16 deviation).
(CPU autocast keeps softmax / LayerNorm in bf16, CUDA autocast promotes them to fp32 -- SURVEY
section 5 -- so this is an upper estimate of the reference-on-GPU noise; it is only applied to the
whole-model case, every per-module and per-block case keeps the plain 2e-2.)
"""
import json
import os
import sys
from pathlib import Path

import torch

REF = Path(os.environ.get("OGV_REFERENCE", "/root/reference"))
sys.dont_write_bytecode = True
sys.path.insert(0, str(REF))
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from src.Model_A_OutGridNet import MaxOutNet  # noqa: E402
from src.Model_B_OutGridNet import OutlookerFrontGridNet  # noqa: E402
from src.stage_config import StageCfg  # noqa: E402


def main():
    cases = torch.load(ROOT / "tests" / "golden" / "outgrid_golden.pt", map_location="cpu", weights_only=False)
    out = {}
    for name in ("model_a_tiny", "model_b_tiny_eval"):
        case = cases[name]
        mcfg = case["model_cfg"]
        stages = [StageCfg(**dict(s, drop_path=0.0)) for s in mcfg["stages"]]
        if str(mcfg.get("type", "model_a")).lower() in ("b", "model_b"):
            m = OutlookerFrontGridNet(num_classes=mcfg["num_classes"], stages=stages, stem_dim=mcfg["stem_dim"],
                                      outlooker_front_depth=int(mcfg.get("outlooker_front_depth", 2)), dpr_max=0.0)
        else:
            m = MaxOutNet(num_classes=mcfg["num_classes"], stages=stages, stem_dim=mcfg["stem_dim"], dpr_max=0.0)
        m.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in case["state"].items()}, strict=True)
        m.train(bool(case["training"]))
        x = case["x"].float().clone().requires_grad_(True)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            y = m(x)
        (y.float() * case["R"].float()).sum().backward()
        noise = {}
        for k, p in m.named_parameters():
            g = case["grads"][k].double()
            noise[k] = float((p.grad.double() - g).norm() / (g.norm() + 1e-3 * g.numel() ** 0.5))
        noise["__y__"] = float((y.double() - case["y"]).norm() / case["y"].norm())
        noise["__dx__"] = float((x.grad.double() - case["dx"].double()).norm() / case["dx"].double().norm())
        dxe = (x.grad.double() - case["dx"].double()).abs() / (case["dx"].double().abs() + case["dx"].double().abs().max())
        noise["__dx_max_band__"] = float(dxe.max())  # worst elementwise error in units of (|b| + max|b|)
        out[name] = noise
        worst = sorted(noise.items(), key=lambda kv: -kv[1])[:5]
        print(name, "worst:", worst)
    p = ROOT / "tests" / "golden" / "ref_bf16_noise.json"
    p.write_text(json.dumps(out, indent=1, sort_keys=True))
    print("wrote", p)


if __name__ == "__main__":
    main()
