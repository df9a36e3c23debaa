"""GPU diagnostic: every named config (nano ... x) through the bf16/fp16 tcgen05 path at its test size, compared
with the oracle's fp32 forward and with plain-torch 16-bit arithmetic (the reference's own .bfloat16()/.half())."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import pixeltable_yolox_b200 as yx  # noqa: E402
from oracle import yolox_oracle as yo  # noqa: E402
from pixeltable_yolox_b200 import synthetic as syn  # noqa: E402

dev = torch.device("cuda", 0)
names = sys.argv[1:] or ["yolox_nano", "yolox_tiny", "yolox_s", "yolox_m", "yolox_l", "yolox_x"]
for name in names:
    cfg = yx.YoloxConfig.get_named_config(name)
    cfg.model = None
    model = cfg.get_model()
    h, w = cfg.test_size
    x = torch.from_numpy(syn.images(2, h, w, seed=11))
    sd = yo.seeded_state_dict(model.state_dict(), 3, (h, w), calib_x=x)
    model.load_state_dict(sd)
    ref = yo.forward({k: v.to(dev) for k, v in sd.items()}, x.to(dev)).cpu().numpy()       # fp32 torch, same graph
    for dt in (torch.bfloat16, torch.float16):
        model = model.float()
        model.load_state_dict(sd)            # nn.Module.to is in place: reload, or fp16 would inherit bf16-rounded weights
        m = model.to(dev).to(dt).eval()
        out = m(x.to(dev)).float().cpu().numpy()
        sdd = {k: (v.to(dev).to(dt) if v.is_floating_point() else v.to(dev)) for k, v in sd.items()}
        a = yo.ACTS["silu"]
        with torch.no_grad():
            o, _ = yo.head(sdd, yo.pafpn(sdd, x.to(dev).to(dt), a, depthwise=cfg.depthwise) if "depthwise" in yo.pafpn.__code__.co_varnames else yo.pafpn(sdd, x.to(dev).to(dt), a), a)
        nat = o.float().cpu().numpy()
        rel = lambda p, q: np.abs(p - q) / np.maximum(np.abs(q), 1.0)
        e_ours, e_nat = rel(out, ref), rel(nat, ref)
        print(f"{name:11s} {str(dt)[6:]:9s} {h}x{w} finite={np.isfinite(out).all()}  ours vs fp32: med {np.median(e_ours):.2e} p99 {np.quantile(e_ours, .99):.2e}"
              f" | torch-16bit vs fp32: med {np.median(e_nat):.2e} p99 {np.quantile(e_nat, .99):.2e}  launches={m.engine_for(x.to(dev)).launches}", flush=True)
        m.invalidate_engine()
    model = model.float().cpu()
