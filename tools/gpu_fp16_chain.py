"""GPU diagnostic: error growth along the backbone chain (ours 16-bit vs torch 16-bit vs fp32)."""
import sys, copy
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import pixeltable_yolox_b200 as yx
from oracle import yolox_oracle as yo
from pixeltable_yolox_b200 import synthetic as syn
dev = torch.device("cuda", 0)
dt = getattr(torch, sys.argv[1]) if len(sys.argv) > 1 else torch.float16
cfg = yx.YoloxConfig.get_named_config("yolox_s"); cfg.model = None
model = cfg.get_model()
x = torch.from_numpy(syn.images(2, 320, 320, seed=11))
sd = yo.seeded_state_dict(model.state_dict(), 3, (320, 320), calib_x=x)
model.load_state_dict(sd)
bb32 = copy.deepcopy(model.backbone.backbone).to(dev).float().eval()
bb16 = copy.deepcopy(model.backbone.backbone).to(dev).to(dt).eval()
def rel(a, b): return ((a - b).abs() / b.abs().clamp_min(1.0))
def stats(e): return f"med {e.median().item():.2e} p99 {e.flatten().kthvalue(int(e.numel()*0.99)).values.item():.2e}"
with torch.no_grad():
    r = x.to(dev); o = x.to(dev); t = x.to(dev).to(dt)
    for name in ["stem", "dark2", "dark3", "dark4", "dark5"]:
        b32, b16 = getattr(bb32, name), getattr(bb16, name)
        # fp32 reference and torch-native 16-bit through the training-mode (pure torch) path with eval BN
        def torch_fwd(blk, v):
            if hasattr(blk, "_train_forward"): return blk._train_forward(v)
            for sub in blk: v = sub._train_forward(v)
            return v
        r = torch_fwd(b32, r)
        t = torch_fwd(b16, t)
        if name == "stem": o = b16(o)
        else:
            for sub in b16: o = sub(o)
        print(f"{name}: ours {stats(rel(o.float(), r))} | torch {stats(rel(t.float(), r))} | mean|ref| {r.abs().mean().item():.2f}", flush=True)
