"""GPU diagnostic: per-block fp16 error of the tcgen05 path against torch fp32 on the same (fp16-rounded) input."""
import sys
from pathlib import Path
import copy
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import pixeltable_yolox_b200 as yx
from oracle import yolox_oracle as yo
from pixeltable_yolox_b200 import synthetic as syn

dev = torch.device("cuda", 0)
cfg = yx.YoloxConfig.get_named_config("yolox_s"); cfg.model = None
model = cfg.get_model()
x = torch.from_numpy(syn.images(2, 320, 320, seed=11))
sd = yo.seeded_state_dict(model.state_dict(), 3, (320, 320), calib_x=x)
model.load_state_dict(sd)
model = model.to(dev).eval()
bb = model.backbone.backbone
dt = torch.float16 if len(sys.argv) < 2 else getattr(torch, sys.argv[1])
cur = x.to(dev)
def rel(a, b): return ((a - b).abs() / b.abs().clamp_min(1.0))
for name in ["stem", "dark2", "dark3", "dark4", "dark5"]:
    blk = getattr(bb, name)
    blk32 = copy.deepcopy(blk).float().train()
    for m in blk32.modules():
        if isinstance(m, torch.nn.BatchNorm2d): m.eval()
    xin = cur if name == "stem" else cur.to(dt).float()
    with torch.no_grad():
        ref = blk32._train_forward(xin) if hasattr(blk32, "_train_forward") else torch.nn.Sequential(*blk32)[0]._train_forward(xin)
        if not hasattr(blk32, "_train_forward"):
            ref = xin
            for sub in blk32: ref = sub._train_forward(ref)
    blk16 = copy.deepcopy(blk).to(dt).eval()
    with torch.no_grad():
        if name == "stem":
            got = blk16(xin).float()
        else:
            got = xin.to(dt)
            for sub in blk16: got = sub(got)
            got = got.float()
    e = rel(got, ref)
    print(f"{name}: {str(dt)[6:]} out {tuple(got.shape)} med {e.median().item():.2e} p99 {e.flatten().kthvalue(int(e.numel()*0.99)).values.item():.2e} max {e.max().item():.2e}", flush=True)
    cur = ref

# ---- whole neck and head ----
import copy as _c
neck16 = _c.deepcopy(model.backbone).to(dt).eval()
sd32 = {k: v.to(dev) for k, v in sd.items()}
a = yo.ACTS["silu"]
with torch.no_grad():
    ref_feats = yo.pafpn(sd32, x.to(dev), a)
    got_feats = neck16(x.to(dev))
for i, (g, r) in enumerate(zip(got_feats, ref_feats)):
    e = rel(g.float(), r)
    print(f"pafpn out{i}: med {e.median().item():.2e} p99 {e.flatten().kthvalue(int(e.numel()*0.99)).values.item():.2e}", flush=True)
from pixeltable_yolox_b200.engine import run_head
head16 = _c.deepcopy(model.head).to(dt).eval()
with torch.no_grad():
    ref_out, _ = yo.head(sd32, ref_feats, a)
    got_out = run_head(head16, [r.to(dt) for r in ref_feats]).float()
e = rel(got_out, ref_out)
print(f"head (fp32 feats rounded to {str(dt)[6:]}): med {e.median().item():.2e} p99 {e.flatten().kthvalue(int(e.numel()*0.99)).values.item():.2e}")
for c0, c1, nm in ((0, 2, "xy"), (2, 4, "wh"), (4, 5, "obj"), (5, 85, "cls")):
    ee = e[..., c0:c1]
    print(f"   {nm}: med {ee.median().item():.2e} p99 {ee.flatten().kthvalue(int(ee.numel()*0.99)).values.item():.2e}")
