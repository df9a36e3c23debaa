"""GPU diagnostic: PAFPN outputs, ours vs torch-native in the same 16-bit dtype, both against fp32. usage: [dtype]"""
import copy
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import pixeltable_yolox_b200 as yx  # noqa: E402
from oracle import yolox_oracle as yo  # noqa: E402
from pixeltable_yolox_b200 import synthetic as syn  # noqa: E402

dev = torch.device("cuda", 0)
dt = torch.float16 if len(sys.argv) < 2 else getattr(torch, sys.argv[1])
S = int(sys.argv[2]) if len(sys.argv) > 2 else 320
cfg = yx.YoloxConfig.get_named_config("yolox_s"); cfg.model = None
model = cfg.get_model()
x = torch.from_numpy(syn.images(2, S, S, seed=11))
sd = yo.seeded_state_dict(model.state_dict(), 3, (S, S), calib_x=x)
model.load_state_dict(sd)
model = model.to(dev).eval()
a = yo.ACTS["silu"]


def rel(p, q):
    return (p - q).abs() / q.abs().clamp_min(1.0)


def st(e):
    return f"med {e.median().item():.2e} p99 {e.flatten().kthvalue(int(e.numel() * 0.99)).values.item():.2e}"


sd32 = {k: v.to(dev) for k, v in sd.items()}
sd16 = {k: (v.to(dev).to(dt) if v.is_floating_point() else v.to(dev)) for k, v in sd.items()}
with torch.no_grad():
    ref = yo.pafpn(sd32, x.to(dev), a)
    nat = yo.pafpn(sd16, x.to(dev).to(dt), a)
    ours = copy.deepcopy(model.backbone).to(dt).eval()(x.to(dev))
    # the neck alone on identical (fp32-exact, rounded) backbone features
    feats32 = yo.darknet(sd32, x.to(dev), a) if hasattr(yo, "darknet") else None
for i, (o, n, r) in enumerate(zip(ours, nat, ref)):
    print(f"pafpn out{i} {str(dt)[6:]}: ours {st(rel(o.float(), r))} | torch-native {st(rel(n.float(), r))} | ours vs native {st(rel(o.float(), n.float()))}")

with torch.no_grad():
    ref_o, _ = yo.head(sd32, ref, a)
    nat_o, _ = yo.head(sd16, nat, a)
    ours_o = copy.deepcopy(model).to(dt).eval()(x.to(dev))
for c0, c1, nm in ((0, 2, "xy"), (2, 4, "wh"), (4, 5, "obj"), (5, 85, "cls"), (0, 85, "all")):
    print(f"output {nm:3s}: ours {st(rel(ours_o.float()[..., c0:c1], ref_o[..., c0:c1]))} | torch-native {st(rel(nat_o.float()[..., c0:c1], ref_o[..., c0:c1]))}")
