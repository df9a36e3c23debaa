"""On-GPU diagnostic (not pytest): model-level error statistics of the fp32 / bf16 / fp16 paths
against the reference goldens, next to what plain torch bf16/fp16 arithmetic achieves."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import cases  # noqa: E402
import pixeltable_yolox_b200 as yx  # noqa: E402
from oracle import yolox_oracle as yo  # noqa: E402
from pixeltable_yolox_b200 import synthetic as syn  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda", 0)
g = np.load(ROOT / "tests/golden/network.npz")
out_lines = []


def log(*a):
    s = " ".join(str(x) for x in a); print(s, flush=True); out_lines.append(s)


def stats(tag, out, ref):
    rel = np.abs(out - ref) / np.maximum(np.abs(ref), 1.0)
    p = np.abs(out[..., 4:] - ref[..., 4:])
    b = rel[..., :4]
    log(f"  {tag:28s} all max {rel.max():.2e} | box rel med {np.median(b):.2e} p99 {np.quantile(b, .99):.2e} max {b.max():.2e}"
        f" | prob abs med {np.median(p):.2e} p99 {np.quantile(p, .99):.2e} max {p.max():.2e}")


for name, c in cases.NET_CASES.items():
    cfg = yx.YoloxConfig(name, depth=c["depth"], width=c["width"], depthwise=c["depthwise"])
    model = cfg.get_model()
    x = torch.from_numpy(syn.images(c["batch"], c["h"], c["w"], seed=c["seed"] + 500))
    sd = yo.seeded_state_dict(model.state_dict(), c["seed"], (c["h"], c["w"]), calib_x=x)
    model.load_state_dict(sd)
    ref = g[f"{name}/out"]; refu = g[f"{name}/undecoded"]
    log(name, "ref box absmax", np.abs(ref[..., :4]).max())
    m = model.to(dev).eval()
    stats("ours fp32", m(x.to(dev)).cpu().numpy(), ref)
    m.head.decode_in_inference = False; m.invalidate_engine()
    stats("ours fp32 undecoded", m(x.to(dev)).cpu().numpy(), refu)
    m.head.decode_in_inference = True; m.invalidate_engine()
    sdc = {k: v.to(dev) for k, v in sd.items()}
    stats("torch-cuda fp32 oracle", yo.forward(sdc, x.to(dev)).cpu().numpy(), ref)
    for dt in (torch.bfloat16, torch.float16):
        mm = m.to(dt)
        stats(f"ours {str(dt)[6:]}", mm(x.to(dev)).float().cpu().numpy(), ref)
        # plain torch arithmetic in the same dtype (what the reference's .half()/.bfloat16() does)
        sdd = {k: (v.to(dt) if v.is_floating_point() else v) for k, v in sdc.items()}
        a = yo.ACTS["silu"]
        with torch.no_grad():
            o, _ = yo.head(sdd, yo.pafpn(sdd, x.to(dev).to(dt), a), a)
        stats(f"torch {str(dt)[6:]}", o.float().cpu().numpy(), ref)
    m.float()
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "model_diag.txt").write_text("\n".join(out_lines))
