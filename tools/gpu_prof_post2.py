"""Timing of postprocess on the bench's own prediction tensor at several thresholds (diagnostic)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from pixeltable_yolox_b200 import ops  # noqa: E402
from pixeltable_yolox_b200 import synthetic as syn  # noqa: E402

dev = torch.device("cuda", 0)
args = bench.parse()
cfg, model = bench.build_model(args, dev)
model = model.to(torch.bfloat16).eval()
x = torch.from_numpy(syn.images(64, 640, 640, seed=7)).to(dev)
pred = model(x).float().contiguous()
for thr in (0.3, 0.5, 0.7, 0.9, 0.99, 0.9999):
    for variant in (3,):
        for _ in range(3):
            d, i, c = ops.postprocess_device(pred.clone(), 80, thr, 0.65, variant, max_det=1000)
        torch.cuda.synchronize()
        xx = pred.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        d, i, c = ops.postprocess_device(xx, 80, thr, 0.65, variant, max_det=1000)
        e1.record(); torch.cuda.synchronize()
        sc = pred[..., 4] * pred[..., 5:].max(-1).values
        print(f"thr {thr}: {e0.elapsed_time(e1)*1e3:.0f} us, kept/img mean {c.float().mean().item():.1f} max {c.max().item()}, cand/img mean {(sc >= thr).sum(1).float().mean().item():.1f} max {(sc >= thr).sum(1).max().item()}", flush=True)
