"""Tiny driver (ncu / timing) for the postprocess kernels on dense and sparse synthetic scenes."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from pixeltable_yolox_b200 import ops  # noqa: E402
from pixeltable_yolox_b200 import synthetic as syn  # noqa: E402

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
one = syn.dense_scene(4, anchors=8400, seed=13)
dense = torch.from_numpy(np.concatenate([one] * (B // 4), 0)).to(dev)
sp = syn.sparse_scene(4, anchors=8400, seed=5)
sparse = torch.from_numpy(np.concatenate([sp] * (B // 4), 0)).to(dev)
for name, t, thr in (("dense thr0.001", dense, 0.001), ("dense thr0.5", dense, 0.5), ("sparse thr0.5", sparse, 0.5)):
    for variant in (0, 1):
        for _ in range(2):
            d, i, c = ops.postprocess_device(t.clone(), 80, thr, 0.65, variant, max_det=1000)
        torch.cuda.synchronize()
        x = t.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        d, i, c = ops.postprocess_device(x, 80, thr, 0.65, variant, max_det=1000)
        e1.record(); torch.cuda.synchronize()
        print(f"{name} variant {variant}: {e0.elapsed_time(e1)*1e3:.0f} us, kept/img {c.float().mean().item():.0f}", flush=True)
