"""Per-phase cycle counts of sort_nms_kernel (YX_NMS_DEBUG) for one dense image at three thresholds and for the
every-anchor-kept scene. usage: python tools/gpu_nms_phases.py"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from pixeltable_yolox_b200 import ops, synthetic as syn  # noqa: E402
from pixeltable_yolox_b200.boxes import NMS_VARIANTS  # noqa: E402

dev = torch.device("cuda", 0)
A = 8400
dense = torch.from_numpy(syn.dense_scene(2, anchors=A, seed=13)).to(dev)
grid = np.zeros((2, A, 85), dtype=np.float32)
gx, gy = np.meshgrid(np.arange(100), np.arange(84))
grid[:, :, 0] = (gx.reshape(-1) * 6.0 + 3.0)[None]; grid[:, :, 1] = (gy.reshape(-1) * 6.0 + 3.0)[None]
grid[:, :, 2:4] = 5.0
rng = np.random.default_rng(10)
grid[:, :, 4] = rng.uniform(0.5, 1.0, (2, A))
np.put_along_axis(grid[:, :, 5:], rng.integers(0, 3, (2, A))[..., None], 0.9, axis=2)       # three classes only
grid = torch.from_numpy(grid).to(dev)
for name, pred, thr in (("dense", dense, 0.001), ("dense", dense, 0.25), ("dense", dense, 0.5), ("all kept, 3 classes", grid, 0.01)):
    ops.postprocess_device(pred.clone(), 80, thr, 0.65, NMS_VARIANTS["auto"])
    torch.cuda.synchronize()
    print(f"== {name} thr {thr}", flush=True)
    os.environ["YX_NMS_DEBUG"] = "1"
    ops.postprocess_device(pred.clone(), 80, thr, 0.65, NMS_VARIANTS["auto"])
    torch.cuda.synchronize()
    del os.environ["YX_NMS_DEBUG"]
