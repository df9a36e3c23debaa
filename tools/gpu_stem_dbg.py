"""Stem kernel timing under YX_STEM_DEBUG experiment bits. usage: gpu_stem_dbg.py bits [bits ...]"""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from pixeltable_yolox_b200 import synthetic as syn  # noqa: E402

dev = torch.device("cuda", 0)
bits = sys.argv[1:] or ["0", "4", "8", "12"]
sys.argv = sys.argv[:1]
args = bench.parse()
cfg, model = bench.build_model(args, dev)
model = model.to(torch.bfloat16).eval()
x = torch.from_numpy(syn.images(64, 640, 640, seed=7)).to(dev)
for bt in bits:
    os.environ["YX_STEM_DEBUG"] = bt
    model.invalidate_engine()
    eng = model.engine_for(x)
    eng.forward(x)
    prof = eng.builder.profile(); prof = eng.builder.profile()
    print(f"debug={bt:>3s} stem {prof[0]['ms'] * 1e3:7.1f} us")
