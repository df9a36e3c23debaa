"""Stem kernel timing per tile width (YX_STEM_TILE experiment switch). usage: gpu_stem_tiles.py [tw ...]"""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from pixeltable_yolox_b200 import synthetic as syn  # noqa: E402

dev = torch.device("cuda", 0)
tiles = sys.argv[1:] or ["0", "16", "24", "32", "40"]
sys.argv = sys.argv[:1]
args = bench.parse()
cfg, model = bench.build_model(args, dev)
model = model.to(torch.bfloat16).eval()
x = torch.from_numpy(syn.images(64, 640, 640, seed=7)).to(dev)
ref = None
for tw in tiles:
    os.environ["YX_STEM_TILE"] = tw
    model.invalidate_engine()
    for name, xx in (("fp32", x), ("u8", x.to(torch.uint8))):
        eng = model.engine_for(xx)
        out = eng.forward(xx).clone()
        prof = eng.builder.profile(); prof = eng.builder.profile()
        if ref is None:
            ref = out
        print(f"tw={tw:>3s} {name:4s} stem {prof[0]['ms'] * 1e3:7.1f} us  same output: {torch.equal(out, ref)}")
