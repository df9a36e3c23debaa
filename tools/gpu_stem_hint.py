"""Stem kernel timing per mbarrier.try_wait suspend hint (YX_MBAR_HINT, ns). usage: gpu_stem_hint.py ns [ns ...]"""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from pixeltable_yolox_b200 import synthetic as syn  # noqa: E402

dev = torch.device("cuda", 0)
hints = sys.argv[1:] or ["20000", "2000", "500", "100", "0"]
sys.argv = sys.argv[:1]
args = bench.parse()
cfg, model = bench.build_model(args, dev)
model = model.to(torch.bfloat16).eval()
x = torch.from_numpy(syn.images(64, 640, 640, seed=7)).to(dev)
for h in hints:
    os.environ["YX_MBAR_HINT"] = h
    model.invalidate_engine()
    eng = model.engine_for(x)
    eng.forward(x)
    prof = eng.builder.profile(); prof = eng.builder.profile()
    print(f"hint={h:>6s} ns  stem {prof[0]['ms'] * 1e3:7.1f} us")
