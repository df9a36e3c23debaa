"""YX_STEM_DEBUG=2 trace of the stem kernel (one eager launch). usage: YX_STEM_DEBUG=2 python tools/gpu_stem_trace.py"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from pixeltable_yolox_b200 import synthetic as syn  # noqa: E402

dev = torch.device("cuda", 0)
sys.argv = sys.argv[:1]
args = bench.parse()
cfg, model = bench.build_model(args, dev)
model = model.to(torch.bfloat16).eval()
model.use_cuda_graph = False
x = torch.from_numpy(syn.images(64, 640, 640, seed=7)).to(dev)
model(x)
torch.cuda.synchronize()
print("ok")
