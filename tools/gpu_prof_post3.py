import os, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench
from pixeltable_yolox_b200 import ops, synthetic as syn
dev = torch.device("cuda", 0)
args = bench.parse()
cfg, model = bench.build_model(args, dev)
model = model.to(torch.bfloat16).eval()
x = torch.from_numpy(syn.images(64, 640, 640, seed=7)).to(dev)
pred = model(x).float().contiguous()
ops.postprocess_device(pred.clone(), 80, 0.5, 0.65, 3, max_det=1000)
torch.cuda.synchronize()
os.environ["YX_NMS_DEBUG"] = "1"
ops.postprocess_device(pred.clone(), 80, 0.5, 0.65, 3, max_det=1000)
torch.cuda.synchronize()
