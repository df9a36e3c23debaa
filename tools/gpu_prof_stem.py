import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench
from pixeltable_yolox_b200 import synthetic as syn
dev = torch.device("cuda", 0)
args = bench.parse()
cfg, model = bench.build_model(args, dev)
model = model.to(torch.bfloat16).eval()
x = torch.from_numpy(syn.images(64, 640, 640, seed=7)).to(dev)
if len(sys.argv) > 1 and sys.argv[1] == "u8":
    x = x.to(torch.uint8)
for _ in range(3):
    model(x)
torch.cuda.synchronize()
print("ok")
