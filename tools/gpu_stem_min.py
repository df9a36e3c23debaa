import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from pixeltable_yolox_b200.engine import Builder
from pixeltable_yolox_b200.network_blocks import Focus
dev = torch.device("cuda", 0)
dt = torch.bfloat16
blk = Focus(3, 32, ksize=3).eval().to(dev).to(dt)
b = Builder(dev, dt, use_plan=False)
img = torch.randint(0, 256, (1, 3, 64, 64)).float().to(dev)
out = blk.lower_image(b, img)
torch.cuda.synchronize()
print("ok", out.t.float().abs().mean().item())
