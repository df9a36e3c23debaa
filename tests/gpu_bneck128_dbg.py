import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from pixeltable_yolox_b200 import ops
from pixeltable_yolox_b200.ops import View
dev = torch.device("cuda", 0)
B, hw, c = int(sys.argv[1]), 40, 128
x = torch.randn(B, hw, hw, c, device=dev).to(torch.bfloat16)
w1 = (torch.randn(c, 1, c, device=dev) / c ** 0.5).to(torch.bfloat16)
w2 = (torch.randn(c, 9, c, device=dev) / (9 * c) ** 0.5).to(torch.bfloat16)
b1 = torch.zeros(c, device=dev); b2 = torch.zeros(c, device=dev)
o = torch.empty_like(x)
ops.bottleneck_fwd(View(x), w1, b1, w2, b2, View(o), 1, True)
torch.cuda.synchronize()
print("B", B, "ok", o.float().abs().mean().item())
