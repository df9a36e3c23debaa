"""Tiny driver for ncu: the two conv shapes that dominate yolox_s (3x3 and 1x1, 128->128 @80x80, B=64)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from pixeltable_yolox_b200 import ops  # noqa: E402
from pixeltable_yolox_b200.ops import View  # noqa: E402

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for (cin, cout, k, hw) in ((128, 128, 1, 80), (128, 128, 3, 80)):
    x = torch.randn(B, hw, hw, cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(cout, k * k, cin, device=dev) / (k * k * cin) ** 0.5).to(torch.bfloat16)
    bias = torch.zeros(cout, device=dev)
    o = torch.empty(B, hw, hw, cout, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        ops.conv_bn_act(View(x), w, bias, View(o), k, 1, 1)
    torch.cuda.synchronize()
print("ok")
