"""Tiny driver for ncu: selected conv shapes of yolox_s. usage: gpu_prof_conv.py B shape[,shape...]
shape = cin:cout:k:stride:hw"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from pixeltable_yolox_b200 import ops  # noqa: E402
from pixeltable_yolox_b200.ops import View  # noqa: E402

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
shapes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["128:128:1:1:80", "128:128:3:1:80"]
for sh in shapes:
    cin, cout, k, s, hw = (int(v) for v in sh.split(":"))
    oh = (hw + 2 * ((k - 1) // 2) - k) // s + 1
    x = torch.randn(B, hw, hw, cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(cout, k * k, cin, device=dev) / (k * k * cin) ** 0.5).to(torch.bfloat16)
    bias = torch.zeros(cout, device=dev)
    o = torch.empty(B, oh, oh, cout, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        ops.conv_bn_act(View(x), w, bias, View(o), k, s, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.conv_bn_act(View(x), w, bias, View(o), k, s, 1)
    e1.record(); torch.cuda.synchronize()
    print(sh, f"{e0.elapsed_time(e1)*1e3:.1f} us")
print("ok")
