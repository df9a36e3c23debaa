"""Driver for ncu / timing: the head prediction GEMM (256 -> 96 block-diagonal reg|obj|cls, decode epilogue).
usage: gpu_prof_head.py [B] [hw]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from pixeltable_yolox_b200 import ops  # noqa: E402
from pixeltable_yolox_b200.ops import View  # noqa: E402

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
hw = int(sys.argv[2]) if len(sys.argv) > 2 else 80
nc, A = 80, hw * hw
x = torch.randn(B, hw, hw, 256, device=dev).to(torch.bfloat16)
w = (torch.randn(96, 1, 256, device=dev) / 16).to(torch.bfloat16)
bias = torch.randn(96, device=dev)
out = torch.empty(B, A, 5 + nc, device=dev)
head = {"out_ptr": out.data_ptr(), "anchors": A, "anchor_off": 0, "nc": nc, "decode": 3, "stride": 8.0}
for _ in range(3):
    ops.conv_bn_act(View(x), w, bias, None, 1, 1, 0, head=head)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.conv_bn_act(View(x), w, bias, None, 1, 1, 0, head=head)
e1.record(); torch.cuda.synchronize()
print(f"head pred 256->96 @{hw}: {e0.elapsed_time(e1) * 1e3 / 5:.1f} us")
