"""One wgrad layer in a loop (for `ncu --set full -k regex:wgrad_tc`). usage: python tools/gpu_prof_wgrad.py [B ci co H k s]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from pixeltable_yolox_b200 import ops  # noqa: E402

B, ci, co, H, k, s = (int(v) for v in (sys.argv[1:7] if len(sys.argv) >= 7 else (8, 128, 128, 80, 3, 1)))
dev = torch.device("cuda", 0)
pad = (k - 1) // 2
OH = (H + 2 * pad - k) // s + 1
x = torch.randn(B, ci, H, H, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
dy = torch.randn(B, co, OH, OH, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
w = torch.empty(co, ci, k, k, device=dev).contiguous(memory_format=torch.channels_last)
for _ in range(5):
    dw = ops.conv_wgrad(x, dy, w, k, s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.conv_wgrad(x, dy, w, k, s)
e1.record()
torch.cuda.synchronize()
flops = 2.0 * B * OH * OH * ci * co * k * k
print(f"wgrad {ci}->{co} k{k} s{s} @{H} batch {B}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call (wgrad + reduce, eager launches), {flops / 1e9:.2f} GFLOP")
