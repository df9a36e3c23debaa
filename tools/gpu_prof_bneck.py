import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from pixeltable_yolox_b200 import ops
from pixeltable_yolox_b200.ops import View
dev = torch.device("cuda", 0)
B, hw, c = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 80, int(sys.argv[2]) if len(sys.argv) > 2 else 64
dt = torch.float16 if len(sys.argv) > 3 and sys.argv[3] == "fp16" else torch.bfloat16
x = torch.randn(B, hw, hw, c, device=dev).to(dt)
w1 = (torch.randn(c, 1, c, device=dev) / c ** 0.5).to(dt)
w2 = (torch.randn(c, 9, c, device=dev) / (9 * c) ** 0.5).to(dt)
b1 = torch.zeros(c, device=dev); b2 = torch.zeros(c, device=dev)
o = torch.empty_like(x)
for _ in range(3):
    ops.bottleneck_fwd(View(x), w1, b1, w2, b2, View(o), 1, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.bottleneck_fwd(View(x), w1, b1, w2, b2, View(o), 1, True)
e1.record(); torch.cuda.synchronize()
print(f"bneck {c}@{hw} {dt}: {e0.elapsed_time(e1)*1e3/5:.1f} us")
