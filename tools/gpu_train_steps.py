"""A few eager training steps of the --train bench configuration (yolox_s, 8 images, bf16 autocast, direct gradients, batched
weight packing) with no profiler of its own: the target of `ncu -k regex:... -s <skip> -c <count>` launch lists.
usage: python tools/gpu_train_steps.py [steps]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from pixeltable_yolox_b200 import train_conv  # noqa: E402
from pixeltable_yolox_b200.optim import FusedSgdEma  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
sys.argv = [sys.argv[0], "--train"]
args = bench.parse()
dev = torch.device("cuda", 0)
cfg, model = bench.build_model(args, dev)
model = model.train().to(memory_format=torch.channels_last)
opt = FusedSgdEma(model, lr=1e-3, direct_grads=True)
train_conv.attach_packer(model, torch.bfloat16)
x, lab, _ = bench.train_batch(args, 0, 8)
x, lab = x.to(dev).contiguous(memory_format=torch.channels_last), lab.to(dev)
for i in range(steps):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = model(x, lab)
    opt.zero_grad()
    out["total_loss"].backward()
    opt.step()
    torch.cuda.synchronize()
    print(f"step {i}: loss {float(out['total_loss']):.4f}", flush=True)
