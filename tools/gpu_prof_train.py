"""Kernel-time breakdown of one training step (yolox_s, 8 images, bf16 autocast) with torch.profiler: which part of the
torch / cuDNN side the step spends its GPU time in. usage: python tools/gpu_prof_train.py"""
import os
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from pixeltable_yolox_b200.optim import FusedSgdEma  # noqa: E402

fmt = sys.argv[1] if len(sys.argv) > 1 else "channels_last"
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
sys.argv = [sys.argv[0], "--train", "--train-format", fmt]
args = bench.parse()
dev = torch.device("cuda", 0)
cfg, model = bench.build_model(args, dev)
model.train()
if fmt == "channels_last":
    model = model.to(memory_format=torch.channels_last)
direct = os.environ.get("YX_TRAIN_CONV", "1") != "0"
opt = FusedSgdEma(model, lr=1e-3, direct_grads=direct)
if direct:
    from pixeltable_yolox_b200 import train_conv
    train_conv.attach_packer(model, torch.bfloat16)
x, lab, _ = bench.train_batch(args, 0, 8)
x, lab = x.to(dev), lab.to(dev)
if fmt == "channels_last":
    x = x.contiguous(memory_format=torch.channels_last)


def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = model(x, lab)
    opt.zero_grad()
    out["total_loss"].backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
rows = prof.key_averages()
tot = sum(r.device_time_total for r in rows)
print(f"total GPU kernel time per step: {tot / 3 / 1e3:.2f} ms")
groups = {}
for r in rows:
    k = r.key
    name = ("cudnn/conv" if any(s in k for s in ("cudnn", "conv", "gemm", "cutlass", "xmma", "sm90", "sm100", "nvjet", "wgrad", "dgrad")) else
            "batch_norm" if "batch_norm" in k or "bn_" in k else
            "elementwise (silu, add, mul, copy, cast)" if any(s in k for s in ("elementwise", "vectorized", "copy", "Copy", "silu", "fill")) else
            "ours (yx::)" if "yx::" in k else
            "cat / index / reduce / other")
    g = groups.setdefault(name, [0.0, 0])
    g[0] += r.device_time_total; g[1] += r.count
for k, (t, n) in sorted(groups.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:44s} {t / 3 / 1e3:7.2f} ms  {100 * t / tot:5.1f}%  {n // 3} launches")
for r in sorted(rows, key=lambda r: -r.device_time_total)[:top]:
    print(f"    {r.key[:90]:90s} {r.device_time_total / 3 / 1e3:7.3f} ms x{r.count // 3}")
