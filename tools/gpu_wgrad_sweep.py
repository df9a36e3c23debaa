"""Sweep the wgrad work split (slices per CTA, pixel split) on representative layers: us per call (wgrad + reduce),
CUDA-graph timed. usage: python tools/gpu_wgrad_sweep.py"""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from pixeltable_yolox_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
LAYERS = [(128, 128, 80, 3, 1), (256, 256, 40, 3, 1), (64, 64, 160, 3, 1), (128, 128, 40, 3, 1)] if B >= 32 else [(256, 256, 20, 3, 1), (512, 256, 20, 1, 1), (128, 128, 40, 3, 1), (256, 128, 40, 1, 1), (128, 128, 80, 3, 1),
          (64, 64, 80, 3, 1), (32, 32, 160, 3, 1), (16, 32, 320, 3, 1), (32, 64, 320, 3, 2), (128, 256, 40, 3, 2)]


def timeit(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for ci, co, H, k, s in LAYERS:
    pad = (k - 1) // 2
    OH = (H + 2 * pad - k) // s + 1
    x = torch.randn(B, ci, H, H, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    dy = torch.randn(B, co, OH, OH, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    w = torch.empty(co, ci, k, k, device=dev).contiguous(memory_format=torch.channels_last)
    row = []
    for sg in (0, 1, 2, 4):
        for ks in ((0, 8, 16, 32, 64, 148) if B >= 32 else (0, 1, 2, 4, 8, 16, 32, 64)):
            for kp in (0, 32):
                for name, v in (("YX_WGRAD_SG", sg), ("YX_WGRAD_KSPLIT", ks), ("YX_WGRAD_KP", kp)):
                    if v:
                        os.environ[name] = str(v)
                    else:
                        os.environ.pop(name, None)
                try:
                    row.append((timeit(lambda: ops.conv_wgrad(x, dy, w, k, s)), sg, ks, kp))
                except Exception as e:      # noqa: BLE001
                    row.append((1e9, sg, ks, kp))
    default = [r for r in row if r[1:] == (0, 0, 0)][0][0]
    row.sort()
    print(f"{ci:4d}->{co:4d} @{H:3d} k{k} s{s}: default {default:6.1f} us | best " +
          "  ".join(f"{t:6.1f} (sg {sg} ks {ks} kp {kp})" for t, sg, ks, kp in row[:5]), flush=True)
