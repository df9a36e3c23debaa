"""Per-layer times of the training conv stack (yolox_s, 8 images, bf16): our forward / dgrad / wgrad launches against
torch's cuDNN calls on the same channels_last tensors. usage: python tools/gpu_prof_trainconv.py [model] [batch]"""
import sys
from collections import OrderedDict
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import pixeltable_yolox_b200 as yx  # noqa: E402
from pixeltable_yolox_b200 import ops, train_conv  # noqa: E402
from pixeltable_yolox_b200._lib import YX_ACT_NONE  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "yolox_s"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda", 0)
model = yx.YoloxConfig.get_named_config(name).get_model().to(dev).train()
shapes = OrderedDict()


def hook(m, inp, out):
    if m.groups != 1:
        return
    key = (inp[0].shape[1], out.shape[1], inp[0].shape[2], inp[0].shape[3], m.kernel_size[0], m.stride[0])
    shapes[key] = shapes.get(key, 0) + 1


for m in model.modules():
    if isinstance(m, torch.nn.Conv2d):
        m.register_forward_hook(hook)
import os
os.environ["YX_TRAIN_CONV"] = "0"
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    x = torch.rand(B, 3, 640, 640, device=dev) * 255
    lab = torch.zeros(B, 120, 5, device=dev)
    lab[:, 0] = torch.tensor([1.0, 320, 320, 100, 100], device=dev)
    model(x, lab)
os.environ["YX_TRAIN_CONV"] = "1"


def timeit(fn, reps=20):
    """GPU time per call: `reps` calls captured in one CUDA graph (no host launch overhead between them), replayed."""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


cl = torch.channels_last
tot = [0.0] * 6
print(f"{'ci':>5} {'co':>5} {'HxW':>9} k s  n | ours fwd  dgrad  wgrad | cudnn fwd  dgrad  wgrad   (us per launch)")
for (ci, co, H, W, k, s), cnt in shapes.items():
    cip, cop = (ci + 15) // 16 * 16, (co + 15) // 16 * 16
    pad = (k - 1) // 2
    OH, OW = (H + 2 * pad - k) // s + 1, (W + 2 * pad - k) // s + 1
    xx = torch.randn(B, cip, H, W, device=dev).bfloat16().contiguous(memory_format=cl)
    dy = torch.randn(B, cop, OH, OW, device=dev).bfloat16().contiguous(memory_format=cl)
    w = torch.randn(cop, cip, k, k, device=dev).contiguous(memory_format=cl) * 0.05
    wf, wd = ops.pack_train_weights(w, torch.bfloat16, cop, cip, True)
    y = torch.empty_like(dy)
    dx = torch.empty_like(xx)
    zb_o, zb_i = torch.zeros(cop, device=dev), torch.zeros(cip, device=dev)
    t_f = timeit(lambda: ops.conv_bn_act(ops._nhwc(xx), wf, zb_o, ops._nhwc(y), k, s, YX_ACT_NONE))
    if s == 1:
        t_d = timeit(lambda: ops.conv_bn_act(ops._nhwc(dy), wd, zb_i, ops._nhwc(dx), k, 1, YX_ACT_NONE))
    else:
        t_dil = timeit(lambda: ops.conv_bn_act(ops._nhwc(ops.dilate2(dy, H, W)), wd, zb_i, ops._nhwc(dx), k, 1, YX_ACT_NONE))
        _, wd4 = ops.pack_train_weights(w, torch.bfloat16, cop, cip, True, subpixel=True)
        zb4 = torch.zeros(4 * cip, device=dev)
        t_d = timeit(lambda: ops.conv_bn_act(ops._nhwc(dy), wd4, zb4, ops._nhwc(dx), 3, 1, YX_ACT_NONE, shuffle2_c=cip))
        print(f"      stride-2 dgrad: zero-stuffed {t_dil:.1f} us, sub-pixel conv {t_d:.1f} us")
    t_w = timeit(lambda: ops.conv_wgrad(xx, dy, w, k, s))
    wb = w.bfloat16()
    c_f = timeit(lambda: F.conv2d(xx, wb, None, s, pad))
    c_d = timeit(lambda: torch.ops.aten.convolution_backward(dy, xx, wb, None, (s, s), (pad, pad), (1, 1), False, (0, 0), 1, (True, False, False)))
    c_w = timeit(lambda: torch.ops.aten.convolution_backward(dy, xx, wb, None, (s, s), (pad, pad), (1, 1), False, (0, 0), 1, (False, True, False)))
    for i, t in enumerate((t_f, t_d, t_w, c_f, c_d, c_w)):
        tot[i] += t * cnt
    print(f"{ci:5d} {co:5d} {H:4d}x{W:<4d} {k} {s} {cnt:2d} | {t_f:8.1f} {t_d:6.1f} {t_w:6.1f} | {c_f:9.1f} {c_d:6.1f} {c_w:6.1f}")
print(f"sum over the network (us): ours fwd {tot[0]:.0f} dgrad {tot[1]:.0f} wgrad {tot[2]:.0f} | cudnn fwd {tot[3]:.0f} dgrad {tot[4]:.0f} wgrad {tot[5]:.0f}")
