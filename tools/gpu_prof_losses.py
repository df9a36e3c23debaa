"""Timing of the training-side kernels on config-4-like inputs: SimOTA assignment + fused head losses vs the same
losses through torch ops (boolean-mask gathers, one-hot, BCE, autograd backward). usage: gpu_prof_losses.py [B]"""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from oracle.simota_oracle import anchor_grid  # noqa: E402
from pixeltable_yolox_b200 import ops  # noqa: E402
from pixeltable_yolox_b200 import synthetic as syn  # noqa: E402
from pixeltable_yolox_b200.losses import IouLoss  # noqa: E402

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
hw = [(80, 80), (40, 40), (20, 20)]
lab_np = syn.labels(B, max_gt=120, seed=3, size=640.0)
pred_np = syn.train_head_output(B, hw, (8, 16, 32), lab_np, seed=4)
xs, ys, st = (torch.from_numpy(a).to(dev) for a in anchor_grid(hw, (8, 16, 32)))
pred = torch.from_numpy(pred_np).to(dev)
lab = torch.from_numpy(lab_np).to(dev)
A, nc = pred.shape[1], pred.shape[2] - 5


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


def device_time(fn, n=20):
    """Device time of one call: captured in a CUDA graph once, replayed n times (no Python / allocator time)."""
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return timed(g.replay, n)


asg = ops.simota_assign(pred, lab, xs, ys, st, nc, levels=3)
t_asg = timed(lambda: ops.simota_assign(pred, lab, xs, ys, st, nc, levels=3))
t_asg_dev = device_time(lambda: ops.simota_assign(pred, lab, xs, ys, st, nc, levels=3))
print(f"simota_assign: {t_asg_dev:.1f} us on the device (graph replay), {t_asg:.1f} us per Python call; "
      f"{pred.numel() * 4 / t_asg_dev / 1e3:.0f} GB/s of the prediction tensor")
t_loss = timed(lambda: ops.head_losses(pred, lab, asg))
nbytes = 2 * pred.numel() * 4
print(f"B={B} A={A}: simota_assign {t_asg:.1f} us, head_losses {t_loss:.1f} us = {nbytes / t_loss / 1e3:.0f} GB/s "
      f"({pred.numel() * 4 / 1e6:.1f} MB read + the same written)")

iou_loss = IouLoss(reduction="none")


def torch_losses():
    p = pred.clone().requires_grad_(True)
    fg = asg["fg_mask"].bool()
    b_idx, a_idx = fg.nonzero(as_tuple=True)
    g_idx = asg["matched_gt"][b_idx, a_idx].long()
    reg_t = lab[b_idx, g_idx, 1:5]
    cls_t = F.one_hot(asg["matched_cls"][b_idx, a_idx].long(), nc).float() * asg["matched_iou"][b_idx, a_idx].unsqueeze(-1)
    num_fg = asg["num_fg"].sum().clamp(min=1).float()
    l_iou = iou_loss(p[..., :4].reshape(-1, 4)[fg.reshape(-1)], reg_t).sum() / num_fg
    l_obj = F.binary_cross_entropy_with_logits(p[..., 4].reshape(-1), fg.reshape(-1).float(), reduction="none").sum() / num_fg
    l_cls = F.binary_cross_entropy_with_logits(p[..., 5:].reshape(-1, nc)[fg.reshape(-1)], cls_t, reduction="none").sum() / num_fg
    (5 * l_iou + l_obj + l_cls).backward()
    return p.grad


t_torch = timed(torch_losses)
print(f"the same losses + backward through torch ops (what the reference runs after its per-image loop): {t_torch:.1f} us")
