"""TEST INFRASTRUCTURE (checker only, never on the product path): numpy restatement of the YOLOX head losses and of
their gradients w.r.t. the prediction tensor.

  reference: YoloxHead.get_losses (yolox/models/yolo_head.py:354-418), IOUloss (yolox/models/losses.py:7-51),
  torch.nn.BCEWithLogitsLoss(reduction="none"), torch.nn.L1Loss(reduction="none"), get_l1_target (:412-418).
Pinned by tests/golden/losses.npz: loss values and autograd gradients of the unmodified reference on the seeded
cases of tests/cases.py (tests/test_oracle_golden.py). float64 throughout.
"""
from __future__ import annotations

import numpy as np


def _bce(x, t):
    """BCEWithLogits value and d/dx."""
    e = np.exp(-np.abs(x))
    sig = np.where(x >= 0, 1.0 / (1.0 + e), e / (1.0 + e))
    return np.maximum(x, 0) - x * t + np.log1p(e), sig - t


def _dmax(a, b):       # d max(a, b) / d a, ties split (torch.maximum backward)
    return np.where(a > b, 1.0, np.where(a == b, 0.5, 0.0))


def _dmin(a, b):
    return np.where(a < b, 1.0, np.where(a == b, 0.5, 0.0))


def iou_loss(pred, target, loss_type="iou"):
    """losses.py:14-48 on [n,4] cxcywh rows. Returns (loss [n], dloss/dpred [n,4])."""
    pred = pred.astype(np.float64); target = target.astype(np.float64)
    pc, pwh, tc, twh = pred[:, :2], pred[:, 2:], target[:, :2], target[:, 2:]
    plo, phi, tlo, thi = pc - pwh / 2, pc + pwh / 2, tc - twh / 2, tc + twh / 2
    tl, br = np.maximum(plo, tlo), np.minimum(phi, thi)
    area_p, area_g = pwh.prod(1), twh.prod(1)
    en = (tl < br).all(1).astype(np.float64)
    wh = br - tl
    area_i = wh.prod(1) * en
    area_u = area_p + area_g - area_i
    U = area_u + 1e-16
    iou = area_i / U
    a_lo, a_hi = _dmax(plo, tlo), _dmin(phi, thi)                  # [n,2] each
    dwh_c = a_hi - a_lo                                            # d (br - tl) / d centre
    dwh_s = 0.5 * (a_hi + a_lo)                                    # d (br - tl) / d size
    other = wh[:, ::-1]
    dI = np.concatenate([dwh_c * other, dwh_s * other], 1) * en[:, None]
    dP = np.concatenate([np.zeros_like(pc), pwh[:, ::-1]], 1)
    dU = dP - dI
    diou = (dI * U[:, None] - area_i[:, None] * dU) / (U * U)[:, None]
    if loss_type == "iou":
        return 1 - iou ** 2, -2 * iou[:, None] * diou
    clo, chi = np.minimum(plo, tlo), np.maximum(phi, thi)
    cwh = chi - clo
    area_c = cwh.prod(1)
    C = np.maximum(area_c, 1e-16)
    g = iou - (area_c - area_u) / C
    c_lo, c_hi = _dmin(plo, tlo), _dmax(phi, thi)
    dC = np.concatenate([(c_hi - c_lo) * cwh[:, ::-1], 0.5 * (c_hi + c_lo) * cwh[:, ::-1]], 1)
    dCk = dC * (area_c >= 1e-16)[:, None]
    dterm = ((dC - dU) * C[:, None] - (area_c - area_u)[:, None] * dCk) / (C * C)[:, None]
    inside = ((g >= -1) & (g <= 1))[:, None]
    return 1 - np.clip(g, -1, 1), np.where(inside, -(diou - dterm), 0.0)


def head_losses(pred, labels, fg_mask, matched_gt, matched_iou, matched_cls, origin=None, x_shift=None, y_shift=None,
                stride=None, loss_type="iou", reg_weight=5.0):
    """pred [B,A,5+nc]; labels [B,G,5] (cls,cx,cy,w,h); dense assignment arrays [B,A].
    Returns (sums [4] = iou/obj/cls/l1 un-normalised, grad [B,A,5+nc] of reg_weight*iou+obj+cls, grad_origin or None)."""
    pred = pred.astype(np.float64)
    B, A, nch = pred.shape
    nc = nch - 5
    fg = fg_mask.astype(bool)
    grad = np.zeros_like(pred)
    l_obj, d_obj = _bce(pred[..., 4], fg.astype(np.float64))
    grad[..., 4] = d_obj
    bi, ai = np.nonzero(fg)
    gt = labels.astype(np.float64)[bi, matched_gt[bi, ai]]
    l_iou, d_iou = iou_loss(pred[bi, ai, :4], gt[:, 1:5], loss_type)
    grad[bi, ai, :4] = reg_weight * d_iou
    tgt = np.zeros((len(bi), nc))
    tgt[np.arange(len(bi)), matched_cls[bi, ai]] = matched_iou[bi, ai].astype(np.float64)
    l_cls, d_cls = _bce(pred[bi, ai, 5:], tgt)
    grad[bi, ai, 5:] = d_cls
    sums = [l_iou.sum(), l_obj.sum(), l_cls.sum(), 0.0]
    g_or = None
    if origin is not None:
        st = stride.astype(np.float64)[ai]
        t = np.stack([gt[:, 1] / st - x_shift.astype(np.float64)[ai], gt[:, 2] / st - y_shift.astype(np.float64)[ai],
                      np.log(gt[:, 3] / st + 1e-8), np.log(gt[:, 4] / st + 1e-8)], 1)
        df = origin.astype(np.float64)[bi, ai] - t
        sums[3] = np.abs(df).sum()
        g_or = np.zeros(origin.shape)
        g_or[bi, ai] = np.sign(df)
    return np.array(sums), grad, g_or
