"""numpy/C restatement (test infrastructure) of the SimOTA label assignment
(YoloxHead.get_assignments / get_geometry_constraint / simota_matching,
yolox/models/yolo_head.py:420-574)."""
from __future__ import annotations

import numpy as np

from ._native import lib
from .postprocess_oracle import bboxes_iou

f32 = np.float32


def simota_matching(cost: np.ndarray, ious: np.ndarray):
    """yolo_head.py:542-574 on [G, n] matrices. Returns (match_gt [n] int32 with -1 = unmatched,
    match_iou [n] fp32, num_fg)."""
    cost = np.ascontiguousarray(cost, dtype=f32)
    ious = np.ascontiguousarray(ious, dtype=f32)
    G, n = cost.shape
    mg = np.empty((n,), dtype=np.int32)
    mi = np.empty((n,), dtype=f32)
    nf = lib().oracle_simota_matching(cost.ctypes.data, ious.ctypes.data, G, n, mg.ctypes.data, mi.ctypes.data)
    return mg, mi, int(nf)


def geometry_constraint(gt_boxes, strides, x_shifts, y_shifts):
    """yolo_head.py:511-540. gt_boxes [G,4] cxcywh; per-anchor arrays [A]."""
    xc = ((x_shifts + f32(0.5)) * strides)[None, :]
    yc = ((y_shifts + f32(0.5)) * strides)[None, :]
    dist = (strides * f32(1.5))[None, :]
    l = gt_boxes[:, 0:1] - dist; r = gt_boxes[:, 0:1] + dist
    t = gt_boxes[:, 1:2] - dist; b = gt_boxes[:, 1:2] + dist
    deltas = np.stack([xc - l, yc - t, r - xc, b - yc], 2)
    is_in = deltas.min(axis=-1) > 0
    anchor_filter = is_in.sum(axis=0) > 0
    return anchor_filter, is_in[:, anchor_filter]


def _sigmoid(x):
    return (f32(1) / (f32(1) + np.exp(-x.astype(f32)))).astype(f32)


def cost_matrices(pred, gt, num_classes, strides, x_shifts, y_shifts):
    """pred [A, 5+nc] (decoded boxes, raw logits), gt [G,5] (cls,cx,cy,w,h) -> fg_mask, cost, ious."""
    pred = pred.astype(f32); gt = gt.astype(f32)
    fg_mask, geom = geometry_constraint(gt[:, 1:5], strides.astype(f32), x_shifts.astype(f32), y_shifts.astype(f32))
    p = pred[fg_mask]
    ious = bboxes_iou(gt[:, 1:5], p[:, :4], xyxy=False)
    iou_loss = -np.log(ious + f32(1e-8))
    prob = np.sqrt(_sigmoid(p[:, 5:]) * _sigmoid(p[:, 4:5])).astype(f32)        # [A', nc]
    onehot = np.eye(num_classes, dtype=f32)[gt[:, 0].astype(np.int64)]           # [G, nc]
    with np.errstate(divide="ignore"):
        logp = np.maximum(np.log(prob), f32(-100.0))
        log1mp = np.maximum(np.log(f32(1) - prob), f32(-100.0))
    # F.binary_cross_entropy(..., reduction="none").sum(-1)
    cls_cost = -(onehot[:, None, :] * logp[None] + (f32(1) - onehot[:, None, :]) * log1mp[None]).sum(-1)
    cost = cls_cost.astype(f32) + f32(3.0) * iou_loss + f32(1e6) * (~geom).astype(f32)
    return fg_mask, cost.astype(f32), ious.astype(f32)


def get_assignments(pred, gt, num_classes, strides, x_shifts, y_shifts):
    """Dense per-anchor result: fg [A] bool, matched_gt [A] int32 (-1), matched_iou [A], num_fg."""
    A = pred.shape[0]
    fg = np.zeros((A,), dtype=bool)
    mgt = np.full((A,), -1, dtype=np.int32)
    miou = np.zeros((A,), dtype=f32)
    if gt.shape[0] == 0:
        return fg, mgt, miou, 0
    fg_mask, cost, ious = cost_matrices(pred, gt, num_classes, strides, x_shifts, y_shifts)
    if cost.shape[1] == 0:
        return fg, mgt, miou, 0
    mg, mi, nf = simota_matching(cost, ious)
    cand = np.where(fg_mask)[0]
    sel = mg >= 0
    fg[cand[sel]] = True
    mgt[cand[sel]] = mg[sel]
    miou[cand[sel]] = mi[sel]
    return fg, mgt, miou, nf


def anchor_grid(hw, strides):
    """x_shifts, y_shifts, expanded_strides per anchor (yolo_head.py:213-231), level-major."""
    xs, ys, st = [], [], []
    for (h, w), s in zip(hw, strides):
        yv, xv = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
        xs.append(xv.reshape(-1).astype(f32)); ys.append(yv.reshape(-1).astype(f32))
        st.append(np.full((h * w,), s, dtype=f32))
    return np.concatenate(xs), np.concatenate(ys), np.concatenate(st)
