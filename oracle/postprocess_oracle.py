"""numpy/C restatement (test infrastructure) of utils.boxes.postprocess and bboxes_iou
(yolox/utils/boxes.py:31-101) and of torchvision.ops.batched_nms as called at boxes.py:62-67."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np

from ._native import lib

f32 = np.float32


def nms(boxes: np.ndarray, scores: np.ndarray, thr: float) -> np.ndarray:
    """torchvision.ops.nms on CPU: indices of kept boxes in descending-score order."""
    boxes = np.ascontiguousarray(boxes, dtype=f32)
    scores = np.ascontiguousarray(scores, dtype=f32)
    n = boxes.shape[0]
    keep = np.empty((max(n, 1),), dtype=np.int64)
    k = lib().oracle_nms(boxes.ctypes.data, scores.ctypes.data, n, float(thr), keep.ctypes.data)
    return keep[:k].copy()


def batched_nms_offset(boxes, scores, idxs, thr):
    """torchvision _batched_nms_coordinate_trick: boxes + idxs.to(boxes) * (boxes.max() + 1), one NMS."""
    if boxes.shape[0] == 0:
        return np.empty((0,), dtype=np.int64)
    boxes = boxes.astype(f32)
    max_coordinate = boxes.max()
    offsets = idxs.astype(f32) * f32(max_coordinate + f32(1))
    return nms(boxes + offsets[:, None], scores, thr)


def batched_nms_per_class(boxes, scores, idxs, thr):
    """torchvision _batched_nms_vanilla: NMS per class, then kept indices sorted by descending score."""
    keep_mask = np.zeros(scores.shape[0], dtype=bool)
    for c in np.unique(idxs):
        cur = np.where(idxs == c)[0]
        keep_mask[cur[nms(boxes[cur], scores[cur], thr)]] = True
    kept = np.where(keep_mask)[0]
    order = np.argsort(-scores[kept].astype(f32), kind="stable")
    return kept[order]


def batched_nms(boxes, scores, idxs, thr, variant="auto_cpu"):
    if variant == "offset":
        return batched_nms_offset(boxes, scores, idxs, thr)
    if variant == "per_class":
        return batched_nms_per_class(boxes, scores, idxs, thr)
    # torchvision/ops/boxes.py batched_nms: 4000 on CPU; on CUDA 100000 (torchvision >= 0.19, 0.26 installed) or
    # 20000 (torchvision 0.17.2, the reference's pin: "auto_tv017")
    limit = {"auto_cpu": 4000, "auto_tv017": 20000}.get(variant, 100000)
    if boxes.size > limit:
        return batched_nms_per_class(boxes, scores, idxs, thr)
    return batched_nms_offset(boxes, scores, idxs, thr)


def postprocess(prediction: np.ndarray, num_classes: int, conf_thre=0.7, nms_thre=0.45, class_agnostic=False,
                variant="auto_cpu", return_indices=False):
    """boxes.py:31-75. prediction [B, A, 5+nc] fp32 is modified in place (cxcywh -> xyxy).
    Returns list of [n,7] arrays (or None); with return_indices also the kept anchor indices."""
    assert prediction.dtype == f32
    cx, cy, w, h = (prediction[:, :, i].copy() for i in range(4))
    prediction[:, :, 0] = cx - w / f32(2)
    prediction[:, :, 1] = cy - h / f32(2)
    prediction[:, :, 2] = cx + w / f32(2)
    prediction[:, :, 3] = cy + h / f32(2)
    outs: List[Optional[np.ndarray]] = []
    idxs_out: List[Optional[np.ndarray]] = []
    thr32 = f32(conf_thre)   # torch compares the fp32 tensor with the scalar cast to fp32
    for image_pred in prediction:
        cls = image_pred[:, 5:5 + num_classes]
        class_pred = np.argmax(cls, axis=1)              # first index of the maximum, like torch.max
        class_conf = cls[np.arange(cls.shape[0]), class_pred]
        score = image_pred[:, 4] * class_conf            # fp32 product
        mask = score >= thr32
        anchors = np.where(mask)[0]
        if anchors.size == 0:
            outs.append(None); idxs_out.append(None)
            continue
        det = np.concatenate([image_pred[anchors, :5], class_conf[anchors, None],
                              class_pred[anchors, None].astype(f32)], axis=1)
        sc = det[:, 4] * det[:, 5]
        if class_agnostic:
            keep = nms(det[:, :4], sc, nms_thre)
        else:
            keep = batched_nms(det[:, :4], sc, det[:, 6].astype(np.int64), nms_thre, variant)
        outs.append(det[keep]); idxs_out.append(anchors[keep])
    return (outs, idxs_out) if return_indices else outs


def bboxes_iou(a: np.ndarray, b: np.ndarray, xyxy=True) -> np.ndarray:
    """boxes.py:78-101."""
    if a.shape[1] != 4 or b.shape[1] != 4:
        raise IndexError
    a = a.astype(f32); b = b.astype(f32)
    if xyxy:
        tl = np.maximum(a[:, None, :2], b[:, :2]); br = np.minimum(a[:, None, 2:], b[:, 2:])
        area_a = np.prod(a[:, 2:] - a[:, :2], 1); area_b = np.prod(b[:, 2:] - b[:, :2], 1)
    else:
        tl = np.maximum(a[:, None, :2] - a[:, None, 2:] / f32(2), b[:, :2] - b[:, 2:] / f32(2))
        br = np.minimum(a[:, None, :2] + a[:, None, 2:] / f32(2), b[:, :2] + b[:, 2:] / f32(2))
        area_a = np.prod(a[:, 2:], 1); area_b = np.prod(b[:, 2:], 1)
    en = (tl < br).astype(f32).prod(axis=2)
    area_i = np.prod(br - tl, 2) * en
    with np.errstate(divide="ignore", invalid="ignore"):
        return (area_i / (area_a[:, None] + area_b - area_i)).astype(f32)
