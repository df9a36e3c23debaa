"""Box utilities with the reference's signatures (yolox/utils/boxes.py:31-101), executed by the
sm_100a kernels (csrc/yx_postprocess.cu, csrc/yx_misc.cu). CUDA tensors only."""
from __future__ import annotations

from typing import List, Optional

import torch

from . import _lib, ops

__all__ = ["postprocess", "bboxes_iou", "NMS_VARIANTS"]

NMS_VARIANTS = {
    "offset": _lib.NMS_OFFSET,        # torchvision _batched_nms_coordinate_trick
    "per_class": _lib.NMS_PER_CLASS,  # torchvision _batched_nms_vanilla
    "auto": 3,                        # what the installed torchvision (>= 0.19) picks on CUDA (<= 100000 coordinates -> offset)
    "auto_tv017": 5,                  # what torchvision 0.17.2 (the reference's pin) picks on CUDA (<= 20000 coordinates -> offset)
    "auto_cpu": 4,                    # what torchvision picks on CPU  (<= 4000 coordinates -> offset)
}


def postprocess(prediction: torch.Tensor, num_classes: int, conf_thre: float = 0.7, nms_thre: float = 0.45,
                class_agnostic: bool = False, nms_variant: str = "auto") -> List[Optional[torch.Tensor]]:
    """Score filter + (class-aware) NMS for a decoded head tensor [B, A, 5+nc].

    Same contract as the reference: ``prediction[:, :, :4]`` is converted to corner form IN PLACE
    (boxes.py:32-37) and the result is a list with one ``[n, 7]`` tensor
    (x1, y1, x2, y2, obj_conf, class_conf, class_pred) per image, or None when nothing passes
    ``obj*cls >= conf_thre``; rows are in descending score order (torchvision.ops.batched_nms).
    ``nms_variant`` picks torchvision's arithmetic variant (see NMS_VARIANTS); the two variants can
    disagree on a few boxes of dense scenes because of fp32 rounding of the offset boxes.
    """
    ops.require_cuda(prediction, "postprocess")
    B = prediction.shape[0]
    output: List[Optional[torch.Tensor]] = [None for _ in range(B)]
    if B == 0 or prediction.shape[1] == 0:
        return output
    work = prediction
    if prediction.dtype != torch.float32 or not prediction.is_contiguous():
        work = prediction.float().contiguous()
    variant = _lib.NMS_AGNOSTIC if class_agnostic else NMS_VARIANTS[nms_variant]
    dets, _, counts = ops.postprocess_device(work, num_classes, conf_thre, nms_thre, variant, inplace_xyxy=True)
    if work is not prediction:
        prediction[:, :, :4] = work[:, :, :4].to(prediction.dtype)
    counts_h = counts.cpu().tolist()  # the one device->host sync of the call
    for i, n in enumerate(counts_h):
        if n > 0:
            out = dets[i, :n]
            output[i] = out if prediction.dtype == torch.float32 else out.to(prediction.dtype)
    return output


def bboxes_iou(bboxes_a: torch.Tensor, bboxes_b: torch.Tensor, xyxy: bool = True) -> torch.Tensor:
    if bboxes_a.shape[1] != 4 or bboxes_b.shape[1] != 4:
        raise IndexError
    out = ops.bboxes_iou_device(bboxes_a, bboxes_b, xyxy)
    return out if bboxes_a.dtype == torch.float32 else out.to(bboxes_a.dtype)
