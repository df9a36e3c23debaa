"""Evaluator result path: CocoEvaluator.convert_to_coco_format (yolox/evaluators/coco_evaluator.py:205-251)
fed by the detections of YoloxModule.detect(). The per-detection arithmetic (box / scale, xyxy -> xywh, score = obj * class_conf,
category lookup) and the compaction over the batch run in one kernel (csrc/yx_preproc.cu: coco_rows_kernel); ONE device->host
copy replaces the reference's per-image .cpu() and per-row .item() calls; the host only assembles the JSON dictionaries."""
from __future__ import annotations

from collections import defaultdict
from typing import Optional, Sequence

import torch

from . import ops


def convert_to_coco_format(dets: torch.Tensor, det_count: torch.Tensor, info_imgs, ids, img_size: Sequence[int],
                           class_ids: Optional[Sequence[int]] = None, return_outputs: bool = False):
    """dets [B, max_det, 7] / det_count [B] from YoloxModule.detect(); info_imgs = (heights, widths) of the original images;
    ids = image ids; class_ids = dataset.class_ids (COCO category ids by class index). Returns the reference's `data_list`
    (and `image_wise_data` with return_outputs), same keys, same values."""
    heights = torch.as_tensor(info_imgs[0], dtype=torch.float64)
    widths = torch.as_tensor(info_imgs[1], dtype=torch.float64)
    scale = torch.minimum(img_size[0] / heights, img_size[1] / widths).to(torch.float32)      # min(H / img_h, W / img_w) in double
    image_ids = torch.as_tensor([int(i) for i in ids], dtype=torch.int64)
    cid = None if class_ids is None else torch.as_tensor(list(class_ids), dtype=torch.int32)
    bbox, score, cat, iid = ops.coco_rows(dets, det_count, scale, image_ids, cid)
    bbox_l, score_l, cat_l, iid_l = bbox.tolist(), score.tolist(), cat.tolist(), iid.tolist()
    data_list = [{"image_id": i, "category_id": c, "bbox": b, "score": s, "segmentation": []}
                 for i, c, b, s in zip(iid_l, cat_l, bbox_l, score_l)]
    if not return_outputs:
        return data_list
    image_wise = defaultdict(dict)
    for i, c, b, s in zip(iid_l, cat_l, bbox_l, score_l):
        e = image_wise.setdefault(i, {"bboxes": [], "scores": [], "categories": []})
        e["bboxes"].append([b[0], b[1], b[0] + b[2], b[1] + b[3]])     # the reference records xyxy here (before xyxy2xywh)
        e["scores"].append(s)
        e["categories"].append(c)
    return data_list, image_wise
