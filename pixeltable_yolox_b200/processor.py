"""YoloxProcessor with the reference's interface (yolox/models/processor.py:13-60).
Image -> tensor stays host-side numpy/cv2 exactly like the reference's ValTransform/preproc
(yolox/data/data_augment.py:140-156, 234-241); postprocess runs the sm_100a kernels."""
from __future__ import annotations

from typing import Iterable, TypedDict, Union

import numpy as np
import torch

from . import boxes
from .config import YoloxConfig


class Detections(TypedDict):
    bboxes: list
    scores: list
    labels: list


def letterbox(img: np.ndarray, input_size, dtype=np.float32) -> np.ndarray:
    """Resize keeping the aspect ratio into the top-left corner of a 114-grey canvas, HWC uint8 ->
    CHW float32 in 0..255 (no mean/std), data_augment.py:140-156. ``dtype=np.uint8`` keeps the
    pixels as bytes (same values, a quarter of the host->device traffic; the stem kernel converts)."""
    import cv2

    if img.ndim == 3:
        canvas = np.full((input_size[0], input_size[1], 3), 114, dtype=np.uint8)
    else:
        canvas = np.full(tuple(input_size), 114, dtype=np.uint8)
    r = min(input_size[0] / img.shape[0], input_size[1] / img.shape[1])
    nh, nw = int(img.shape[0] * r), int(img.shape[1] * r)
    canvas[:nh, :nw] = cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR).astype(np.uint8)
    return np.ascontiguousarray(canvas.transpose(2, 0, 1), dtype=dtype)


class YoloxProcessor:
    config: YoloxConfig

    def __init__(self, model_name_or_config: Union[str, YoloxConfig]):
        if isinstance(model_name_or_config, str):
            self.config = YoloxConfig.get_named_config(model_name_or_config)
        elif isinstance(model_name_or_config, YoloxConfig):
            self.config = model_name_or_config
        else:
            raise ValueError("model_name_or_config must be a string or YoloxConfig")
        self.nms_variant = "auto"
        self.dtype = torch.float32      # torch.uint8: upload bytes, YoloxModule converts on the device
        # a CUDA device: the letterbox (resize + pad + HWC->CHW) runs there too, bit-exact with cv2 (csrc/yx_preproc.cu);
        # only the decoded image bytes cross PCIe and the returned tensor is already on the device
        self.device = None

    def __call__(self, inputs: Iterable) -> torch.Tensor:
        if self.device is not None and torch.device(self.device).type == "cuda":
            from . import ops

            return ops.letterbox_u8([np.asarray(im) for im in inputs], self.config.test_size, torch.device(self.device),
                                    torch.uint8 if self.dtype == torch.uint8 else torch.float32)
        npdt = np.uint8 if self.dtype == torch.uint8 else np.float32
        return torch.stack([torch.from_numpy(letterbox(np.array(im), self.config.test_size, npdt)) for im in inputs])

    def postprocess(self, images: Iterable, tensor: torch.Tensor, threshold: float = 0.5) -> list:
        outputs = boxes.postprocess(tensor, self.config.num_classes, threshold, self.config.nmsthre,
                                    class_agnostic=False, nms_variant=self.nms_variant)
        return self._format(images, outputs)

    def format_detections(self, images: Iterable, dets: torch.Tensor, counts: torch.Tensor) -> list:
        """Detections of YoloxModule.detect() ([B, max_det, 7] rows + per-image counts) in the reference's result
        format (processor.py:46-59): one device->host copy for the whole batch."""
        n = counts.cpu().tolist()
        rows = dets[:, :max(max(n), 1)].cpu() if len(n) else dets.cpu()
        return self._format(images, [rows[i, :k] if k > 0 else None for i, k in enumerate(n)])

    def _format(self, images: Iterable, outputs) -> list:
        results = []
        for i, image in enumerate(images):
            ratio = min(self.config.test_size[0] / image.height, self.config.test_size[1] / image.width)
            if outputs[i] is None:
                results.append(Detections(bboxes=[], scores=[], labels=[]))
                continue
            rows = outputs[i].float().cpu()        # one device->host copy per image
            bx = (rows[:, :4] / ratio).tolist()
            results.append(Detections(
                bboxes=[tuple(b) for b in bx],
                scores=[float(r[4]) * float(r[5]) for r in rows.tolist()],
                labels=[int(r[6]) for r in rows.tolist()],
            ))
        return results
