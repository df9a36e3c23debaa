"""numpy restatement (TEST INFRASTRUCTURE, never imported by the product) of the pieces either side of the path:
  * letterbox: `preproc` (yolox/data/data_augment.py:140-156) with cv2.resize(INTER_LINEAR) on uint8 restated from OpenCV's
    published fixed-point algorithm (third-party dependency, opencv-python >= 4.10 in pyproject.toml, 4.13 installed; not
    under /root/reference). tests/test_oracle_golden.py pins this restatement against cv2 itself on seeded images.
  * coco_rows: the arithmetic of CocoEvaluator.convert_to_coco_format (yolox/evaluators/coco_evaluator.py:205-251).
  * sgd_ema: torch.optim.SGD's single-tensor update and ModelEMA.update (yolox/utils/ema.py:46-58) in numpy fp32."""
from __future__ import annotations

import numpy as np

f32 = np.float32


def _coeffs(src: int, dst: int, clamp: bool):
    scale = np.float64(src) / np.float64(dst)
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(f32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(f32)).astype(f32)
    if clamp:                                   # x only: resize.cpp clamps sx / fx, y rows are clipped when they are fetched
        lo = s < 0
        f[lo] = 0; s[lo] = 0
        hi = s >= src - 1
        f[hi] = 0; s[hi] = src - 1
    a1 = np.rint(f * f32(2048)).astype(np.int64)
    a0 = np.rint((f32(1.0) - f) * f32(2048)).astype(np.int64)
    return s, a0, a1


def resize_linear_u8(img: np.ndarray, nw: int, nh: int) -> np.ndarray:
    """cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR) for uint8 HWC / HW images."""
    squeeze = img.ndim == 2
    if squeeze:
        img = img[..., None]
    h, w = img.shape[:2]
    if w == 2 * nw and h == 2 * nh:             # exact 2:1 goes to INTER_AREA: rounded mean of the 2x2 block
        a = img.astype(np.int64)
        out = ((a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2] + 2) >> 2).astype(np.uint8)
        return out[..., 0] if squeeze else out
    sx, ax0, ax1 = _coeffs(w, nw, True)
    sy, ay0, ay1 = _coeffs(h, nh, False)
    src = img.astype(np.int64)
    rows = src[:, sx] * ax0[None, :, None] + src[:, np.minimum(sx + 1, w - 1)] * ax1[None, :, None]
    s0, s1 = rows[np.clip(sy, 0, h - 1)], rows[np.clip(sy + 1, 0, h - 1)]
    b0, b1 = ay0[:, None, None], ay1[:, None, None]
    out = np.clip((((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2, 0, 255).astype(np.uint8)
    return out[..., 0] if squeeze else out


def letterbox(img: np.ndarray, input_size, dtype=np.float32) -> np.ndarray:
    """data_augment.py:140-156: [C, H, W] (or [H, W]) in 0..255."""
    if img.ndim == 3:
        canvas = np.full((input_size[0], input_size[1], img.shape[2]), 114, dtype=np.uint8)
    else:
        canvas = np.full(tuple(input_size), 114, dtype=np.uint8)
    r = min(input_size[0] / img.shape[0], input_size[1] / img.shape[1])
    nh, nw = int(img.shape[0] * r), int(img.shape[1] * r)
    canvas[:nh, :nw] = resize_linear_u8(img, nw, nh)
    if canvas.ndim == 3:
        canvas = canvas.transpose(2, 0, 1)
    return np.ascontiguousarray(canvas, dtype=dtype)


def coco_rows(dets, img_hw, img_size, class_ids=None):
    """dets: list of [n_i, 7] fp32 arrays or None (postprocess output); img_hw: list of (h, w). Returns flat
    (bbox_xywh [N,4] fp32, score [N] fp32, category [N] int, image index [N] int)."""
    bb, sc, ct, ix = [], [], [], []
    for i, (d, (h, w)) in enumerate(zip(dets, img_hw)):
        if d is None:
            continue
        scale = f32(min(img_size[0] / float(h), img_size[1] / float(w)))
        b = (d[:, :4] / scale).astype(f32)
        b[:, 2] = b[:, 2] - b[:, 0]
        b[:, 3] = b[:, 3] - b[:, 1]
        bb.append(b); sc.append((d[:, 4] * d[:, 5]).astype(f32))
        c = d[:, 6].astype(np.int64)
        ct.append(c if class_ids is None else np.asarray(class_ids)[c]); ix.append(np.full(len(d), i))
    if not bb:
        return np.zeros((0, 4), f32), np.zeros((0,), f32), np.zeros((0,), np.int64), np.zeros((0,), np.int64)
    return np.concatenate(bb), np.concatenate(sc), np.concatenate(ct), np.concatenate(ix)
