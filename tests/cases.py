"""Seeded test cases shared by tests/golden/make_golden.py (which runs the reference on them in the
build container) and by the tests (which run the oracle and the CUDA path on them)."""
from __future__ import annotations

import numpy as np

from pixeltable_yolox_b200 import synthetic as syn

# name -> (generator, kwargs, conf_thre, nms_thre)
POST_CASES = {
    "dense_a2100_t001": (syn.dense_scene, dict(batch=2, anchors=2100, seed=11, clusters=25, size=320.0), 0.001, 0.65),
    "dense_a2100_t25": (syn.dense_scene, dict(batch=2, anchors=2100, seed=12, clusters=25, size=320.0), 0.25, 0.65),
    "dense_a8400_t001": (syn.dense_scene, dict(batch=2, anchors=8400, seed=13), 0.001, 0.65),
    "dense_a8400_t5": (syn.dense_scene, dict(batch=3, anchors=8400, seed=14), 0.5, 0.65),
    "dense_a8400_nms45": (syn.dense_scene, dict(batch=2, anchors=8400, seed=15), 0.3, 0.45),
    "sparse_a8400": (syn.sparse_scene, dict(batch=6, anchors=8400, seed=5), 0.5, 0.65),
    "sparse_a3549": (syn.sparse_scene, dict(batch=4, anchors=3549, seed=6, size=416.0), 0.25, 0.65),
    "sparse_small": (syn.sparse_scene, dict(batch=3, anchors=300, seed=7, objects=4, size=160.0), 0.3, 0.65),
    "empty": (syn.sparse_scene, dict(batch=2, anchors=500, seed=8, objects=0), 0.9, 0.65),
}

SIMOTA_MATCH_CASES = {"g1": (1, 21), "g3": (3, 22), "g17": (17, 23), "g50": (50, 24), "g120": (120, 25)}

# get_assignments end to end: (image size, label counts per image, seed)
SIMOTA_ASSIGN_CASES = {
    "s640": dict(size=640, counts=[0, 1, 7, 23, 50, 120], seed=31),
    "s416": dict(size=416, counts=[3, 12], seed=32),
}
STRIDES = (8, 16, 32)


def post_case(name):
    gen, kw, conf, nms = POST_CASES[name]
    return gen(**kw), conf, nms


def assign_case(name):
    c = SIMOTA_ASSIGN_CASES[name]
    size = c["size"]
    hw = [(size // s, size // s) for s in STRIDES]
    lab = syn.labels(len(c["counts"]), max_gt=120, seed=c["seed"], size=float(size), counts=c["counts"])
    pred = syn.train_head_output(len(c["counts"]), hw, STRIDES, lab, seed=c["seed"] + 1)
    return pred, lab, hw


# small networks for the forward golden: (depth, width, depthwise, H, W, batch, seed)
NET_CASES = {
    "w25_d33_64": dict(depth=0.33, width=0.25, depthwise=False, h=64, w=64, batch=2, seed=0),
    "w25_d33_dw_96": dict(depth=0.33, width=0.25, depthwise=True, h=96, w=64, batch=2, seed=1),
    "w50_d33_96x128": dict(depth=0.33, width=0.50, depthwise=False, h=96, w=128, batch=1, seed=2),
    "w375_d33_64": dict(depth=0.33, width=0.375, depthwise=False, h=64, w=64, batch=2, seed=3),
}


# BASELINE.json config 1: yolox_nano 416x416 batch 1 through the reference's user API (tests/golden/config1.npz)
CONFIG1 = dict(name="yolox_nano", seed=7, image_hw=(480, 640), thresholds=(0.25, 0.6))


def config1_image() -> np.ndarray:
    """Seeded 480x640 RGB uint8 image with smooth structure (so that the letterbox resize interpolates real gradients)."""
    h, w = CONFIG1["image_hw"]
    rng = np.random.default_rng(CONFIG1["seed"])
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.stack([127 + 120 * np.sin(xx / 37.0 + k) * np.cos(yy / 23.0 - k) for k in range(3)], -1)
    img += rng.normal(0, 12, size=img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


# evaluator result path (tests/golden/coco.npz): postprocess output of a seeded case -> COCO json rows
COCO_CASE = dict(post_case="sparse_a8400", img_size=(640, 640), seed=17)


def coco_case_meta(n_images: int):
    """Original image sizes, image ids and a COCO-style class-index -> category-id table (91-id space with gaps)."""
    rng = np.random.default_rng(COCO_CASE["seed"])
    hs = rng.integers(240, 1300, size=n_images).tolist()
    ws = rng.integers(240, 1300, size=n_images).tolist()
    ids = (rng.permutation(100000)[:n_images] + 1).tolist()
    class_ids = sorted(rng.permutation(91)[:80].tolist())
    return hs, ws, ids, class_ids


def checksum(a: np.ndarray) -> str:
    import hashlib

    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


# training blocks: the reference's BaseConv in train mode (conv -> BatchNorm2d with batch statistics -> SiLU), forward and the
# autograd backward of a seeded output gradient: name -> (batch, in_c, out_c, H, W, ksize, stride, seed)
TRAIN_BLOCK_CASES = {
    "c3s1_16_32": (2, 16, 32, 12, 10, 3, 1, 31),
    "c3s2_32_48": (2, 32, 48, 14, 12, 3, 2, 32),
    "c1s1_64_16": (3, 64, 16, 7, 9, 1, 1, 33),
    "c3s2_odd": (1, 16, 16, 9, 11, 3, 2, 34),
}


def train_block_inputs(name):
    """x, conv weight, gamma, beta, output gradient of a TRAIN_BLOCK case (fp32 numpy, seeded)."""
    import numpy as np

    B, ci, co, H, W, k, s, seed = TRAIN_BLOCK_CASES[name]
    r = np.random.default_rng(seed)
    pad = (k - 1) // 2
    oh, ow = (H + 2 * pad - k) // s + 1, (W + 2 * pad - k) // s + 1
    x = r.standard_normal((B, ci, H, W)).astype(np.float32)
    w = (r.standard_normal((co, ci, k, k)) * 0.2).astype(np.float32)
    gamma = (r.random(co) + 0.5).astype(np.float32)
    beta = (r.standard_normal(co) * 0.2).astype(np.float32)
    go = r.standard_normal((B, co, oh, ow)).astype(np.float32)
    return x, w, gamma, beta, go
