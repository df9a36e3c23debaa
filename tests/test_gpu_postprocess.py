"""-m gpu: score filter + NMS through the C-ABI; kept rows must be BIT-EXACT against the committed
reference goldens and the oracle (integer/index work and fp32 compares)."""
import numpy as np
import pytest
import torch

import cases
from conftest import unpack_list

pytestmark = pytest.mark.gpu

import pixeltable_yolox_b200 as yx  # noqa: E402
from oracle import postprocess_oracle as po  # noqa: E402
from pixeltable_yolox_b200 import _lib, ops  # noqa: E402
from pixeltable_yolox_b200 import synthetic as syn  # noqa: E402


def _check_lists(got, want):
    assert len(got) == len(want)
    for i, (g, w) in enumerate(zip(got, want)):
        assert (g is None) == (w is None), f"image {i}"
        if g is not None:
            np.testing.assert_array_equal(g.cpu().numpy(), w, err_msg=f"image {i}")


@pytest.mark.parametrize("name", list(cases.POST_CASES))
def test_postprocess_matches_reference_golden(cuda, name, golden_post):
    pred, conf, nms = cases.post_case(name)
    t = torch.from_numpy(pred).to(cuda)
    got = yx.postprocess(t, 80, conf, nms, nms_variant="auto_cpu")       # the variant the CPU reference took
    _check_lists(got, unpack_list(golden_post, f"{name}/ref_dets"))
    assert cases.checksum(t[:, :, :4].cpu().numpy()) == str(golden_post[f"{name}/xyxy_sha"])   # in-place xyxy
    got = yx.postprocess(torch.from_numpy(pred).to(cuda), 80, conf, nms, class_agnostic=True)
    _check_lists(got, unpack_list(golden_post, f"{name}/ref_agnostic"))


@pytest.mark.parametrize("name", list(cases.POST_CASES))
@pytest.mark.parametrize("variant", ["offset", "per_class"])
def test_nms_variants_kept_indices(cuda, name, variant, golden_post):
    pred, conf, nms = cases.post_case(name)
    t = torch.from_numpy(pred).to(cuda)
    _, idx, cnt = ops.postprocess_device(t, 80, conf, nms, yx.boxes.NMS_VARIANTS[variant])
    cnt = cnt.cpu().tolist()
    for b, want in enumerate(unpack_list(golden_post, f"{name}/{variant}_idx")):
        if want is None:
            assert cnt[b] == 0
        else:
            np.testing.assert_array_equal(idx[b, :cnt[b]].cpu().numpy(), want)


def test_auto_variant_follows_torchvision_cuda_rule(cuda):
    pred = syn.dense_scene(2, anchors=8400, seed=41)
    want, _ = po.postprocess(pred.copy(), 80, 0.3, 0.65, variant="offset", return_indices=True)
    _check_lists(yx.postprocess(torch.from_numpy(pred).to(cuda), 80, 0.3, 0.65), want)


def test_auto_tv017_variant_follows_the_pinned_torchvision_rule(cuda):
    """torchvision 0.17.2 (poetry.lock pin of the reference) switches to per-class NMS above 20 000 box coordinates on
    CUDA; the installed 0.26 (which generated the goldens) above 100 000. Image 0 has > 5 000 candidates, image 1 fewer."""
    pred = syn.dense_scene(2, anchors=8400, seed=43)
    pred[1, 3000:, 4] = 0.0                                     # < 5 000 candidates -> offset trick under both rules
    per_class, _ = po.postprocess(pred.copy(), 80, 0.05, 0.65, variant="per_class", return_indices=True)
    offset, _ = po.postprocess(pred.copy(), 80, 0.05, 0.65, variant="offset", return_indices=True)
    oracle, _ = po.postprocess(pred.copy(), 80, 0.05, 0.65, variant="auto_tv017", return_indices=True)
    got = yx.postprocess(torch.from_numpy(pred).to(cuda), 80, 0.05, 0.65, nms_variant="auto_tv017")
    _check_lists(got, [per_class[0], offset[1]])
    _check_lists(got, oracle)


def test_large_anchor_count_takes_global_sort_path(cuda):
    """A > 16384 candidates: keys no longer fit shared memory."""
    pred = syn.dense_scene(1, anchors=20000, seed=42, clusters=150, size=1280.0)
    for variant in ("offset", "per_class"):
        want, _ = po.postprocess(pred.copy(), 80, 0.001, 0.65, variant=variant, return_indices=True)
        _check_lists(yx.postprocess(torch.from_numpy(pred).to(cuda), 80, 0.001, 0.65, nms_variant=variant), want)


def test_ties_and_degenerate_boxes(cuda):
    """identical scores (stable order), zero-area boxes (0/0 -> NaN -> kept), IoU exactly at threshold."""
    A = 64
    pred = np.zeros((1, A, 85), dtype=np.float32)
    pred[0, :, 0:2] = 100.0
    pred[0, :, 2:4] = 20.0
    pred[0, :, 4] = 0.9
    pred[0, :, 5] = 0.5                      # every score identical
    pred[0, 10:20, 2:4] = 0.0                # zero-area boxes
    pred[0, 30:40, 0] += np.arange(10, dtype=np.float32) * 7.0
    pred[0, 40:, 5 + 3] = 0.8                # another class, higher score
    for variant in ("offset", "per_class"):
        want, _ = po.postprocess(pred.copy(), 80, 0.1, 0.5, variant=variant, return_indices=True)
        _check_lists(yx.postprocess(torch.from_numpy(pred).to(cuda), 80, 0.1, 0.5, nms_variant=variant), want)


def test_negative_coordinates_disable_class_gating(cuda):
    """With negative coordinates the offset trick lets boxes of adjacent classes overlap; the kernel must
    then compare across classes exactly like torchvision does."""
    pred = syn.dense_scene(2, anchors=3000, seed=46, clusters=30, size=320.0)
    pred[:, :, 0:2] -= 250.0
    for variant in ("offset", "per_class"):
        want, _ = po.postprocess(pred.copy(), 80, 0.2, 0.5, variant=variant, return_indices=True)
        _check_lists(yx.postprocess(torch.from_numpy(pred).to(cuda), 80, 0.2, 0.5, nms_variant=variant), want)


def test_single_class_dense_scene_overflows_the_shared_kept_list(cuda):
    """Every box in one class, low overlap: thousands kept, more than the shared-memory list holds."""
    rng = np.random.default_rng(47)
    A = 12000
    pred = np.zeros((1, A, 85), dtype=np.float32)
    pred[0, :, 0:2] = rng.uniform(0, 4000, size=(A, 2))
    pred[0, :, 2:4] = rng.uniform(10, 30, size=(A, 2))
    pred[0, :, 4] = rng.uniform(0.5, 1.0, size=A)
    pred[0, :, 5 + 7] = rng.uniform(0.5, 1.0, size=A)
    for variant in ("offset", "per_class"):
        want, _ = po.postprocess(pred.copy(), 80, 0.1, 0.3, variant=variant, return_indices=True)
        assert len(want[0]) > 7000
        _check_lists(yx.postprocess(torch.from_numpy(pred).to(cuda), 80, 0.1, 0.3, nms_variant=variant), want)


def test_empty_batch_and_no_candidates(cuda):
    assert yx.postprocess(torch.zeros(0, 100, 85, device=cuda), 80) == []
    out = yx.postprocess(torch.zeros(3, 100, 85, device=cuda), 80, 0.5, 0.65)
    assert out == [None, None, None]


def test_score_filter_compact_is_ordered(cuda):
    pred = syn.dense_scene(2, anchors=3000, seed=43)
    cand, idx, cnt = ops.score_filter_compact(torch.from_numpy(pred).to(cuda), 80, 0.3)
    p = pred.copy()
    cls = p[:, :, 5:]
    for b in range(2):
        conf = cls[b].max(1); lab = cls[b].argmax(1)
        score = p[b, :, 4] * conf
        keep = np.where(score >= np.float32(0.3))[0]
        n = int(cnt[b])
        assert n == len(keep)
        np.testing.assert_array_equal(idx[b, :n].cpu().numpy(), keep)
        got = cand[b, :n].cpu().numpy()
        np.testing.assert_array_equal(got[:, 4], p[b, keep, 4]); np.testing.assert_array_equal(got[:, 5], conf[keep])
        np.testing.assert_array_equal(got[:, 6], lab[keep].astype(np.float32)); np.testing.assert_array_equal(got[:, 7], score[keep])


def test_batched_nms_standalone(cuda):
    rng = np.random.default_rng(44)
    B, n = 3, 700
    boxes = rng.uniform(0, 200, size=(B, n, 4)).astype(np.float32); boxes[..., 2:] += boxes[..., :2]
    scores = rng.uniform(0, 1, size=(B, n)).astype(np.float32)
    cls = rng.integers(0, 5, size=(B, n)).astype(np.int32)
    counts = np.array([700, 0, 333], dtype=np.int32)
    for variant, name in ((0, "offset"), (1, "per_class")):
        keep, kc = ops.batched_nms(torch.from_numpy(boxes).to(cuda), torch.from_numpy(scores).to(cuda),
                                   torch.from_numpy(cls).to(cuda), torch.from_numpy(counts).to(cuda), 0.5, variant)
        for b in range(B):
            c = counts[b]
            want = po.batched_nms(boxes[b, :c], scores[b, :c], cls[b, :c].astype(np.int64), 0.5, name) if c else np.empty(0, np.int64)
            assert int(kc[b]) == len(want)
            np.testing.assert_array_equal(keep[b, :len(want)].cpu().numpy(), want)


def test_processor_postprocess_matches_oracle(cuda):
    from PIL import Image

    proc = yx.YoloxProcessor("yolox_s")
    pred = syn.sparse_scene(2, anchors=8400, seed=45)
    images = [Image.new("RGB", (640, 480)), Image.new("RGB", (480, 640))]
    res = proc.postprocess(images, torch.from_numpy(pred).to(cuda), threshold=0.5)
    want = po.postprocess(pred.copy(), 80, 0.5, 0.65, variant="offset")
    for im, r, w in zip(images, res, want):
        ratio = min(640 / im.height, 640 / im.width)
        if w is None:
            assert r["bboxes"] == []
            continue
        assert r["labels"] == [int(v) for v in w[:, 6]]
        np.testing.assert_allclose(np.array(r["bboxes"], dtype=np.float32), w[:, :4] / np.float32(ratio), rtol=1e-6)
        np.testing.assert_allclose(r["scores"], [float(a) * float(b) for a, b in zip(w[:, 4], w[:, 5])], rtol=0)


@pytest.mark.parametrize("variant", ["offset", "per_class", "auto"])
def test_single_warp_path_equals_chunked_path_and_oracle(cuda, variant, monkeypatch):
    """n <= 32 candidates take the single-warp path (registers only); the chunked shared-memory path on the same input
    (YX_NMS_NO_SMALL) and the oracle must keep the same rows: ties, zero-area boxes, negative coordinates, n = 1 / 32."""
    rng = np.random.default_rng(9)
    for n_img, A in ((3, 32), (2, 24), (2, 1), (1, 40)):
        pred = np.zeros((n_img, A, 85), dtype=np.float32)
        pred[:, :, 0:2] = rng.uniform(-5, 120, (n_img, A, 2))
        pred[:, :, 2:4] = rng.uniform(10, 60, (n_img, A, 2))
        pred[:, :, 4] = np.round(rng.uniform(0.3, 1.0, (n_img, A)), 1)            # many tied scores
        cls = rng.integers(0, 4, (n_img, A))
        np.put_along_axis(pred[:, :, 5:], cls[..., None], 0.9, axis=2)
        pred[0, :3, 2:4] = 0.0                                                    # zero-area boxes
        if A > 8:
            pred[0, 5:8] = pred[0, 4]                                             # exact duplicates
        want, _ = po.postprocess(pred.copy(), 80, 0.2, 0.5, variant="offset" if variant == "auto" else variant, return_indices=True)
        got = yx.postprocess(torch.from_numpy(pred).to(cuda), 80, 0.2, 0.5, nms_variant=variant)
        _check_lists(got, want)
        monkeypatch.setenv("YX_NMS_NO_SMALL", "1")
        _check_lists(yx.postprocess(torch.from_numpy(pred).to(cuda), 80, 0.2, 0.5, nms_variant=variant), want)
        monkeypatch.delenv("YX_NMS_NO_SMALL")


def test_every_anchor_kept_stays_in_shared_memory(cuda):
    """8 400 disjoint boxes (nothing suppresses anything; what conf 0.01 gives on random weights): the kept list of a
    640^2 image fits the shared-memory list, all rows come back in score order."""
    A = 8400
    pred = np.zeros((2, A, 85), dtype=np.float32)
    gx, gy = np.meshgrid(np.arange(100), np.arange(84))
    pred[:, :, 0] = (gx.reshape(-1) * 6.0 + 3.0)[None]; pred[:, :, 1] = (gy.reshape(-1) * 6.0 + 3.0)[None]
    pred[:, :, 2:4] = 5.0
    rng = np.random.default_rng(10)
    pred[:, :, 4] = rng.uniform(0.5, 1.0, (2, A))
    np.put_along_axis(pred[:, :, 5:], rng.integers(0, 80, (2, A))[..., None], 0.9, axis=2)
    want, _ = po.postprocess(pred.copy(), 80, 0.01, 0.65, variant="offset", return_indices=True)
    assert all(len(w) == A for w in want)
    _check_lists(yx.postprocess(torch.from_numpy(pred).to(cuda), 80, 0.01, 0.65, nms_variant="offset"), want)


def _cluster_scenes():
    rng = np.random.default_rng(21)
    scenes = {}
    scenes["dense_8400"] = (syn.dense_scene(2, anchors=8400, seed=13), 0.001, 0.65)
    scenes["dense_8400_thr3"] = (syn.dense_scene(3, anchors=8400, seed=15), 0.3, 0.45)
    neg = syn.dense_scene(2, anchors=3000, seed=46, clusters=30, size=320.0)
    neg[:, :, 0:2] -= 250.0                                                       # class window J > 0
    scenes["negative_coords"] = (neg, 0.2, 0.5)
    A = 12000                                                                     # one class, thousands kept: list overflow
    one = np.zeros((1, A, 85), dtype=np.float32)
    one[0, :, 0:2] = rng.uniform(0, 4000, size=(A, 2)); one[0, :, 2:4] = rng.uniform(10, 30, size=(A, 2))
    one[0, :, 4] = rng.uniform(0.5, 1.0, size=A); one[0, :, 5 + 7] = rng.uniform(0.5, 1.0, size=A)
    scenes["one_class_12000"] = (one, 0.1, 0.3)
    A = 2500                                                                      # tied scores, duplicates, zero areas
    tie = np.zeros((2, A, 85), dtype=np.float32)
    tie[:, :, 0:2] = rng.uniform(-20, 600, (2, A, 2)); tie[:, :, 2:4] = rng.uniform(10, 90, (2, A, 2))
    tie[:, :, 4] = np.round(rng.uniform(0.3, 1.0, (2, A)), 1)
    np.put_along_axis(tie[:, :, 5:], rng.integers(0, 6, (2, A))[..., None], 0.9, axis=2)
    tie[0, :40, 2:4] = 0.0; tie[0, 100:140] = tie[0, 99]
    tie[1, 600:, 4] = 0.0                                                         # ragged: image 1 has 600 candidates
    scenes["ties_2500"] = (tie, 0.2, 0.5)
    few = syn.dense_scene(2, anchors=8400, seed=17)
    few[0, 40:, 4] = 0.0; few[1, 33:, 4] = 0.0                                    # 33 .. 40 candidates: short runs, empty runs
    scenes["few_of_8400"] = (few, 0.0005, 0.65)
    return scenes


@pytest.mark.parametrize("scene", ["dense_8400", "dense_8400_thr3", "negative_coords", "one_class_12000", "ties_2500", "few_of_8400"])
def test_cluster_nms_equals_single_cta_and_oracle(cuda, scene, monkeypatch):
    """sort_nms_kernel<true> (one thread-block cluster of 2 / 4 / 8 CTAs per image: split sort with a rank merge through
    distributed shared memory, kept list / mask rows dealt over the CTAs) keeps exactly the rows of the single-CTA
    kernel and of the oracle, for every variant."""
    pred, conf, nms = _cluster_scenes()[scene]
    for variant in ("offset", "per_class", "agnostic"):
        agn = variant == "agnostic"
        want, _ = po.postprocess(pred.copy(), 80, conf, nms, variant="offset" if agn else variant, class_agnostic=agn,
                                 return_indices=True)
        rows = {}
        for R in (1, 2, 4, 8):
            monkeypatch.setenv("YX_NMS_CLUSTER", str(R))
            dets, idx, cnt = ops.postprocess_device(torch.from_numpy(pred).to(cuda), 80, conf, nms,
                                                    _lib.NMS_AGNOSTIC if agn else yx.boxes.NMS_VARIANTS[variant])
            cnt = cnt.cpu().tolist()
            rows[R] = [(dets[b, :cnt[b]].cpu().numpy(), idx[b, :cnt[b]].cpu().numpy()) for b in range(len(cnt))]
        monkeypatch.delenv("YX_NMS_CLUSTER")
        for b, w in enumerate(want):
            for R in (1, 2, 4, 8):
                d, _ = rows[R][b]
                if w is None:
                    assert d.shape[0] == 0, f"{variant} R={R} image {b}"
                else:
                    np.testing.assert_array_equal(d, w, err_msg=f"{variant} R={R} image {b}")
                np.testing.assert_array_equal(rows[R][b][1], rows[1][b][1], err_msg=f"{variant} R={R} image {b} (indices)")
