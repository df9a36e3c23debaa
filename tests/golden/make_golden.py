"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) and the
torchvision build it calls on the seeded cases of tests/cases.py. Build-container only:
    python tests/golden/make_golden.py
The committed .npz files are what pins the oracle (tests/test_oracle_golden.py) and, through it,
the CUDA path; the GPU box never sees /root/reference."""
import sys
from pathlib import Path

import numpy as np
import torch
import torchvision

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from golden._ref_shim import import_reference  # noqa: E402

import cases  # noqa: E402
from oracle import yolox_oracle as yo  # noqa: E402

OUT = Path(__file__).resolve().parent
ref = import_reference()
torch.set_num_threads(8)


def pack_list(prefix, arrays, out):
    out[prefix + "_n"] = np.array([-1 if a is None else len(a) for a in arrays], dtype=np.int64)
    for i, a in enumerate(arrays):
        if a is not None:
            out[f"{prefix}_{i}"] = np.asarray(a)


def gen_postprocess():
    from torchvision.ops import boxes as tvb

    out = {"torchvision": np.array(torchvision.__version__), "torch": np.array(torch.__version__)}
    for name in cases.POST_CASES:
        pred, conf, nms = cases.post_case(name)
        out[f"{name}/in_sha"] = np.array(cases.checksum(pred))
        # (1) the reference's own postprocess on CPU (torchvision picks the variant by candidate count)
        t = torch.from_numpy(pred.copy())
        res = ref["boxes"].postprocess(t, 80, conf, nms, class_agnostic=False)
        pack_list(f"{name}/ref_dets", [None if r is None else r.numpy() for r in res], out)
        out[f"{name}/xyxy_sha"] = np.array(cases.checksum(t[:, :, :4].numpy()))
        res_ag = ref["boxes"].postprocess(torch.from_numpy(pred.copy()), 80, conf, nms, class_agnostic=True)
        pack_list(f"{name}/ref_agnostic", [None if r is None else r.numpy() for r in res_ag], out)
        # (2) both torchvision variants explicitly, as kept ANCHOR indices
        t = torch.from_numpy(pred.copy())
        box = torch.stack([t[..., 0] - t[..., 2] / 2, t[..., 1] - t[..., 3] / 2, t[..., 0] + t[..., 2] / 2,
                           t[..., 1] + t[..., 3] / 2], -1)
        for variant, fn in (("offset", tvb._batched_nms_coordinate_trick), ("per_class", tvb._batched_nms_vanilla)):
            kept = []
            for b in range(t.shape[0]):
                conf_c, cls_c = torch.max(t[b, :, 5:85], 1)
                score = t[b, :, 4] * conf_c
                m = score >= conf
                idx = torch.nonzero(m).flatten()
                if idx.numel() == 0:
                    kept.append(None)
                    continue
                k = fn(box[b][idx], score[idx], cls_c[idx].float(), nms)
                kept.append(idx[k].numpy())
            pack_list(f"{name}/{variant}_idx", kept, out)
    np.savez_compressed(OUT / "postprocess.npz", **out)
    print("postprocess.npz", len(out))


def gen_simota():
    out = {}
    head = ref["YoloxHead"](80)
    for name, (g, seed) in cases.SIMOTA_MATCH_CASES.items():
        from pixeltable_yolox_b200.synthetic import simota_case

        cost, ious = simota_case(g, seed)
        n = cost.shape[1]
        fg_mask = torch.ones(n, dtype=torch.bool)
        num_fg, cls, pious, minds = head.simota_matching(torch.from_numpy(cost), torch.from_numpy(ious),
                                                         torch.arange(g).float(), g, fg_mask)
        out[f"{name}/in_sha"] = np.array(cases.checksum(cost) + cases.checksum(ious))
        out[f"{name}/fg"] = fg_mask.numpy()
        out[f"{name}/matched"] = minds.numpy()
        out[f"{name}/ious"] = pious.numpy()
        out[f"{name}/num_fg"] = np.array(num_fg)
    for name in cases.SIMOTA_ASSIGN_CASES:
        pred, lab, hw = cases.assign_case(name)
        from oracle.simota_oracle import anchor_grid

        xs, ys, st = anchor_grid(hw, cases.STRIDES)
        out[f"{name}/in_sha"] = np.array(cases.checksum(pred) + cases.checksum(lab))
        tp = torch.from_numpy(pred)
        for b in range(pred.shape[0]):
            gts = lab[b][lab[b].sum(1) > 0]
            G = len(gts)
            if G == 0:
                continue
            r = head.get_assignments(b, G, torch.from_numpy(gts[:, 1:5]), torch.from_numpy(gts[:, 0]), tp[b, :, :4],
                                     torch.from_numpy(st)[None], torch.from_numpy(xs)[None], torch.from_numpy(ys)[None],
                                     tp[:, :, 5:], tp[:, :, 4:5])
            gcls, fg, pious, minds, num_fg = r
            out[f"{name}/{b}/fg"] = fg.numpy()
            out[f"{name}/{b}/matched"] = minds.numpy()
            out[f"{name}/{b}/ious"] = pious.numpy()
            out[f"{name}/{b}/cls"] = gcls.numpy()
            out[f"{name}/{b}/num_fg"] = np.array(num_fg)
    # bboxes_iou
    rng = np.random.default_rng(77)
    a = rng.uniform(0, 300, size=(40, 4)).astype(np.float32); b = rng.uniform(0, 300, size=(55, 4)).astype(np.float32)
    a[:, 2:] += a[:, :2]; b[:, 2:] += b[:, :2]
    out["iou/xyxy"] = ref["boxes"].bboxes_iou(torch.from_numpy(a), torch.from_numpy(b), True).numpy()
    out["iou/cxcywh"] = ref["boxes"].bboxes_iou(torch.from_numpy(a), torch.from_numpy(b), False).numpy()
    out["iou/a"], out["iou/b"] = a, b
    np.savez_compressed(OUT / "simota.npz", **out)
    print("simota.npz", len(out))


def gen_network():
    from pixeltable_yolox_b200.synthetic import images

    out = {}
    for name, c in cases.NET_CASES.items():
        m = ref["YoloxModule"](
            ref["YoloPafpn"](c["depth"], c["width"], depthwise=c["depthwise"]),
            ref["YoloxHead"](80, c["width"], depthwise=c["depthwise"]))
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.eps = 1e-3
        x = torch.from_numpy(images(c["batch"], c["h"], c["w"], seed=c["seed"] + 500))
        sd = yo.seeded_state_dict(m.state_dict(), c["seed"], (c["h"], c["w"]), calib_x=x)
        m.load_state_dict(sd)
        m.eval()
        with torch.no_grad():
            y = m(x)
            feats = m.backbone(x)
        out[f"{name}/out"] = y.numpy()
        out[f"{name}/pan2_sha"] = np.array(cases.checksum(feats[0].numpy()))
        out[f"{name}/keys"] = np.array(len(sd))
        out[f"{name}/sd_sha"] = np.array(cases.checksum(np.concatenate([v.float().reshape(-1).numpy() for v in sd.values()])))
        # training-branch head tensor (decoded boxes, raw logits) for the SimOTA path
        m.head.decode_in_inference = False
        with torch.no_grad():
            out[f"{name}/undecoded"] = m(x).numpy()
    np.savez_compressed(OUT / "network.npz", **out)
    print("network.npz", len(out))


def gen_config1():
    """BASELINE.json config 1: yolox_nano 416x416 batch 1 on CPU through the reference's own user API --
    YoloxProcessor.__call__ (letterbox) -> YoloxModule.forward -> YoloxProcessor.postprocess -- on a seeded 480x640 PIL image,
    with the seeded non-degenerate weights of cases.CONFIG1 (the state_dict is re-derived from the seed by the tests)."""
    from PIL import Image
    from yolox.config import YoloxConfig
    from yolox.models import YoloxProcessor

    c = cases.CONFIG1
    cfg = YoloxConfig.get_named_config(c["name"])
    cfg.model = None
    m = cfg.get_model()
    img = Image.fromarray(cases.config1_image())
    proc = YoloxProcessor(c["name"])
    x = proc([img])
    sd = yo.seeded_state_dict(m.state_dict(), c["seed"], tuple(cfg.test_size), calib_x=x)
    m.load_state_dict(sd)
    m.eval()
    with torch.no_grad():
        y = m(x)
    out = {"x_sha": np.array(cases.checksum(x.numpy())), "x_shape": np.array(x.shape), "out": y.numpy().copy(),
           "keys": np.array(len(sd)),
           "sd_sha": np.array(cases.checksum(np.concatenate([v.float().reshape(-1).numpy() for v in sd.values()])))}
    for thr in c["thresholds"]:
        det = proc.postprocess([img], y.clone(), threshold=thr)[0]
        out[f"t{thr}/bboxes"] = np.array(det["bboxes"], dtype=np.float64).reshape(-1, 4)
        out[f"t{thr}/scores"] = np.array(det["scores"], dtype=np.float64)
        out[f"t{thr}/labels"] = np.array(det["labels"], dtype=np.int64)
        print("config1 thr", thr, "detections", len(det["labels"]))
    np.savez_compressed(OUT / "config1.npz", **out)
    print("config1.npz", len(out))


def gen_coco():
    """CocoEvaluator.convert_to_coco_format (yolox/evaluators/coco_evaluator.py:205-251) of the unmodified reference on the
    reference's own postprocess output of a seeded case; `self` is a stand-in carrying the two attributes the method reads."""
    import types

    from yolox.evaluators.coco_evaluator import CocoEvaluator

    c = cases.COCO_CASE
    pred, conf, nms = cases.post_case(c["post_case"])
    res = ref["boxes"].postprocess(torch.from_numpy(pred.copy()), 80, conf, nms, class_agnostic=False)
    hs, ws, ids, class_ids = cases.coco_case_meta(len(res))
    me = types.SimpleNamespace(img_size=c["img_size"],
                               dataloader=types.SimpleNamespace(dataset=types.SimpleNamespace(class_ids=class_ids)))
    data_list, image_wise = CocoEvaluator.convert_to_coco_format(me, [None if r is None else r.clone() for r in res],
                                                                 [torch.tensor(hs), torch.tensor(ws)], torch.tensor(ids),
                                                                 return_outputs=True)
    out = {"in_sha": np.array(cases.checksum(pred)), "n": np.array(len(data_list)),
           "image_id": np.array([d["image_id"] for d in data_list], dtype=np.int64),
           "category_id": np.array([d["category_id"] for d in data_list], dtype=np.int64),
           "bbox": np.array([d["bbox"] for d in data_list], dtype=np.float64).reshape(-1, 4),
           "score": np.array([d["score"] for d in data_list], dtype=np.float64),
           "wise_ids": np.array(sorted(image_wise.keys()), dtype=np.int64),
           "wise_first_bbox": np.array([image_wise[k]["bboxes"][0] for k in sorted(image_wise.keys())], dtype=np.float64)}
    np.savez_compressed(OUT / "coco.npz", **out)
    print("coco.npz", len(data_list), "rows")


def loss_case_inputs(name):
    """pred / labels / anchor grid / raw regression outputs of a SIMOTA_ASSIGN case (the inverse decode of pred)."""
    from oracle.simota_oracle import anchor_grid

    pred, lab, hw = cases.assign_case(name)
    xs, ys, st = anchor_grid(hw, cases.STRIDES)
    origin = np.stack([pred[..., 0] / st - xs, pred[..., 1] / st - ys, np.log(pred[..., 2] / st), np.log(pred[..., 3] / st)],
                      -1).astype(np.float32)
    return pred, lab, xs, ys, st, origin


def gen_losses():
    """YoloxHead.get_losses of the unmodified reference (values + autograd gradients) on the SimOTA assign cases,
    with the assignment it used (captured from its own get_assignments) so the loss path can be checked on its own."""
    out = {}
    for name in cases.SIMOTA_ASSIGN_CASES:
        pred, lab, xs, ys, st, origin = loss_case_inputs(name)
        B, A, _ = pred.shape
        for l1 in (False, True):
            head = ref["YoloxHead"](80)
            head.use_l1 = l1
            cap = []
            orig = head.get_assignments

            def wrapped(*a, **k):
                r = orig(*a, **k)
                cap.append((a[0], r))
                return r

            head.get_assignments = wrapped
            tp = torch.from_numpy(pred).clone().requires_grad_(True)
            to = torch.from_numpy(origin).clone().requires_grad_(True)
            res = head.get_losses(None, [torch.from_numpy(xs)[None]], [torch.from_numpy(ys)[None]],
                                  [torch.from_numpy(st)[None]], torch.from_numpy(lab), tp, [to], torch.float32)
            res[0].backward()
            key = f"{name}/l1_{int(l1)}"
            out[key + "/losses"] = np.array([float(v) for v in res], dtype=np.float64)
            g = tp.grad.numpy()
            fg = np.zeros((B, A), dtype=bool); mg = np.full((B, A), -1, dtype=np.int32)
            mi = np.zeros((B, A), dtype=np.float32); mc = np.zeros((B, A), dtype=np.int32)
            for b, (gcls, fgm, pious, minds, num_fg) in cap:
                fgn = fgm.numpy()
                fg[b] = fgn; mg[b, fgn] = minds.numpy(); mi[b, fgn] = pious.numpy(); mc[b, fgn] = gcls.numpy().astype(np.int32)
            nz = g.copy(); nz[..., 4] = 0; nz[fg] = 0
            assert not nz.any(), "reference gradient outside the foreground rows / objectness column"
            if not l1:
                out[f"{name}/fg"], out[f"{name}/matched_gt"] = fg, mg
                out[f"{name}/matched_iou"], out[f"{name}/matched_cls"] = mi, mc
                out[f"{name}/in_sha"] = np.array(cases.checksum(pred) + cases.checksum(lab) + cases.checksum(origin))
            out[key + "/grad_obj"] = g[..., 4].copy()
            out[key + "/grad_fg"] = g[fg].copy()
            if l1:
                out[key + "/grad_origin_fg"] = to.grad.numpy()[fg].copy()
    np.savez_compressed(OUT / "losses.npz", **out)
    print("losses.npz", len(out))


def gen_trainblock():
    """The reference's BaseConv (yolox/models/network_blocks.py:27-52) in train mode on seeded inputs: forward output, the
    BatchNorm running statistics it leaves behind, and the autograd gradients w.r.t. input, conv weight, gamma and beta
    (what Trainer.train_one_iter's loss.backward(), core/trainer.py:104-118, computes per block). fp32 on CPU."""
    out = {}
    for name, (B, ci, co, H, W, k, s, seed) in cases.TRAIN_BLOCK_CASES.items():
        x, w, gamma, beta, go = cases.train_block_inputs(name)
        from yolox.models.network_blocks import BaseConv as RefBaseConv      # the unmodified reference module

        blk = RefBaseConv(ci, co, k, s).train()
        with torch.no_grad():
            blk.conv.weight.copy_(torch.from_numpy(w)); blk.bn.weight.copy_(torch.from_numpy(gamma)); blk.bn.bias.copy_(torch.from_numpy(beta))
        xt = torch.from_numpy(x).requires_grad_(True)
        y = blk(xt)
        y.backward(torch.from_numpy(go))
        out[name + "_y"] = y.detach().numpy()
        out[name + "_dx"] = xt.grad.numpy()
        out[name + "_dw"] = blk.conv.weight.grad.numpy()
        out[name + "_dgamma"] = blk.bn.weight.grad.numpy()
        out[name + "_dbeta"] = blk.bn.bias.grad.numpy()
        out[name + "_running_mean"] = blk.bn.running_mean.numpy().copy()
        out[name + "_running_var"] = blk.bn.running_var.numpy().copy()
        out[name + "_eps_momentum"] = np.array([blk.bn.eps, blk.bn.momentum], dtype=np.float64)
    np.savez_compressed(OUT / "trainblock.npz", **out)
    print("trainblock.npz", len(out))


if __name__ == "__main__":
    which = sys.argv[1:] or ["postprocess", "simota", "network", "losses", "config1", "coco", "trainblock"]
    if "trainblock" in which:
        gen_trainblock()
    if "postprocess" in which:
        gen_postprocess()
    if "simota" in which:
        gen_simota()
    if "network" in which:
        gen_network()
    if "losses" in which:
        gen_losses()
    if "config1" in which:
        gen_config1()
    if "coco" in which:
        gen_coco()
