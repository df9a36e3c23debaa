"""sort_nms_kernel: one CTA per image (YX_NMS_CLUSTER=1) against the thread-block-cluster kernel (host's choice of 2 / 4 / 8
CTAs per image) on the dense scenes of config 5, with equality of the kept rows checked on every case.
usage: python tools/gpu_nms_cluster.py [out.json]"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from pixeltable_yolox_b200 import ops, synthetic as syn  # noqa: E402
from pixeltable_yolox_b200.boxes import NMS_VARIANTS  # noqa: E402

dev = torch.device("cuda", 0)
A = 8400


def timed(pred, thr, reps=20):
    """median device time (us) of filter + sort_nms, and of the filter alone (nms threshold path identical)"""
    work = [pred.clone() for _ in range(reps + 3)]
    for w in work[:3]:
        out = ops.postprocess_device(w, 80, thr, 0.65, NMS_VARIANTS["auto"])
    torch.cuda.synchronize()
    ts = []
    for w in work[3:]:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = ops.postprocess_device(w, 80, thr, 0.65, NMS_VARIANTS["auto"])
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts)), out


def all_kept(batch, classes):
    grid = np.zeros((batch, A, 85), dtype=np.float32)
    gx, gy = np.meshgrid(np.arange(100), np.arange(84))
    grid[:, :, 0] = (gx.reshape(-1) * 6.0 + 3.0)[None]; grid[:, :, 1] = (gy.reshape(-1) * 6.0 + 3.0)[None]
    grid[:, :, 2:4] = 9.0                                                  # neighbours overlap a little (IoU 0.2): divisions run
    rng = np.random.default_rng(10)
    grid[:, :, 4] = rng.uniform(0.5, 1.0, (batch, A))
    np.put_along_axis(grid[:, :, 5:], rng.integers(0, classes, (batch, A))[..., None], 0.9, axis=2)
    return grid


rows = []
PHASES_ONLY = "--phases-only" in sys.argv
if "--flush-sweep" in sys.argv:
    # batch rows at which a survivor batch is resolved (YX_NMS_FLUSH): time + equality of the kept rows against the default
    dense = torch.from_numpy(syn.dense_scene(64, anchors=A, seed=13)).to(dev)
    for thr in (0.001, 0.25, 0.5):
        base_t, base_o = timed(dense, thr)
        cnt = base_o[2].cpu().tolist()
        for fl in ("128", "192", "256", "320", "384", "448", "512"):
            os.environ["YX_NMS_FLUSH"] = fl
            t, o = timed(dense, thr)
            del os.environ["YX_NMS_FLUSH"]
            same = bool(torch.equal(o[2], base_o[2])) and all(bool(torch.equal(o[1][b, :cnt[b]], base_o[1][b, :cnt[b]])) for b in range(64))
            print(json.dumps(dict(conf_thre=thr, flush_rows=int(fl), postprocess_us=round(t, 1), default_us=round(base_t, 1), rows_equal=same)), flush=True)
    sys.exit(0)
for batch in (() if PHASES_ONLY else (64, 16, 4, 1)):
    dense = torch.from_numpy(syn.dense_scene(batch, anchors=A, seed=13)).to(dev)
    kept3 = torch.from_numpy(all_kept(batch, 3)).to(dev)
    for name, pred, thr in (("dense", dense, 0.001), ("dense", dense, 0.25), ("dense", dense, 0.5), ("all_kept_3_classes", kept3, 0.01)):
        os.environ["YX_NMS_CLUSTER"] = "1"
        t1, o1 = timed(pred, thr)
        del os.environ["YX_NMS_CLUSTER"]
        tc, oc = timed(pred, thr)
        same = bool(torch.equal(o1[2], oc[2]))
        cnt = o1[2].cpu().tolist()
        for b in range(batch):
            same = same and bool(torch.equal(o1[0][b, :cnt[b]], oc[0][b, :cnt[b]])) and bool(torch.equal(o1[1][b, :cnt[b]], oc[1][b, :cnt[b]]))
        row = dict(scene=name, batch=batch, conf_thre=thr, kept_per_image=float(np.mean(cnt)), postprocess_us_one_cta=round(t1, 1),
                   postprocess_us_cluster=round(tc, 1), speedup=round(t1 / tc, 2), rows_equal=same)
        print(json.dumps(row), flush=True)
        rows.append(row)
for batch in (64, 1):
    dense = torch.from_numpy(syn.dense_scene(batch, anchors=A, seed=13)).to(dev)
    for cap in ("1", "2", None):
        if cap:
            os.environ["YX_NMS_CLUSTER"] = cap
        ops.postprocess_device(dense.clone(), 80, 0.001, 0.65, NMS_VARIANTS["auto"])
        torch.cuda.synchronize()
        os.environ["YX_NMS_DEBUG"] = "1"
        print(f"== phases (clk), batch {batch}, cluster cap {cap or 'auto'}: dense thr 0.001", flush=True)
        ops.postprocess_device(dense.clone(), 80, 0.001, 0.65, NMS_VARIANTS["auto"])
        torch.cuda.synchronize()
        del os.environ["YX_NMS_DEBUG"]
        os.environ.pop("YX_NMS_CLUSTER", None)
if len(sys.argv) > 1 and not PHASES_ONLY:
    Path(sys.argv[1]).write_text(json.dumps(rows, indent=1))
assert all(r["rows_equal"] for r in rows), "cluster kernel keeps different rows"
