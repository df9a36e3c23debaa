"""GPU diagnostic: detect() (fused score filter + sort/NMS) == forward() + stand-alone postprocess over a sweep of named
configs, input sizes, batch sizes, dtypes and thresholds (bit-exact rows)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import pixeltable_yolox_b200 as yx  # noqa: E402
from pixeltable_yolox_b200 import synthetic as syn  # noqa: E402

dev = torch.device("cuda", 0)
bad = 0
for name, size in (("yolox_nano", 416), ("yolox_tiny", 416), ("yolox_s", 640), ("yolox_s", 352)):
    cfg = yx.YoloxConfig.get_named_config(name)
    cfg.model = None
    torch.manual_seed(1)
    model = cfg.get_model().to(dev)
    syn.randomize_and_calibrate(model, syn.images(2, size, size, seed=5), seed=2)
    for dt in (torch.bfloat16, torch.float16):
        m = model.to(dt).eval()
        for B in (1, 3, 7):
            x = torch.from_numpy(syn.images(B, size, size, seed=40 + B)).to(dev)
            pred = m(x)
            for thr in (0.01, 0.3, 0.6):
                want = yx.postprocess(pred.clone(), 80, thr, 0.65)
                dets, _, cnt = m.detect(x, conf_thre=thr, nms_thre=0.65)
                ok = all((w is None and int(cnt[b]) == 0) or (w is not None and int(cnt[b]) == len(w) and torch.equal(dets[b, :len(w)], w))
                         for b, w in enumerate(want))
                bad += not ok
                print(f"{name:10s} {size} {str(dt)[6:]:8s} B={B} thr={thr}: kept {[int(c) for c in cnt]} {'ok' if ok else 'MISMATCH'}", flush=True)
        m.invalidate_engine()
        model = model.float()
print("mismatches:", bad)
sys.exit(1 if bad else 0)
