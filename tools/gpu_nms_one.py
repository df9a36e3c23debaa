"""Three postprocess calls on the dense scene of config 5 ([64, 8400, 85], conf 0.001): the command behind the
`ncu --set full -k regex:sort_nms -s 2 -c 1` capture summarised in profiles/r2_nms_cluster.md.
usage: python tools/gpu_nms_one.py [batch]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from pixeltable_yolox_b200 import ops, synthetic as syn  # noqa: E402
from pixeltable_yolox_b200.boxes import NMS_VARIANTS  # noqa: E402

dev = torch.device("cuda", 0)
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
pred = torch.from_numpy(syn.dense_scene(batch, anchors=8400, seed=13)).to(dev)
for _ in range(3):
    _, _, cnt = ops.postprocess_device(pred.clone(), 80, 0.001, 0.65, NMS_VARIANTS["auto"])
torch.cuda.synchronize()
print("kept per image:", float(cnt.float().mean()))
