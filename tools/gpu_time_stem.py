import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench
from pixeltable_yolox_b200 import synthetic as syn
dev = torch.device("cuda", 0)
args = bench.parse()
cfg, model = bench.build_model(args, dev)
model = model.to(torch.bfloat16).eval()
x = torch.from_numpy(syn.images(64, 640, 640, seed=7)).to(dev)
for name, xx in (("fp32", x), ("u8", x.to(torch.uint8))):
    eng = model.engine_for(xx)
    eng.forward(xx)
    prof = eng.builder.profile(); prof = eng.builder.profile()
    print(name, "stem us:", round(prof[0]["ms"] * 1e3, 1), "total us:", round(sum(p["ms"] for p in prof) * 1e3, 1))
a = model(x); b = model(x.to(torch.uint8))
print("u8 == fp32 output:", torch.equal(a, b))
