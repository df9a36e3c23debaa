import sys, time, faulthandler
faulthandler.dump_traceback_later(15, exit=True)
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import pixeltable_yolox_b200 as yx
from pixeltable_yolox_b200 import synthetic as syn
dev = torch.device("cuda", 0)
cfg = yx.YoloxConfig("dbg", depth=0.33, width=0.5)
model = cfg.get_model().to(dev)
syn.randomize_and_calibrate(model, syn.images(2, 160, 160, seed=1), seed=0)
model = model.bfloat16().eval()
x = torch.from_numpy(syn.images(4, 160, 160, seed=2)).to(dev)
print("eager+capture...", flush=True)
t0 = time.time()
y = model(x)
torch.cuda.synchronize()
print("first call done", time.time() - t0, flush=True)
y2 = model(x)
torch.cuda.synchronize()
print("replay done", torch.equal(y, y2), flush=True)
