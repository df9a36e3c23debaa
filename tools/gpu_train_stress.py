"""Race hunt for the captured training step (side-stream branches, wgrad stream, direct gradient accumulation): with lr = 0 the
parameters never change, so every replay of the graph must give the same loss bit for bit and the same gradients (up to the fp32
atomics of the SPP backward). usage: python tools/gpu_train_stress.py [replays]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from pixeltable_yolox_b200 import train_conv  # noqa: E402
from pixeltable_yolox_b200.optim import FusedSgdEma  # noqa: E402

replays = int(sys.argv[1]) if len(sys.argv) > 1 else 200
sys.argv = [sys.argv[0], "--train"]
args = bench.parse()
dev = torch.device("cuda", 0)
cfg, model = bench.build_model(args, dev)
model = model.train().to(memory_format=torch.channels_last)
opt = FusedSgdEma(model, lr=0.0, direct_grads=True)
train_conv.attach_packer(model, torch.bfloat16)
x, lab, _ = bench.train_batch(args, 0, 8)
x, lab = x.to(dev).contiguous(memory_format=torch.channels_last), lab.to(dev)


def eager():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = model(x, lab)
    opt.zero_grad()
    out["total_loss"].backward()
    opt.step(0.0)
    return out["total_loss"].detach()


side = torch.cuda.Stream(dev)
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        l_eager = eager().clone()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, capture_error_mode="thread_local"):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = model(x, lab)
    loss = out["total_loss"]
    opt.zero_grad()
    loss.backward()
    opt.join()
    grads = opt.flat_grad.clone()
    opt.step_captured()
opt.set_hyper(0.0)
g.replay(); torch.cuda.synchronize()
l0, g0 = loss.detach().clone(), grads.clone()
bad_loss, worst = 0, 0.0
for i in range(replays):
    opt.set_hyper(0.0)
    g.replay()
    torch.cuda.synchronize()
    if not torch.equal(loss.detach(), l0):
        bad_loss += 1
    worst = max(worst, float((grads - g0).abs().max() / g0.abs().max()))
print(f"{replays} replays: eager loss {float(l_eager):.6f}, graph loss {float(l0):.6f}; replays with a different loss: {bad_loss}; "
      f"worst gradient deviation relative to the largest gradient: {worst:.2e}")
print("OK" if bad_loss == 0 and worst < 1e-2 else "RACE?")
