"""torchrun target (>= 2 GPUs): the fused all-reduce + SGD + EMA kernel over NVLink peer memory against
ncclAllReduce + yx_sgd_ema_step from the same state, then its time against that sequence.
usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/gpu_allreduce_sgd_check.py"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import pixeltable_yolox_b200 as yx  # noqa: E402
from pixeltable_yolox_b200 import train_conv  # noqa: E402
from pixeltable_yolox_b200.optim import FusedSgdEma  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)                               # identical parameters on every rank
model = yx.YoloxConfig.get_named_config("yolox_s").get_model().to(dev).train()
opt = FusedSgdEma(model, lr=0.01, ema=True, direct_grads=True, peer_group=dist.group.WORLD)
g = torch.Generator(device=dev).manual_seed(100 + rank)    # different gradients per rank


def fill():
    opt.flat_grad.copy_(torch.randn(opt.flat_grad.shape, generator=g, device=dev) * 0.1)


def snapshot():
    return ([v.clone() for v in model.state_dict().values()], [b.clone() for b in opt.bufs], [v.clone() for v in opt.ema.state_dict().values()])


def restore(s):
    with torch.no_grad():
        for dst, src in zip(model.state_dict().values(), s[0]): dst.copy_(src)
        for dst, src in zip(opt.bufs, s[1]): dst.copy_(src)
        for dst, src in zip(opt.ema.state_dict().values(), s[2]): dst.copy_(src)


# two eager steps through NCCL: momentum buffers, pointer table
for _ in range(2):
    fill()
    dist.all_reduce(opt.flat_grad); opt.flat_grad.div_(world)
    opt.step(0.01)
ok = True
for it in range(3):
    fill()
    grads = opt.flat_grad.clone()
    before = snapshot()
    upd = opt.updates
    # reference: NCCL all-reduce + the plain fused optimizer launch
    dist.all_reduce(opt.flat_grad); opt.flat_grad.div_(world)
    opt.set_hyper(0.02)
    opt.step_captured()
    want = snapshot()
    restore(before); opt.updates = upd
    opt.flat_grad.copy_(grads)
    opt.set_hyper(0.02)
    opt.step_allreduce_captured()
    torch.cuda.synchronize()
    got = snapshot()
    worst = 0.0
    for part in range(3):
        for a, b in zip(got[part], want[part]):
            if a.dtype.is_floating_point:
                worst = max(worst, float((a - b).abs().max() / (b.abs().max() + 1e-12)))
    ok = ok and worst < 1e-5
    if rank == 0:
        print(f"iteration {it}: fused all-reduce + SGD + EMA vs ncclAllReduce + yx_sgd_ema_step: worst relative difference {worst:.2e}", flush=True)
# parameters identical on every rank
chk = torch.stack([v.double().sum() for v in model.state_dict().values() if v.dtype.is_floating_point]).sum().reshape(1)
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
same = all(float(c) == float(allc[0]) for c in allc)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def nccl_path():
    dist.all_reduce(opt.flat_grad); opt.flat_grad.div_(world); opt.step_captured()


t_nccl = timed(nccl_path)
t_fused = timed(opt.step_allreduce_captured)
if rank == 0:
    print(f"parameters identical across {world} ranks: {same}; ncclAllReduce + div + yx_sgd_ema_step {t_nccl:.1f} us, fused kernel {t_fused:.1f} us "
          f"({opt.flat_grad.numel() * 4 / 1e6:.1f} MB of gradients)", flush=True)
    print("OK" if ok and same else "MISMATCH", flush=True)
dist.barrier()
dist.destroy_process_group()
