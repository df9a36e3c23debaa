"""On-GPU diagnostic sweep (not a pytest file): every conv geometry the network uses, the tcgen05
kernel and the SIMT kernel against torch's fp32 conv on the same rounded operands. Prints a table;
writes gpurun_out/conv_diag.txt."""
import sys
import time
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from pixeltable_yolox_b200 import _lib, ops  # noqa: E402
from pixeltable_yolox_b200.ops import View  # noqa: E402

dev = torch.device("cuda", 0)
lines = []


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    lines.append(s)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    with open(ROOT / "gpurun_out" / "conv_diag.txt", "a") as f:
        f.write(s + "\n")


def one(B, H, W, cin, cout, k, s, dtype, res=False, ups=False, in_extra=0, out_extra=0, act="silu", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = (torch.randn(B, H, W, cin + in_extra, generator=g) * 1.0).to(dev).to(dtype)
    w = (torch.randn(cout, k * k, cin, generator=g) / (k * k * cin) ** 0.5).to(dev).to(dtype)
    bias = torch.randn(cout, generator=g).to(dev)
    pad = (k - 1) // 2
    oh = (H + 2 * pad - k) // s + 1
    ow = (W + 2 * pad - k) // s + 1
    xin = View(x, in_extra // 2 // 8 * 8, cin)
    r = None
    rv = None
    if res:
        r = torch.randn(B, oh, ow, cout, generator=g).to(dev).to(dtype)
        rv = View(r)
    # torch reference on the rounded operands
    xt = xin.torch().float().permute(0, 3, 1, 2)
    wt = w.float().reshape(cout, k, k, cin).permute(0, 3, 1, 2)
    y = F.conv2d(xt, wt, bias, s, pad)
    y = {"silu": F.silu, None: lambda t: t}[act](y)
    if res:
        y = y + r.float().permute(0, 3, 1, 2)
    y = y.permute(0, 2, 3, 1).contiguous()
    out = {}
    for name, simt in (("tc", False), ("simt", True)):
        o = torch.full((B, oh, ow, cout + out_extra), 7.0, device=dev, dtype=dtype)
        ov = View(o, out_extra // 2 // 8 * 8, cout)
        u = uv = None
        if ups:
            u = torch.full((B, 2 * oh, 2 * ow, cout), 7.0, device=dev, dtype=dtype)
            uv = View(u)
        try:
            ops.conv_bn_act(xin, w, bias, ov, k, s, _lib.ACT_CODES[act], res=rv, ups=uv, simt=simt)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            out[name] = f"EXC {e}"
            continue
        got = ov.torch().float()
        err = (got - y).abs() / y.abs().clamp_min(1.0)
        msg = f"max {err.max().item():.3e} mean {err.mean().item():.3e}"
        if not torch.isfinite(got).all():
            msg += " NONFINITE"
        if ups:
            ue = (u.float()[:, ::2, ::2] - got).abs().max().item() + (u.float()[:, 1::2, 1::2] - got).abs().max().item()
            msg += f" ups {ue:.1e}"
        if out_extra:
            untouched = (o[..., :ov.c_off] == 7).all().item() and (o[..., ov.c_off + cout:] == 7).all().item()
            msg += f" slice_ok {untouched}"
        if err.max().item() > 0.05:
            bad = (err > 0.05).nonzero()
            msg += f" BAD n={bad.shape[0]} first={bad[0].tolist()} got={got[tuple(bad[0])].item():.4f} want={y[tuple(bad[0])].item():.4f}"
            # which rows/cols are bad
            rows = bad[:, 1].unique().tolist()[:12]; cols = bad[:, 2].unique().tolist()[:12]; ch = bad[:, 3].unique().tolist()[:12]
            msg += f" h={rows} w={cols} c={ch}"
        out[name] = msg
    log(f"B{B} {H}x{W} cin{cin} cout{cout} k{k} s{s} {str(dtype)[6:]} res{int(res)} ups{int(ups)} ie{in_extra} oe{out_extra} | tc: {out['tc']} | simt: {out['simt']}")


def main():
    log(torch.cuda.get_device_name(0), torch.version.cuda)
    bf = torch.bfloat16
    # 1x1 flat
    one(2, 16, 16, 64, 64, 1, 1, bf)
    one(2, 16, 16, 64, 64, 1, 1, torch.float16)
    one(1, 20, 12, 128, 128, 1, 1, bf)          # M=240: partial tile
    one(2, 16, 16, 32, 32, 1, 1, bf)            # SW64
    one(2, 16, 16, 16, 16, 1, 1, bf)            # SW32
    one(2, 16, 16, 256, 256, 1, 1, bf)
    one(2, 16, 16, 512, 96, 1, 1, bf)
    one(2, 16, 16, 128, 512, 1, 1, bf)          # 2 n-tiles
    one(2, 16, 16, 64, 64, 1, 1, bf, res=True, ups=True, in_extra=64, out_extra=64)
    # 3x3 s1
    one(2, 16, 16, 64, 64, 3, 1, bf)
    one(2, 20, 20, 128, 128, 3, 1, bf)          # tile 20x6
    one(1, 40, 40, 128, 128, 3, 1, bf)
    one(1, 80, 80, 128, 256, 3, 1, bf)
    one(2, 24, 40, 32, 32, 3, 1, bf)
    one(2, 24, 40, 16, 32, 3, 1, bf)
    one(2, 16, 16, 64, 64, 3, 1, bf, res=True, in_extra=64, out_extra=64)
    one(2, 16, 16, 64, 64, 3, 1, torch.float16)
    # 3x3 s2
    one(2, 16, 16, 64, 64, 3, 2, bf)
    one(2, 40, 40, 32, 64, 3, 2, bf)
    one(1, 80, 80, 128, 128, 3, 2, bf)
    one(2, 20, 20, 256, 512, 3, 2, bf)
    one(1, 64, 48, 16, 32, 3, 2, bf)
    # fp32 simt only sanity (tc path routes fp32 to simt)
    one(2, 16, 16, 32, 48, 3, 1, torch.float32)
    # quick timing of the heavy layer: 3x3 128->128 @80x80, B=64
    B = 64
    x = torch.randn(B, 80, 80, 128, device=dev).to(bf)
    w = (torch.randn(128, 9, 128, device=dev) / 34).to(bf)
    bias = torch.zeros(128, device=dev)
    o = torch.empty(B, 80, 80, 128, device=dev, dtype=bf)
    for (cin, cout, k, hw) in ((128, 128, 3, 80), (128, 128, 1, 80), (256, 256, 3, 40), (64, 64, 3, 160)):
        x = torch.randn(B, hw, hw, cin, device=dev).to(bf)
        w = (torch.randn(cout, k * k, cin, device=dev) / (k * k * cin) ** 0.5).to(bf)
        bias = torch.zeros(cout, device=dev)
        o = torch.empty(B, hw, hw, cout, device=dev, dtype=bf)
        for _ in range(3):
            ops.conv_bn_act(View(x), w, bias, View(o), k, 1, 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 10
        for _ in range(n):
            ops.conv_bn_act(View(x), w, bias, View(o), k, 1, 1)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fl = 2.0 * B * hw * hw * cin * cout * k * k
        by = (x.numel() + o.numel()) * 2
        log(f"time B64 {hw}x{hw} {cin}->{cout} k{k}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s  {by / ms / 1e6:.0f} GB/s (in+out)")


if __name__ == "__main__":
    main()
