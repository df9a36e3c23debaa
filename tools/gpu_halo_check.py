"""GPU diagnostic (not a pytest): halo-mode 3x3 convs against torch for both UMMA base-offset
conventions, then timings of the yolox_s 3x3 shapes with halo on/off.
usage: python tools/gpu_halo_check.py <baseoff><bres> | time [baseoff]"""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from tests.test_gpu_kernels import _conv_case  # noqa: E402
from pixeltable_yolox_b200 import ops  # noqa: E402
from pixeltable_yolox_b200.ops import View  # noqa: E402

dev = torch.device("cuda", 0)
torch.backends.cudnn.allow_tf32 = False
CASES = [
    (2, 16, 16, 64, 64, 3, 1), (2, 20, 20, 128, 128, 3, 1), (1, 40, 40, 128, 128, 3, 1), (1, 80, 80, 32, 64, 3, 1),
    (2, 24, 40, 16, 32, 3, 1), (1, 13, 13, 96, 96, 3, 1), (1, 5, 3, 64, 32, 3, 1), (3, 80, 80, 128, 256, 3, 1),
    (2, 20, 20, 256, 256, 3, 1), (5, 40, 40, 64, 64, 3, 1), (2, 160, 160, 32, 32, 3, 1), (7, 23, 61, 128, 128, 3, 1),
]
MODE = sys.argv[1] if len(sys.argv) > 1 else "00"
if MODE.startswith("s2"):
    CASES = [(2, 16, 16, 64, 64, 3, 2), (2, 40, 40, 32, 64, 3, 2), (1, 80, 80, 128, 128, 3, 2), (2, 20, 20, 256, 512, 3, 2),
             (3, 160, 160, 64, 128, 3, 2), (2, 320, 320, 32, 64, 3, 2), (5, 26, 38, 128, 256, 3, 2), (1, 7, 10, 64, 32, 3, 2),
             (2, 13, 14, 32, 48, 3, 2), (1, 64, 48, 16, 32, 3, 2)]
    MODE = "0" + MODE[2]
for base in (MODE[0],) if not MODE.startswith("time") else ():
    os.environ["YX_HALO_BASEOFF"] = base
    for bres in (MODE[1],):
        os.environ["YX_HALO_BRES"] = bres
        for case in CASES:
            try:
                err = _conv_case(dev, *case, torch.bfloat16, False, False, 0, 0, "silu", simt=False)
                torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                err = repr(e)[:200]
            print(f"baseoff={base} bres={bres} {case} err={err}", flush=True)
os.environ.pop("YX_HALO_BRES", None)
if MODE == "time2":
    B = 64
    for sh in ["32:64:320", "64:128:160", "128:256:80", "256:512:40", "128:128:80", "256:256:40"]:
        cin, cout, hw = (int(v) for v in sh.split(":"))
        x = torch.randn(B, hw, hw, cin, device=dev).to(torch.bfloat16)
        w = (torch.randn(cout, 9, cin, device=dev) / (9 * cin) ** 0.5).to(torch.bfloat16)
        bias = torch.zeros(cout, device=dev)
        o = torch.empty(B, hw // 2, hw // 2, cout, device=dev, dtype=torch.bfloat16)
        res = {}
        for halo in ("0", "1"):
            os.environ["YX_HALO_S2"] = halo
            for _ in range(3):
                ops.conv_bn_act(View(x), w, bias, View(o), 3, 2, 1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.conv_bn_act(View(x), w, bias, View(o), 3, 2, 1)
            e1.record(); torch.cuda.synchronize()
            res[halo] = e0.elapsed_time(e1) * 1e3 / 5
        fl = 2.0 * B * (hw // 2) ** 2 * cin * 9 * cout
        by = 2.0 * B * (hw * hw * cin + (hw // 2) ** 2 * cout)
        print(f"{sh:14s} old {res['0']:7.1f} us  planes {res['1']:7.1f} us  ({fl / res['1'] / 1e6:7.1f} TFLOP/s, {by / res['1'] / 1e3:7.1f} GB/s)", flush=True)
if MODE == "time":
    os.environ["YX_HALO_BASEOFF"] = sys.argv[2] if len(sys.argv) > 2 else "0"
    B = 64
    for sh in ["32:32:160", "64:64:80", "128:128:40", "256:256:20", "128:256:80", "128:128:80", "128:256:40", "128:128:20"]:
        cin, cout, hw = (int(v) for v in sh.split(":"))
        x = torch.randn(B, hw, hw, cin, device=dev).to(torch.bfloat16)
        w = (torch.randn(cout, 9, cin, device=dev) / (9 * cin) ** 0.5).to(torch.bfloat16)
        bias = torch.zeros(cout, device=dev)
        o = torch.empty(B, hw, hw, cout, device=dev, dtype=torch.bfloat16)
        res = {}
        for halo in ("0", "1"):
            os.environ["YX_HALO"] = halo
            for _ in range(3):
                ops.conv_bn_act(View(x), w, bias, View(o), 3, 1, 1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.conv_bn_act(View(x), w, bias, View(o), 3, 1, 1)
            e1.record(); torch.cuda.synchronize()
            res[halo] = e0.elapsed_time(e1) * 1e3 / 5
        fl = 2.0 * B * hw * hw * cin * 9 * cout
        print(f"{sh:14s} old {res['0']:7.1f} us  halo {res['1']:7.1f} us  ({fl / res['1'] / 1e6:7.1f} TFLOP/s)", flush=True)
print("ok")
