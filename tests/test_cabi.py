"""The C-ABI library loads on a GPU-less host and exports every symbol include/yx_b200.h declares."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def header_functions():
    text = (ROOT / "include" / "yx_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(yx_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from pixeltable_yolox_b200 import _lib

    names = header_functions()
    assert len(names) >= 25
    assert sorted(_lib.SIGNATURES) == names, set(names) ^ set(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    from pixeltable_yolox_b200 import _lib

    handle = _lib.lib()
    raw = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in header_functions():
        assert hasattr(raw, name), f"{name} declared in yx_b200.h but not exported"
    assert handle.yx_version() == 100
    assert handle.yx_strerror(-4).decode().startswith("no sm_100 device")


def test_conv_desc_layout_matches_header(tmp_path):
    """sizeof / field offsets of yx_conv_desc and yx_bneck_desc as gcc sees the header must equal the ctypes mirrors."""
    import subprocess

    from pixeltable_yolox_b200._lib import BneckDesc, ConvDesc

    probes = [("yx_conv_desc", ConvDesc, ["in", "w", "bias", "out", "res", "ups", "head_out", "head_stride", "out2_begin",
                                          "out2", "head_cand", "head_counts", "head_conf_thre", "head_xyxy"]),
              ("yx_bneck_desc", BneckDesc, ["x", "w1", "bias2", "out", "out_ld"])]
    lines = []
    for cname, _, fields in probes:
        lines.append(f'printf("%zu\\n", sizeof({cname}));')
        lines += [f'printf("%zu\\n", offsetof({cname}, {f}));' for f in fields]
    src = tmp_path / "probe.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "yx_b200.h"\nint main(void) {\n' + "\n".join(lines) + "\nreturn 0; }\n")
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = []
    for _, cls, fields in probes:
        want.append(ctypes.sizeof(cls))
        want += [getattr(cls, "in_" if f == "in" else f).offset for f in fields]
    assert got == want


def test_no_gpu_is_a_loud_error_not_a_fallback():
    import torch

    from pixeltable_yolox_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert _lib.lib().yx_device_check(0) == -4          # YX_ERR_NO_DEVICE
    assert b"no CUDA device" in _lib.lib().yx_last_error() or _lib.lib().yx_last_error()


def test_plan_lane_api_validates_arguments_without_a_gpu():
    """yx_plan_begin_lane / end_lane / join_lanes are pure host bookkeeping: usable (and checked) on a GPU-less host."""
    from pixeltable_yolox_b200 import _lib

    lib = _lib.lib()
    plan = lib.yx_plan_create()
    try:
        assert lib.yx_plan_begin_lane(plan, 0, -1) == -1           # lane ids are 1..8
        assert lib.yx_plan_begin_lane(plan, 9, -1) == -1
        assert lib.yx_plan_begin_lane(plan, 1, 0) == -1            # op 0 does not exist yet
        assert b"not an existing op" in lib.yx_last_error()
        assert lib.yx_plan_begin_lane(plan, 1, -1) == 0
        assert lib.yx_plan_end_lane(plan) == 0
        assert lib.yx_plan_join_lanes(plan) == 0
        assert lib.yx_plan_num_ops(plan) == 0
    finally:
        lib.yx_plan_destroy(plan)
