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


def test_conv_desc_layout_matches_header():
    """sizeof(yx_conv_desc) computed from the header's field list must equal the ctypes mirror."""
    from pixeltable_yolox_b200._lib import ConvDesc

    # 12 int32, then pointer/int64 pairs ... : recompute with natural alignment
    assert ctypes.sizeof(ConvDesc) == 12 * 4 + 8 * 2 + 8 + 8 + 8 * 2 + 8 * 2 + 8 * 2 + 8 + 4 * 4 + 4 + 4 + 8 * 2


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
