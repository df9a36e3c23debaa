"""Builds/loads oracle/_build/liboracle.so (gcc). Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
SO = HERE / "_build" / "liboracle.so"
_lib = None


def build(force: bool = False) -> Path:
    src = HERE / "oracle.c"
    if force or not SO.exists() or SO.stat().st_mtime < src.stat().st_mtime:
        SO.parent.mkdir(exist_ok=True)
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", str(src), "-o", str(SO), "-lm"], check=True)
    return SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(SO))
        _lib.oracle_nms.restype = C.c_int
        _lib.oracle_nms.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        _lib.oracle_simota_matching.restype = C.c_int
        _lib.oracle_simota_matching.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    return _lib
