import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (sm_100a); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden_post():
    return np.load(GOLDEN / "postprocess.npz")


@pytest.fixture(scope="session")
def golden_simota():
    return np.load(GOLDEN / "simota.npz")


@pytest.fixture(scope="session")
def golden_net():
    return np.load(GOLDEN / "network.npz")


def unpack_list(npz, prefix):
    ns = npz[prefix + "_n"]
    return [None if n < 0 else npz[f"{prefix}_{i}"] for i, n in enumerate(ns)]


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)
