import hashlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_PATH = os.path.join(ROOT, "tests", "golden", "ssd_head_golden.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


@pytest.fixture(scope="session")
def golden():
    return np.load(GOLDEN_PATH)


@pytest.fixture(scope="session")
def priors_cpu():
    from oracle import head
    return head.default_boxes()


def unpack_bits(packed: np.ndarray, shape) -> torch.Tensor:
    n = int(np.prod(shape))
    return torch.from_numpy(np.unpackbits(packed)[:n].reshape(tuple(shape)).astype(bool))
