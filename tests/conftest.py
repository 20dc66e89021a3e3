import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def _load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return _load


def peak_err(a, b):
    """max|a - b| / max|b| — the parity metric of SURVEY.md section 8(d)."""
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


def unpack_fields(d, prefix):
    n = int(d[prefix + "_n"])
    return [(d[f"{prefix}_{i}_data"], tuple(int(v) for v in d[f"{prefix}_{i}_offset"])) for i in range(n)]


# FP64 gate of BASELINE.json north_star: <= 1e-10 peak-normalised
TOL64 = 1e-10


@pytest.fixture(params=["folded", "direct", "czt", "auto"])
def mft_variant(request):
    """Run a test under both executions of K2a (LFD_MFT_AUTO is the default: chirp-z up to 8192-point transforms, else
    the folded DMMA form; LFD_MFT_DIRECT is the plain complex x complex form); restores the default afterwards."""
    from lentil_b200 import _lib
    L = _lib.lib()
    L.lfd_set_mft_variant({"direct": 0, "folded": 1, "czt": 2, "auto": 3}[request.param])
    yield request.param
    L.lfd_set_mft_variant(3)
