import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "nerf-workspaces-explorer_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The tests exercise the in-tree libnwx.so; build it (same Makefile as __graft_entry__.build()) if a
    fresh checkout does not have it yet.  There is still no fallback: a failed build fails the session."""
    import nwx
    if not os.path.exists(nwx._lib.LIB_PATH):
        nwx.build()
    yield


def load_golden(name):
    """Fixture written by tests/golden/make_golden.py from the reference's own outputs."""
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: (torch.from_numpy(z[k]) if z[k].ndim else z[k].item()) for k in z.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


def bits_equal(a: torch.Tensor, b: torch.Tensor) -> bool:
    """Bitwise equality (NaN == NaN)."""
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    view = {1: torch.uint8, 4: torch.int32, 8: torch.int64}[a.element_size()]
    return torch.equal(a.contiguous().view(view), b.contiguous().view(view))
