import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import ctypes
        cudart = ctypes.CDLL("libcudart.so.12")
        n = ctypes.c_int(0)
        return cudart.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def grm10k():
    """The reference's 1000-sample x 10k-marker quick-test set (docs/installation.md:110)."""
    from oracle import oracle as O
    bed, N0, M0, chrs = O.read_bed(os.path.join(GOLDEN, "grm10k"))
    return dict(bed=bed, N0=N0, M0=M0, chrs=np.array([int(c) for c in chrs]), prefix=os.path.join(GOLDEN, "grm10k"))


@pytest.fixture(scope="session")
def chr22():
    from oracle import oracle as O
    bed, N0, M0, chrs = O.read_bed(os.path.join(GOLDEN, "chr22_1000"))
    return dict(bed=bed, N0=N0, M0=M0, chrs=np.array([int(c) for c in chrs]), prefix=os.path.join(GOLDEN, "chr22_1000"))
