import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _has_gpu():
    """driver-level probe (no torch import: the CPU suite should start in milliseconds)"""
    import ctypes
    try:
        cu = ctypes.CDLL("libcuda.so.1")
        if cu.cuInit(0) != 0:
            return False
        n = ctypes.c_int(0)
        return cu.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def cv2_golden():
    return np.load(os.path.join(GOLDEN, "cv2_primitives.npz"))


@pytest.fixture(scope="session")
def mini_golden():
    return np.load(os.path.join(GOLDEN, "pipeline_mini.npz"))


MINI = dict(width=376, height=240, nfeatures=400, nlevels=6, fx=229.327, fy=228.648, cx=183.6, cy=124.2,
            baseline=0.110074)


@pytest.fixture(scope="session")
def mini_cfg():
    return dict(MINI)


@pytest.fixture(scope="session")
def euroc_pair():
    from fasttrack_b200 import synth
    sc = synth.StereoScene(seed=2)
    return sc.pair()
