import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def cpu_golden():
    return dict(np.load(os.path.join(GOLDEN_DIR, "cpu_golden.npz")))


@pytest.fixture(scope="session")
def gpu_golden():
    path = os.path.join(GOLDEN_DIR, "gpu_golden.npz")
    if not os.path.isfile(path):
        pytest.skip("tests/golden/gpu_golden.npz missing (generate with make_golden.py gpu on a B200)")
    return dict(np.load(path))


@pytest.fixture(scope="session")
def cpu_golden_v1():
    return dict(np.load(os.path.join(GOLDEN_DIR, "cpu_golden_v1.npz")))


@pytest.fixture(scope="session")
def gpu_golden_v1():
    path = os.path.join(GOLDEN_DIR, "gpu_golden_v1.npz")
    if not os.path.isfile(path):
        pytest.skip("tests/golden/gpu_golden_v1.npz missing (generate with make_golden.py gpu_v1 on a B200)")
    return dict(np.load(path))


@pytest.fixture(scope="session")
def ref_iou3d():
    """pcdet/ops/iou3d compiled into oracle/_ref/iou3d_cuda.so (skips when it was not built/shipped)."""
    from oracle import ref
    if not ref.iou3d_available():
        pytest.skip("oracle/_ref/iou3d_cuda.so not built (needs /root/reference; run python oracle/build_ref.py)")
    return ref


@pytest.fixture(scope="session")
def capi():
    from oracle import capi as c
    c.load()
    return c


@pytest.fixture(scope="session")
def ref_so():
    """The unmodified reference compiled into oracle/_ref (skips when it was not built/shipped)."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference; run python oracle/build_ref.py)")
    return ref


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
