import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_small.npz")))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def ref(oracle):
    """Compiled reference Cython (oracle/_ref) or None when it has not been built here."""
    return oracle.load_ref()


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import tsr_b200  # noqa: F401
    from tsr_b200 import _lib
    assert _lib.lib().tsr_device_supported() == 1, "libtsr_b200.so loaded but the device is not sm_100"
    return torch.device("cuda")
