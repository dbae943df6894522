import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """the in-tree shared library, built with nvcc if it is missing or stale (no GPU needed to build)"""
    import __graft_entry__ as ge
    return ge.build()


@pytest.fixture(scope="session")
def dev(built_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from libdmet_preview_b200.device import get_device
    return get_device()
