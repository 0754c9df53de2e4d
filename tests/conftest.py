"""pytest configuration: the `gpu` marker and library availability helpers."""
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _have_cuda() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def have_cuda():
    return _have_cuda()


def pytest_collection_modifyitems(config, items):
    # A gpu test on a box without CUDA must FAIL loudly only when explicitly selected with -m gpu there;
    # in the CPU container `-m "not gpu"` deselects them.
    pass
