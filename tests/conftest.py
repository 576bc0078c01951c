import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def golden(name):
    import torch

    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


@pytest.fixture(scope="session")
def lib():
    """The built shared library (built on demand in the CPU container)."""
    from fastdm_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        from fastdm_b200.build import build

        build(verbose=False)
    return _lib.load()
