import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def rmx():
    """The CUDA engine; GPU tests only.  Fails loudly (no CPU fallback) if the .so is missing."""
    import torch
    assert torch.cuda.is_available(), "GPU test run without a CUDA device"
    from radio_mapper_b200 import engine
    return engine
