import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cuda_lib():
    """Loads the C-ABI library; GPU tests must run on the native path, never on a fallback."""
    import torch
    from l3ac_b200 import _lib
    assert torch.cuda.is_available(), "gpu-marked test started without a CUDA device"
    return _lib.load()
