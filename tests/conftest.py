import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("a test marked gpu ran without a CUDA device")
    return torch.device("cuda", 0)


@pytest.fixture(scope="session")
def wr_ctx(cuda_device):
    import worldrenderer_b200 as wr
    return wr.NVDiffRastContextWrapper(str(cuda_device), "cuda")
