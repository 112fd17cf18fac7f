import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def cuda_ops():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from vidchapters_b200.ops import CudaOps
    return CudaOps()


@pytest.fixture(scope="session")
def torch_ops():
    from oracle.torch_ops import TorchOps
    return TorchOps()
