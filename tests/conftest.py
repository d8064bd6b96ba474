import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def cuda_device(request):
    import torch

    if not torch.cuda.is_available():
        expr = request.config.getoption("markexpr", default="") or ""
        if "gpu" in expr and "not gpu" not in expr:   # `-m gpu` was asked for explicitly: a missing device is a failure
            pytest.fail("GPU tests selected (-m gpu) but no CUDA device is visible")
        pytest.skip("needs a CUDA device (run with -m gpu on the GPU box)")
    return 0
