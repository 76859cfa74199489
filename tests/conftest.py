import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib_built():
    from deepliif_b200 import build
    return build.build_library()


@pytest.fixture(autouse=True, scope="session")
def _bounded_cpu_threads():
    """The oracle runs small convolutions at N=1: oversubscribing a 128-core GPU host makes them much slower."""
    import torch
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    yield
