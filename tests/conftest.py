import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def kb():
    """The product package bound to cuda:0; GPU tests only."""
    import kami_b200
    from kami_b200 import api

    if kami_b200.device_count() < 1:
        pytest.fail("no CUDA device: GPU tests must run on the B200 box (there is no CPU fallback)")
    api.init(0)
    return kami_b200
