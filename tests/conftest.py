import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box via gpurun)")


@pytest.fixture(scope="session")
def ctx():
    """One llb200 context for the GPU tests (fails loudly when no device / library)."""
    from lego_loam_b200 import api
    c = api.Context(0)
    yield c
    c.close()
