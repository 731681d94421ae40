import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def mm():
    """The product package; the CUDA library must be present (no fallback)."""
    import monkey_moore_b200 as m
    m.lib()
    return m


@pytest.fixture(scope="session")
def gpu(mm):
    if mm.device_count() == 0:
        pytest.fail("GPU test selected but no CUDA device is visible")
    return mm
