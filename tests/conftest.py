import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def native_lib():
    """Build (if stale) and load the C-ABI library; CPU-only boxes can still load it."""
    from scdrs_b200 import build, _native

    build.build_native()
    return _native.lib()
