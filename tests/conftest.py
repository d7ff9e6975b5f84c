import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: minutes of CPU time")


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """Build (or reuse) the in-tree native libraries once per session."""
    from ecneproject_b200 import build
    build.build_host()
    build.build_oracle()
    if build.nvcc_path():
        build.build_engine()
    yield
