import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/libtermdaw_oracle.so), built on demand.  Test infrastructure only."""
    from helpers import load_oracle
    return load_oracle()


@pytest.fixture(scope="session")
def gpulib():
    """The CUDA extension.  No fallback: a missing .so or device fails the test loudly."""
    import termdaw_b200
    return termdaw_b200.load_library()
