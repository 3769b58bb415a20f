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


def _cuda_device_visible() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests need a B200.  On a host without a CUDA device they are SKIPPED (plain `pytest` stays
    green on the CPU container) unless TDG_REQUIRE_GPU=1 asks for a hard failure.  TDG_TEST_LIB (a test-only
    override, see tests/emu/README.md) points the test bodies at the emulated library instead."""
    if os.environ.get("TDG_TEST_LIB") or os.environ.get("TDG_REQUIRE_GPU") == "1" or _cuda_device_visible():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible (the renderer has no CPU fallback); run under gpurun")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def gpulib():
    """The CUDA extension.  No fallback: a missing .so or device fails the test loudly.  TDG_TEST_LIB is
    the tests' own hook for running the same test bodies against tests/emu (the product package has no
    such override)."""
    alt = os.environ.get("TDG_TEST_LIB")
    if alt:
        from termdaw_b200._capi import Library
        import termdaw_b200
        lib = Library(alt, "tdg_")
        termdaw_b200._lib = lib            # helpers that call termdaw_b200.load_library() see the same library
        return lib
    import termdaw_b200
    return termdaw_b200.load_library()
