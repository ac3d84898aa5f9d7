import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session")
def port():
    """plain-C restatement oracle (always buildable)"""
    from oracle import oracle as O
    if not O.available("port"):
        assert O.build("port"), "could not build oracle/libfm_oracle.so"
    return O.Oracle("port")


@pytest.fixture(scope="session")
def ref():
    """the reference's own headers (only where /root/reference is mounted or a prebuilt _ref travelled)"""
    from oracle import oracle as O
    if not O.available("ref"):
        if os.path.isdir("/root/reference/src"):
            O.build("ref")
    if not O.available("ref"):
        pytest.skip("oracle/_ref not available on this box")
    return O.Oracle("ref")


@pytest.fixture(scope="session")
def gpu_ctx():
    """engine context on cuda:0 -- fails loudly if the CUDA extension is missing"""
    from fmwr_b200 import _lib
    return _lib.Context(0)
