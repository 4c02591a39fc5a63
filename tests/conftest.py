import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with -m gpu)")


@pytest.fixture(scope="session")
def so():
    """the CPU oracle (test infrastructure)"""
    import sgfhe_oracle
    sgfhe_oracle.build()
    return sgfhe_oracle


@pytest.fixture(scope="session")
def sg():
    """the product package (ctypes over libsgfhe_cuda.so)"""
    import sgfhe_jl_b200
    return sgfhe_jl_b200
