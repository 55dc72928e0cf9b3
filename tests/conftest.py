import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def cg():
    import cryogrid_jl_b200
    return cryogrid_jl_b200


@pytest.fixture(scope="session")
def orc():
    import oracle as _o   # oracle/oracle.py -- the CPU checker (test infrastructure)
    _o.lib()
    return _o
