import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    # build the C-ABI library and the oracle if they are missing / stale (nvcc cross-compiles without a GPU)
    import __graft_entry__
    __graft_entry__.build()


@pytest.fixture(scope="session")
def ctx():
    import sweepga_b200 as swg
    c = swg.Context(0)
    yield c
    c.close()
