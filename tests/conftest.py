import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


import oracle.binding  # noqa: E402,F401  -- registers the CPU checker as capi.Engine("orc") for the tests


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def built():
    """Compile the CUDA library and the oracle once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as ge
    ge.build()
    return True
