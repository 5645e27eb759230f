import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """Both native artefacts are built in-tree before any test runs (no GPU needed to compile)."""
    import __graft_entry__ as g
    # always: the make rules track every source and header, so an up-to-date tree costs a no-op `make`, and
    # an edited kernel can never be tested against a stale binary
    g.build()
    yield
