import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built_library():
    """The C-ABI library, built in-tree (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    from cosmopower_b200 import _engine
    if not os.path.isfile(_engine.LIB_PATH):
        g.build()
    return _engine.load_library()
