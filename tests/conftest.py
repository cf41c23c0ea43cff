import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def native_lib():
    """The product library must exist (built in-tree by __graft_entry__.build / phyloacc_b200.build)."""
    import phyloacc_b200 as pab

    return pab.load_library()


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import coracle

    return coracle.lib()
