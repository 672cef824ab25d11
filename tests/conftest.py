import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle as O
    return O.port()


@pytest.fixture(scope="session")
def ref():
    from oracle import oracle as O
    r = O.reference()
    if r is None:
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return r
