import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import akaze_oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def akz():
    """The product package; building is part of __graft_entry__.build()."""
    import __graft_entry__ as G
    G.build()
    import akaze_rust_b200 as A
    return A


@pytest.fixture(scope="session")
def engine(akz):
    eng = akz.Engine(0, 4096, 4096, 4, keep_evolutions=True)
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def fixture_grays(akz):
    g = os.path.join(ROOT, "tests", "golden")
    return akz.load_gray(os.path.join(g, "1.jpg")), akz.load_gray(os.path.join(g, "2.jpg"))
