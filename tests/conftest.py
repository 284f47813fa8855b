import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.build(with_reference=os.path.isdir(o.REFERENCE_SRC))
    return o


@pytest.fixture(scope="session")
def nwb():
    import nw_b200
    return nw_b200
