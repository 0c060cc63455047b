import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "forwarddiff.jl_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def oracle():
    import orc
    return orc.Oracle()


@pytest.fixture(scope="session")
def oracle_nansafe():
    import orc
    return orc.Oracle(nansafe=True)


@pytest.fixture(scope="session")
def ctx():
    """The device context.  GPU tests must run the CUDA library: no GPU or no library -> hard failure."""
    import fdb200
    c = fdb200.Context(0)
    yield c
    c.close()
