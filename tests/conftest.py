import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The checker: the reference's own code when oracle/_ref is present, else the plain-C restatement."""
    import helpers

    return helpers.best_oracle()


@pytest.fixture(scope="session")
def port():
    import helpers

    return helpers.port_oracle()


@pytest.fixture(scope="session")
def ctx():
    import cpecan_b200 as cp

    c = cp.Context(0)
    yield c
    c.close()
