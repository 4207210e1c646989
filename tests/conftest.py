import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load


@pytest.fixture(scope="session")
def w0():
    from oracle.weights import make_weights
    return make_weights("W0")


@pytest.fixture(scope="session")
def w1():
    from oracle.weights import make_weights
    return make_weights("W1")


@pytest.fixture(scope="session")
def oracle_net_w0(w0):
    from oracle.forward import OracleNet
    return OracleNet(w0)
