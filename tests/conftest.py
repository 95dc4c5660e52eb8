import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ilsm():
    import ilsm_b200
    ilsm_b200._build.build()
    return ilsm_b200


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def cfg_small(ilsm):
    """A reduced config-1 (20k-point map) that the oracle finishes in well under a second."""
    return ilsm.synth.config1(n_map=20_000)


@pytest.fixture(scope="session")
def cfg_full(ilsm):
    """BASELINE config 1: OS0-64 frame vs 100k-point map."""
    return ilsm.synth.config1(n_map=100_000)


@pytest.fixture(scope="session")
def ctx(ilsm):
    c = ilsm.Context(0)
    yield c
    c.close()


def pose7(q, t):
    return np.concatenate([np.asarray(q, np.float64), np.asarray(t, np.float64)])
