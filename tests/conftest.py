import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def built():
    """Make sure every native artefact exists (CUDA library cross-compiled by nvcc, oracle, emulation harness)."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def oracle(built):
    from _oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def default_model():
    from hygeia_b200 import model
    vartheta, dim = model.get_known_parameters()
    return dict(vartheta=vartheta, theta=model.default_theta(), dim=dim,
                alpha_beta=model.beta_parameters(model.DEFAULT_MU, model.DEFAULT_SIGMA))


def golden(name):
    import numpy as np
    return np.load(os.path.join(HERE, "golden", name), allow_pickle=False)
