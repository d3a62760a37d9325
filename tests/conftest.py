import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    g.build()
    return g.load_package()


@pytest.fixture(scope="session")
def ctx(pkg):
    c = pkg.Context(0, 20261019)
    yield c


def to_pkg_model(pkg, bm):
    """oracle model objects -> facade model objects"""
    rbms = bm if isinstance(bm, (list, tuple)) else [bm]
    out = []
    for r in rbms:
        if r.kind == "nb":
            out.append(pkg.NegativeBinomialBernoulliRBM(r.weights, r.visbias, r.hidbias, r.inversedispersion))
        else:
            out.append(pkg.BernoulliRBM(r.weights, r.visbias, r.hidbias))
    return out if isinstance(bm, (list, tuple)) else out[0]


def relerr(got, ref, floor=1e-12):
    import numpy as np
    return float(np.max(np.abs(got - ref) / (np.abs(ref) + floor)))
