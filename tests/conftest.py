import os
import sys
import warnings

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
warnings.filterwarnings("ignore", category=SyntaxWarning)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_csr(z, prefix):
    shape = tuple(int(v) for v in z[prefix + "_shape"])
    return sp.csr_matrix((z[prefix + "_data"], z[prefix + "_indices"], z[prefix + "_indptr"]), shape=shape)


@pytest.fixture(scope="session")
def graph_l4():
    z = np.load(os.path.join(GOLDEN, "ref_graph_l4.npz"))
    n = len(z["sizes"])
    return dict(z=z, perm=z["perm"], sizes=z["sizes"], A=load_csr(z, "A"),
                L=[load_csr(z, "L%d" % i) for i in range(n)], Lt=[load_csr(z, "Lt%d" % i) for i in range(n)])


@pytest.fixture(scope="session")
def graph_l1():
    z = np.load(os.path.join(GOLDEN, "ref_graph_l1.npz"))
    n = len(z["sizes"])
    return dict(z=z, perm=z["perm"], sizes=z["sizes"], A=load_csr(z, "A"),
                L=[load_csr(z, "L%d" % i) for i in range(n)], Lt=[load_csr(z, "Lt%d" % i) for i in range(n)])


@pytest.fixture(scope="session")
def layer_cases():
    z = np.load(os.path.join(GOLDEN, "layer_cases.npz"))
    cases = {}
    for key in z.files:
        name, field = key.split(".", 1)
        cases.setdefault(name, {})[field] = z[key]
    return cases


def rel_inf(a, b):
    """||a - b||_inf / ||b||_inf -- the parity norm of BASELINE.md section 4."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    d = np.abs(b).max()
    return float(np.abs(a - b).max() / (d if d > 0 else 1.0))
