import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: longer CPU test")


@pytest.fixture(scope="session")
def lib():
    """The built C-ABI library.  Built on demand on CPU boxes (nvcc cross-compiles)."""
    from gmpnp_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


def admissible_state(rng, n, ns, nu, V=-12.5):
    """Random admissible nodal state: u = exp(N(0,0.5)) rescaled so that
    sum_i nu_i u_i <= 0.95, potential ~ U(V, 0)."""
    import numpy as np
    u = np.exp(rng.normal(0.0, 0.5, size=(n, ns + 1)))
    S = u[:, :ns] @ nu
    scale = np.minimum(1.0, 0.95 / np.maximum(S, 1e-300))
    u[:, :ns] *= scale[:, None]
    u[:, ns] = rng.uniform(V, 0.0, size=n)
    return u
