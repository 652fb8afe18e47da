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


def cube_tet_mesh(n=3, scale=(0.3, 0.3, 1.0)):
    """Small structured tet mesh of a box (n^3 cubes x 6 tets), used for 3D kernel-parity cases."""
    import numpy as np
    from gmpnp_b200 import meshio
    g = np.linspace(0.0, 1.0, n + 1)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    x = np.stack([X.ravel() * scale[0], Y.ravel() * scale[1], Z.ravel() * scale[2]], axis=1)
    idx = lambda i, j, k: (i * (n + 1) + j) * (n + 1) + k
    tets = []
    for i in range(n):
        for j in range(n):
            for k in range(n):
                v = [idx(i + a, j + b, k + c) for a in (0, 1) for b in (0, 1) for c in (0, 1)]
                # Kuhn triangulation along the main diagonal v[0]-v[7]
                for p in ((1, 3), (1, 5), (2, 3), (2, 6), (4, 5), (4, 6)):
                    tets.append([v[0], v[p[0]], v[p[1]], v[7]])
    # jitter interior vertices so that the geometry is generic
    rng = np.random.default_rng(7)
    interior = np.all((x > 1e-12) & (x < np.array(scale) - 1e-12), axis=1)
    x[interior] += rng.uniform(-0.02, 0.02, size=(interior.sum(), 3)) * np.array(scale)
    return meshio.Mesh(x=x, cells=np.array(tets, dtype=np.int32), name=f"cube{n}")
