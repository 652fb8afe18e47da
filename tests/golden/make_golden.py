"""Generate the committed golden vectors of the 1D path from the CPU oracle.

    python tests/golden/make_golden.py

The reference itself cannot run here (FEniCS absent), so these vectors pin the
ORACLE (restated forms + FFC quadrature + dolfin Newton semantics); the CUDA path is compared
against them on the GPU box, where neither /root/reference nor a long oracle run is available.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

from gmpnp_b200 import meshio, params  # noqa: E402
from oracle import solver  # noqa: E402
from conftest import admissible_state  # noqa: E402


def ohp_metrics(x, u, prm):
    g = solver.p1_gradient_projection_1d(x, u[:, 6])
    field = -g[0] * prm.thermal_voltage / prm.length * 1e-9
    c_cat, c_H = u[0, 5] * prm.c0[5], u[0, 0] * prm.c0[0]
    w = (prm.n_water_cat * c_cat + prm.n_water_H * c_H) * 1e-3
    eps = prm.eps_w * ((55 - w) / 55) + 6 * (w / 55)
    return field, eps


def main():
    # 1. kernel parity: residual + block-tridiagonal Jacobian at a random admissible state
    m = meshio.graded_interval(24, 0.002, 16)
    x = m.x[:, 0]
    prm = params.params_1d()
    rng = np.random.default_rng(0)
    u = admissible_state(rng, len(x), 6, prm.nu)
    un = admissible_state(rng, len(x), 6, prm.nu)
    disc = solver.Discretisation(x, m.cells, 7)
    bd, bv = solver.bc_1d(len(x), 7, prm.V)
    F = solver.apply_bc_residual(disc.residual(u.ravel(), un.ravel(), prm, prm.jflux), u.ravel(), bd, bv)
    A = solver.apply_bc_matrix(disc.jacobian(u.ravel(), prm), bd).toarray()
    n = len(x)
    J = np.zeros((n, 3, 7, 7))
    for k in range(n):
        for o, kk in enumerate((k - 1, k, k + 1)):
            if 0 <= kk < n:
                J[k, o] = A[7 * k:7 * k + 7, 7 * kk:7 * kk + 7]
    np.savez_compressed(os.path.join(HERE, "assemble_1d.npz"), x=x, u=u, un=un, F=F.reshape(n, 7), J=J)

    # 2. march parity: reference algorithm, 1 um mesh, 5 steps
    m1 = meshio.load_mesh("1D_variable_1um_mesh_1090")
    x1 = m1.x[:, 0]
    p1 = params.params_1d(L_n=1.0e-6)
    hist, its, _ = solver.march_1d(x1, p1, 5)
    np.savez_compressed(os.path.join(HERE, "march_1um.npz"), hist=hist, its=np.array(its))
    # with the H_OHP controller
    p1h = params.params_1d(L_n=1.0e-6, H_OHP=1.1)
    p1h.extras["H_OHP"] = 1.1
    hist, its, frac = solver.march_1d(x1, p1h, 5, H_OHP=1.1)
    np.savez_compressed(os.path.join(HERE, "march_1um_HOHP.npz"), last=hist[-1], its=np.array(its), frac=frac)

    # 3. config-1 mesh: march 3 steps + the steady ladder to -12.5 (Stern soft KATs, ST:66-68)
    m50 = meshio.load_mesh("1D_variable_50um_mesh_5990")
    x50 = m50.x[:, 0]
    p50 = params.params_1d()
    hist, its, _ = solver.march_1d(x50, p50, 3)
    np.savez_compressed(os.path.join(HERE, "march_50um.npz"), last=hist[-1], its=np.array(its))
    u = None
    rec = {}
    allits = []
    for V in -0.5 * np.arange(1, 26):
        u, its = solver.steady_1d(x50, p50, [V], u0=u)
        allits += its
        if V in (-1.0, -2.5, -5.0, -7.5, -10.0, -12.5):
            f, e = ohp_metrics(x50, u, p50)
            rec[f"u_{V}"] = u.copy()
            rec[f"ohp_{V}"] = np.array([f, e])
    np.savez_compressed(os.path.join(HERE, "steady_50um.npz"), its=np.array(allits), **rec)


if __name__ == "__main__":
    main()
