"""Golden vectors of the 3D path WITH the intended boundary integrals (3D:474-499), from the CPU oracle (~5 min).

    python tests/golden/make_golden_3d_intended.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from gmpnp_b200 import marking, meshio, params  # noqa: E402
from oracle import solver  # noqa: E402


def main():
    mesh = meshio.load_mesh("L_50_R_5")
    p3 = params.params_3d(L=50e-9, R=5e-9)
    dofs, kind, _ = marking.dirichlet_sets(mesh, 50e-9, 5e-9)
    ww, ef, ea = marking.facet_terms(mesh, 50e-9, 5e-9)
    ft = (ww, ef, ea, p3.extras["J_wall"], p3.extras["k_exit"])
    # residual and J x at a random state (kernel parity), then ONE reference time step
    rng = np.random.default_rng(21)
    nv = mesh.x.shape[0]
    u = np.ones((nv, 9)); u[:, 8] = 0.0
    u += 0.05 * rng.random((nv, 9))
    un = np.ones((nv, 9)); un[:, 8] = 0.0
    x = rng.normal(size=(nv, 9))
    disc = solver.Discretisation(mesh.x, mesh.cells, 9, facet_terms=ft)
    eq = p3.extras["eq_scaled"]
    vals = np.array([0.0, p3.V, float(eq[0]), eq[1], eq[2]])[kind.astype(np.int64)]
    F = solver.apply_bc_residual(disc.residual(u.ravel(), un.ravel(), p3), u.ravel(), dofs.astype(np.int64), vals)
    A = solver.apply_bc_matrix(disc.jacobian(u.ravel(), p3), dofs.astype(np.int64))
    Jx = A @ x.ravel()
    sech = lambda a, b, c, d: params.sechenov_co2_scaled(p3, a, b, c, d)
    hist, its, co2s = solver.march_3d(mesh.x, mesh.cells, p3, dofs.astype(np.int64), kind, 1, sechenov=sech, facet_terms=ft)
    np.savez_compressed(os.path.join(HERE, "intended_3d_L50R5.npz"), u=u, x=x, F=F.reshape(nv, 9), Jx=Jx.reshape(nv, 9),
                        step1=hist[1], its=np.array(its))
    print("its", its)


if __name__ == "__main__":
    main()
