"""Generate the committed golden vectors of the 3D path from the CPU oracle (slow: ~10 min).

    python tests/golden/make_golden_3d.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

from gmpnp_b200 import marking, meshio, params  # noqa: E402
from oracle import solver  # noqa: E402
from conftest import admissible_state, cube_tet_mesh  # noqa: E402


def main():
    # 1. kernel parity on a tiny mesh: residual + Jacobian (dense) at a random admissible state
    m = cube_tet_mesh(3)
    prm = params.params_3d(L=50e-9, R=5e-9)
    rng = np.random.default_rng(0)
    u = admissible_state(rng, m.num_vertices, 8, prm.nu, V=-3.0)
    un = admissible_state(rng, m.num_vertices, 8, prm.nu, V=-3.0)
    dofs = np.array([8, 4, 9 * 5 + 8, 9 * 63 + 8, 9 * 63 + 5], dtype=np.int64)
    vals = np.array([0.0, 2.5, -1.0, -1.0, 100.0])
    disc = solver.Discretisation(m.x, m.cells, 9)
    F = solver.apply_bc_residual(disc.residual(u.ravel(), un.ravel(), prm), u.ravel(), dofs, vals)
    A = solver.apply_bc_matrix(disc.jacobian(u.ravel(), prm), dofs).toarray()
    np.savez_compressed(os.path.join(HERE, "assemble_3d.npz"), x=m.x, cells=m.cells, u=u, un=un, dofs=dofs,
                        vals=vals, F=F.reshape(-1, 9), A=A)

    # 2. config 3: L_50_R_5, reference march (2 steps, relaxation 0.9, Sechenov median update)
    mesh = meshio.load_mesh("L_50_R_5")
    p3 = params.params_3d(L=50e-9, R=5e-9)
    dofs, kind, info = marking.dirichlet_sets(mesh, 50e-9, 5e-9)
    sech = lambda a, b, c, d: params.sechenov_co2_scaled(p3, a, b, c, d)
    hist, its, co2s = solver.march_3d(mesh.x, mesh.cells, p3, dofs.astype(np.int64), kind, 2, sechenov=sech)
    np.savez_compressed(os.path.join(HERE, "march_3d_L50R5.npz"), step1=hist[1], step2=hist[2], its=np.array(its),
                        co2=np.array(co2s))



if __name__ == "__main__":
    main()
