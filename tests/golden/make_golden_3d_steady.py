"""Golden vector of the 3D steady solve with voltage continuation (north_star; N1 of the scope table) from the CPU
oracle: L_10_R_5, wall voltage -2 V_T ramped over 4 pseudo-time steps (dv_max = 0.5 V_T), marched to increments
<= 1e-8 with the reference's Newton settings (relaxation 0.9, residual criterion 1e-4) and the Sechenov median feedback.

    python tests/golden/make_golden_3d_steady.py        # ~10 min on one core
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from gmpnp_b200 import marking, meshio, params  # noqa: E402
from oracle import solver  # noqa: E402


def main():
    L, R, V = 10e-9, 5e-9, -2.0
    mesh = meshio.load_mesh("L_10_R_5")
    p3 = params.params_3d(L=L, R=R, voltage_multiplier=V)
    dofs, kind, info = marking.dirichlet_sets(mesh, L, R)
    sech = lambda a, b, c, d: params.sechenov_co2_scaled(p3, a, b, c, d)
    t = time.time()
    u, its, incs, co2 = solver.steady_march_3d(mesh.x, mesh.cells, p3, dofs.astype(np.int64), kind, tol=1e-8,
                                               max_steps=40, n_ramp=4, sechenov=sech)
    print("steps", len(its), "newton", its, "increments", incs, "co2", co2, round(time.time() - t, 1), "s")
    np.savez_compressed(os.path.join(HERE, "steady_3d_L10R5.npz"), u=u, its=np.array(its), incs=np.array(incs), co2=co2,
                        V=V, n_ramp=4)


if __name__ == "__main__":
    main()
