"""Golden vectors of the 3D reaction-diffusion drop-in (3D/rxn_diff_CO2ER_pore.py) from the INDEPENDENT 7-species CPU
oracle ``oracle/rxn_diff3d.py`` (~30 s): two reference time steps on L_10_R_5 incl. the Sechenov update.

    python tests/golden/make_golden_rd3.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from gmpnp_b200 import marking, meshio, params, rxn_diff3d  # noqa: E402
from oracle import rxn_diff3d as oracle_rd3  # noqa: E402

MESH, L, R = "L_10_R_5", 10e-9, 5e-9


def main():
    mesh = meshio.load_mesh(MESH)
    p = rxn_diff3d.params_rxn_diff_3d(L=L, R=R)
    ww, ef, ea = marking.facet_terms(mesh, L, R)
    facets, _, marker = marking.mark_facets(mesh, R / L, marking.wall_tolerance(L, R))
    ev = np.unique(facets[marker == 1])
    sech = lambda a, b, c, d: params.sechenov_co2_scaled(p, a, b, c, d)
    hist, its, co2s = oracle_rd3.march(mesh.x, mesh.cells, p, ev, ww, ef, ea, 2, sechenov=sech)
    np.savez_compressed(os.path.join(HERE, "rxn_diff_3d_L10R5.npz"), steps=hist[1:], its=np.array(its),
                        co2_entry=np.array(co2s, dtype=np.float64))
    print("its", its, "co2", co2s)


if __name__ == "__main__":
    main()
