"""Round-2 golden vectors of the 3D path from the CPU oracle (slow: three processes, ~15 min on 3 cores).

    python tests/golden/make_golden_3d_r02.py

* march_3d_L50R5_6.npz  : BASELINE config 3 (L_50_R_5), the reference march for 6 steps (relaxation 0.9, Sechenov median
                          update): Newton counts, CO2 entry values, states after steps 1, 2, 4, 6;
* march_3d_L100R5.npz   : the DEFAULT geometry of the reference CLI (3D:1127-1143): the wall marker's absolute r^2
                          tolerance pins 320 INTERIOR vertices to the wall potential (3D:350-356, 462; SURVEY finding 4):
                          2 steps;
* march_3d_L50R1.npz    : the degenerate L_50_R_1 (every vertex carries the wall potential, no entry/exit facets): 2 steps.
"""
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))


def run(case):
    name, L, R, n_steps, keep, out = case
    from gmpnp_b200 import marking, meshio, params
    from oracle import solver
    mesh = meshio.load_mesh(name)
    p3 = params.params_3d(L=L, R=R)
    dofs, kind, info = marking.dirichlet_sets(mesh, L, R)
    sech = lambda a, b, c, d: params.sechenov_co2_scaled(p3, a, b, c, d)
    hist, its, co2s = solver.march_3d(mesh.x, mesh.cells, p3, dofs.astype(np.int64), kind, n_steps, sechenov=sech)
    data = {f"step{k}": hist[k] for k in keep}
    np.savez_compressed(os.path.join(HERE, out), its=np.array(its), co2=np.array(co2s),
                        interior_pinned=info["interior_pinned"], phi_V_verts=info["phi_V_verts"], **data)
    return name, its, info


if __name__ == "__main__":
    cases = [("L_50_R_5", 50e-9, 5e-9, 6, (1, 2, 4, 6), "march_3d_L50R5_6.npz"),
             ("L_100_R_5", 100e-9, 5e-9, 2, (1, 2), "march_3d_L100R5.npz"),
             ("L_50_R_1", 50e-9, 1e-9, 2, (1, 2), "march_3d_L50R1.npz")]
    with mp.get_context("spawn").Pool(3) as pool:
        for r in pool.imap_unordered(run, cases):
            print(r, flush=True)
