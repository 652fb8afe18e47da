"""Diagnostic: where does the Newton increment of the CUDA 1D path stall, and is it the linear solve?
For a few sweep points: run the benchmarked continuation, then take single Newton iterations from the final state
(pivot 0 / pivot 1) and print the relative increment of each; the CPU oracle (SuperLU) takes the same iterations from
the same state for comparison."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gmpnp_b200 import meshio, params, solver1d, sweep  # noqa: E402
from gmpnp_b200._lib import NewtonOpts  # noqa: E402
from oracle import solver as osolver  # noqa: E402

PTS = [("Cs", 1.0, 200e-6, -12.5 * 200 / 256), ("K", 1.0, 200e-6, -9.521484375), ("K", 0.1, 1e-6, -12.5),
       ("K", 0.5, 5e-6, -12.5), ("Cs", 0.1, 50e-6, -6.0)]


def main():
    for cation, conc, L_n, V in PTS:
        x = meshio.load_mesh(params.mesh_name_1d(L_n)).x[:, 0]
        prm = params.params_1d(concentration_elec=conc, cation=cation, L_n=L_n, voltage_multiplier=V)
        path = sweep.voltage_paths(np.array([V]), 0.75)
        s = solver1d.Solver1D(x, batch=1)
        s.set_params([prm])
        u = solver1d.bulk_state(1, s.n, "cuda:0")
        o = NewtonOpts.steady(xtol=1e-12, xtol_path=1.0, jac_rule=1, xtol_floor=1e-6)
        o.pivot = 0
        out = s.steady(u, path, o)
        print(f"\n{cation} {conc} M L_n={L_n:g} V={V:g}: status {out['status'].tolist()} its {int(out['iters'].sum())} "
              f"dx {float(out['dx'][0]):.3e} umax {float(u.abs().max()):.3g}")
        u0 = u.clone()
        for piv in (0, 1):
            for jr in (1, 0):
                uu = u0.clone()
                seq = []
                for _ in range(5):
                    o1 = NewtonOpts.steady(xtol=1e-30, jac_rule=jr, maxit=1)
                    o1.pivot = piv
                    r = s.steady(uu, np.array([[V]]), o1)
                    seq.append(float(r["dx"][0]))
                print(f"  GPU pivot={piv} jac_rule={jr}: rel dx per extra iteration:", " ".join(f"{d:.2e}" for d in seq))
        # oracle from the same state
        n = len(x)
        cells = np.stack([np.arange(n - 1), np.arange(1, n)], axis=1)
        disc = osolver.Discretisation(x, cells, 7, jac_rule=1)
        pv = prm.with_(kappa=0.0)
        bd, bv = osolver.bc_1d(n, 7, float(V))
        uo = u0[0].cpu().numpy().ravel().copy()
        seq = []
        for _ in range(3):
            h = []
            uo, k, conv, r0, r = osolver.newton(disc, pv, uo, uo, bd, bv, point_flux=pv.jflux, criterion="increment",
                                                xtol=1e-30, maxit=1, history=h)
            seq.append(h[0][2] / max(1.0, np.abs(uo).max()))
        print("  oracle (SuperLU) from the GPU state:            ", " ".join(f"{d:.2e}" for d in seq),
              " |u_gpu - u_oracle|/|u| =", f"{np.abs(u0[0].cpu().numpy().ravel() - uo).max() / np.abs(uo).max():.2e}")
        s.close()


if __name__ == "__main__":
    main()
