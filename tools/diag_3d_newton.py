"""Diagnostic: the reference's first 3D time step (from u = 0, full wall voltage) for a few voltages around the
divergence threshold of config 3, solved (a) one problem at a time with forced GMRES cluster sizes 1/2/4/8 and
(b) together in one batch: status, Newton / GMRES iteration counts, residuals."""
import os
import subprocess
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

VS = [-1.0, -1.41, -1.5, -1.6, -1.7, -2.0]


def run(Vs):
    from gmpnp_b200 import meshio, params, solver3d
    from gmpnp_b200._lib import NewtonOpts
    mesh = meshio.load_mesh("L_50_R_5")
    plist = [params.params_3d(L=50e-9, R=5e-9, voltage_multiplier=V) for V in Vs]
    pp = solver3d.PoreProblem(mesh, 50e-9, 5e-9, plist)
    s = pp.solver
    s.set_dirichlet(pp.dirichlet_values([float(p.extras["eq_scaled"][0]) for p in plist]))
    u = torch.zeros(len(Vs), s.n, 9, dtype=torch.float64, device="cuda:0")
    un = solver3d.bulk_state(len(Vs), s.n, "cuda:0")
    o = s.newton(u, un, NewtonOpts.reference_3d())
    torch.cuda.synchronize()
    for b, V in enumerate(Vs):
        print(f"   V={V}: status {int(o['status'][b])} newton {int(o['iters'][b])} gmres {int(o['lin_iters'][b])} "
              f"r0 {float(o['r0'][b]):.3e} r {float(o['r'][b]):.3e} umax {float(u[b].abs().max()):.3e}", flush=True)
    s.close()


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run([float(v) for v in sys.argv[1:]])
    else:
        for g in ("1", "8"):
            print("cluster size", g, "one problem per solve")
            for V in VS:
                subprocess.run([sys.executable, __file__, str(V)], env=dict(os.environ, GMPNP_GMRES_CLUSTER=g))
        print("all in one batch (default cluster size)")
        run(VS)
