"""Steady 3D solves per second for a batch of config-3 problems (the `pore3d` part of bench.py alone; diagnostic).

    python tools/steady_time.py --batch 128 [--inexact]        (GMPNP_LIB=... selects an experimental build)
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gmpnp_b200 import meshio, params, solver3d  # noqa: E402
from gmpnp_b200._lib import NewtonOpts  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, nargs="*", default=[128])
ap.add_argument("--inexact", action="store_true")
ap.add_argument("--eta", type=float, nargs="*", default=[], help="forcing terms to time; distance to the 1e-10 iterate is printed")
a = ap.parse_args()
mesh = meshio.load_mesh("L_50_R_5")
for batch in a.batch:
    Vs = np.linspace(-0.5, -1.25, batch)
    plist = [params.params_3d(L=50e-9, R=5e-9, voltage_multiplier=float(V)) for V in Vs]
    pp = solver3d.PoreProblem(mesh, 50e-9, 5e-9, plist)
    settings = [("GMRES(40)/1e-8", NewtonOpts.sweep_3d())]
    if a.inexact:
        settings.append(("eta=1e-4", NewtonOpts.sweep_3d_inexact()))
    ref = None
    if a.eta:
        o = NewtonOpts.sweep_3d(); o.lin_rtol = 1e-10
        settings.insert(0, ("GMRES(40)/1e-10", o))
        settings += [(f"eta={e:g}", NewtonOpts.sweep_3d_inexact(e)) for e in a.eta]
    for name, opts in settings:
        pp.steady(opts=opts, tol=1e-8, max_steps=2, raise_on_failure=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = pp.steady(opts=opts, tol=1e-8, max_steps=20, raise_on_failure=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        n_conv = int(np.sum(out["converged"]))
        dist = ""
        if a.eta:
            if ref is None:
                ref = out["u"].clone()
            else:
                d = (out["u"] - ref).abs().amax(dim=(1, 2)) / ref.abs().amax(dim=(1, 2))
                l2 = ((out["u"] - ref).pow(2).sum(dim=1).sqrt() / ref.pow(2).sum(dim=1).sqrt().clamp_min(1e-300)).amax()
                dist = f", max-norm distance to the 1e-10 iterate {float(d.max()):.2e}, worst per-field rel L2 {float(l2):.2e}"
        print(f"batch {batch} {name}: {ms:.1f} ms, converged {n_conv}/{batch}, {n_conv / ms * 1e3:.1f} steady solves/s, "
              f"steps {int(out['steps'])}, Newton its {int(out['iters'].sum(axis=0).max())}{dist}", flush=True)
    pp.solver.close()
    del pp
    torch.cuda.empty_cache()
