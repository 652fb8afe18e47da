"""Per-mesh timing and iteration statistics of the 1D sweep kernel (diagnostic; also the ncu target).

    python tools/prof_1d.py [--mesh 1e-6] [--voltages 256] [--reps 2] [--pivot 1]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gmpnp_b200 import sweep  # noqa: E402
from gmpnp_b200._lib import NewtonOpts  # noqa: E402
from gmpnp_b200.solver1d import NC  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mesh", type=float, nargs="*", default=[1e-6])
ap.add_argument("--voltages", type=int, default=256)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--pivot", type=int, default=0)
ap.add_argument("--dv", type=float, default=0.75)
ap.add_argument("--xtol_path", type=float, default=1.0)
ap.add_argument("--jac_rule", type=int, default=1)
a = ap.parse_args()

pts = sweep.config2_points(a.voltages, meshes=tuple(a.mesh))
sw = sweep.Sweep1D(pts, device=0, dv_max=a.dv, xtol_path=a.xtol_path, jac_rule=a.jac_rule)      # the bench's settings
sw.upload()
opts = sw.opts(pivot=a.pivot)
for g in sw.groups:
    s = g["solver"]
    for rep in range(a.reps):
        u = g["u"]
        u.fill_(1.0)
        u[:, :, NC - 1] = 0.0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = s.steady(u, g["d_path"], opts)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    its = out["iters"].cpu().numpy()
    st = out["status"].cpu().numpy()
    tot = its.sum(1)
    nbytes = tot.sum() * 1072 * s.n
    print(f"mesh L_n={g['L_n']:g} N={s.n} batch={s.batch}: {ms:.1f} ms, its total {tot.sum()} mean {tot.mean():.1f} "
          f"max {tot.max()}, failed {int((st != 0).sum())}, alg GB/s {nbytes / ms / 1e6:.1f}, "
          f"row-iterations/s {tot.sum() * s.n / ms / 1e3:.3g}M", flush=True)
    bad = np.nonzero(st != 0)[0]
    for b in bad[:8]:
        p = sw.points[g["idx"][b]]
        print("   failed:", p.cation, p.conc, p.V, "status", st[b], "its", its[b][:30].tolist())
    # iterations per stage for the longest path
    b = int(np.argmax(tot))
    p = sw.points[g["idx"][b]]
    print("   longest:", p.cation, p.conc, p.V, its[b].tolist())
