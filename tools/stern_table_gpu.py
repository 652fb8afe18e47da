"""GPU check against the reference's own result table: march the default 1D configuration (0.1 M KHCO3, K+,
1D_variable_50um_mesh_5990, MPNP, 10 A/m2, H2_FE 0.2) at the five wall voltages of 1D/Stern_CO2ER.py:66-68 from u = 0
to t = 0.2 s -- the 20 000 steps of 1e-5 s that the non-dry run of 1D/MPNP_CO2ER_EDL.py integrates as executed
(``edl1d`` docstring) -- with the CUDA march (``gmpnp_march_1d``) and compare field_OHP / eps_rel_OHP with the table.
The CPU oracle reproduces the table to 9-10 digits; this is the same check for the product path.

    python tools/stern_table_gpu.py                 # two coarse marches (2e-4 s, 1e-4 s) + extrapolation to 1e-5 s: ~1 min
    python tools/stern_table_gpu.py --exact         # the 20 000 steps themselves (consistent Jacobian: ~6 min on a B200)
    python tools/stern_table_gpu.py --exact --jac_rule 0     # ... with FFC's rule pair, i.e. the reference's iteration path
                                                             # (measured: at -12.5 V_T that Newton iteration reaches
                                                             # maxit = 50 at some step, so that voltage stops early)

Product API only (no oracle); one JSON line per mode.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gmpnp_b200 import meshio, params, solver1d  # noqa: E402
from gmpnp_b200._lib import NewtonOpts  # noqa: E402
from gmpnp_b200.stern import OHP_DICT  # noqa: E402  (the table, ST:66-68)

VS = list(OHP_DICT.keys())


def plist(dt):
    return [params.params_1d(voltage_multiplier=V, time_step=dt) for V in VS]


def ohp(s, u, prm):
    f = s.field(u)[:, 0].cpu().numpy() * prm.thermal_voltage / prm.length * 1.0e-9          # 1D:802-805
    U = u[:, 0, :].cpu().numpy()
    w = (prm.n_water_cat * U[:, 5] * prm.c0[5] + prm.n_water_H * U[:, 0] * prm.c0[0]) * 1.0e-3
    return f, prm.eps_w * ((55 - w) / 55) + 6 * (w / 55)                                      # 1D:895-900


def march_to(s, opts, t_end=0.2, dt=1.0e-5, dt2=0.0, n1=100, grow=1.25):
    dev = s.device
    u = torch.zeros(s.batch, s.n, 7, dtype=torch.float64, device=dev)
    un = solver1d.bulk_state(s.batch, s.n, dev)
    its = 0
    if dt2 <= 0:
        n = int(round(t_end / dt))
        s.set_params(plist(dt))
        out = s.march(u, un, n, opts)
        # a voltage whose Newton iteration hits maxit at some step (FFC's rule pair at -12.5 V_T) stops alone; the
        # others are reported (status per voltage in the JSON line)
        march_to.status = out["status"].tolist()
        return u, n, int(out["iters"].sum(dim=1).max())
    s.set_params(plist(dt))
    out = s.march(u, un, n1, opts)
    assert not out["status"].any().item(), out["status"].tolist()
    its += int(out["iters"].sum(dim=1).max())
    t, cur, n = n1 * dt, dt, n1
    while t_end - t > 1e-15:
        cur = min(dt2, cur * grow, t_end - t)
        s.set_params(plist(cur))
        out = s.march(u, un, 1, opts)
        assert not out["status"].any().item(), (n, out["status"].tolist())
        its += int(out["iters"].sum(dim=1).max())
        t += cur
        n += 1
    return u, n, its


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--exact", action="store_true")
    ap.add_argument("--jac_rule", type=int, default=1)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args()
    mesh = meshio.load_mesh("1D_variable_50um_mesh_5990")
    s = solver1d.Solver1D(mesh.x[:, 0], batch=len(VS), device=a.device)
    opts = NewtonOpts.reference_1d()
    opts.jac_rule = a.jac_rule
    prm = plist(1.0e-5)[0]
    gold_E = np.array([OHP_DICT[V]["E"] for V in VS])
    gold_eps = np.array([OHP_DICT[V]["eps"] for V in VS])
    t0 = time.time()
    if a.exact:
        u, n, its = march_to(s, opts)
        f, e = ohp(s, u, prm)
        line = {"mode": "exact", "steps": n, "newton_max": its, "status_per_voltage": march_to.status}
    else:
        ua, na, ia = march_to(s, opts, dt2=2.0e-4)
        fa, ea = ohp(s, ua, prm)
        ub, nb, ib = march_to(s, opts, dt2=1.0e-4)
        fb, eb = ohp(s, ub, prm)
        f, e = fb - 0.9 * (fa - fb), eb - 0.9 * (ea - eb)                 # first order in dt: value at dt = 1e-5
        line = {"mode": "extrapolated from dt = 2e-4, 1e-4", "steps": [na, nb], "newton_max": [ia, ib],
                "field_rel_dev_dt2e-4": (fa / gold_E - 1).tolist(), "field_rel_dev_dt1e-4": (fb / gold_E - 1).tolist()}
    torch.cuda.synchronize()
    line.update({"V": VS, "jac_rule": a.jac_rule, "field_OHP": f.tolist(), "eps_rel_OHP": e.tolist(),
                 "field_rel_dev": (f / gold_E - 1).tolist(), "eps_rel_dev": (e / gold_eps - 1).tolist(),
                 "wall_s": round(time.time() - t0, 1)})
    print(json.dumps(line))
    s.close()


if __name__ == "__main__":
    main()
