"""CPU study (test infrastructure; drives the oracle): can the ONLY result values the reference holds -- the five
(field_OHP, eps_rel_OHP) pairs pasted into 1D/Stern_CO2ER.py:66-68 -- be reproduced by the oracle?

Hypothesis: they come from the default non-dry run of 1D/MPNP_CO2ER_EDL.py (0.1 M KHCO3, K+, 50 um mesh, MPNP,
10 A/m2, H2_FE 0.2, no H_OHP).  That run takes 10 000 + 10 000 backward-Euler steps (1D:271-290); the loop rebinds the
Python name ``del_t`` to the second Constant at t >= T_1 (1D:643-646), but the form F was built with the FIRST Constant
object (1D:458 ff.), so every one of the 20 000 steps is integrated with dt_1 = 1e-5 s (physical end time 0.2 s, far from
the ~3 s the boundary layer needs -- which is why the values sit between the 1 ms and the steady state).

The march follows ``oracle.solver.march_1d`` literally (u = 0 start, dolfin Newton semantics, residual criterion 1e-4);
only the sparse LU is replaced by LAPACK's banded LU (same solution to round-off, 10x faster).  OHP metrics are logged
every ``--every`` steps as JSON lines.

    OMP_NUM_THREADS=1 python tests/studies/stern_pin_study.py --V -2.5 --steps 20000 --out /tmp/stern_-2.5.jsonl
"""
import argparse
import json
import os
import sys
import time

import numpy as np
from scipy.linalg import solve_banded

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from gmpnp_b200 import meshio, params  # noqa: E402
from oracle import forms, solver  # noqa: E402

STERN = {-2.5: (-0.08032108300135771, 74.56149297894756), -5.0: (-0.2524415478848975, 57.64572780716129),
         -7.5: (-0.4612956299192668, 50.16243860179017), -10.0: (-0.6149631587776277, 49.311548142969336),
         -12.5: (-0.7310301485096051, 49.2556833480052)}
KL = 13                                    # half bandwidth of the 7x7 block-tridiagonal matrix


def ohp_metrics(x, u, prm):
    """field_OHP (1D:802-805, 893) and eps_rel_OHP (1D:895-900)."""
    g = solver.p1_gradient_projection_1d(x, u[:, 6])
    field = -g[0] * prm.thermal_voltage / prm.length * 1e-9
    w = (prm.n_water_cat * u[0, 5] * prm.c0[5] + prm.n_water_H * u[0, 0] * prm.c0[0]) * 1e-3
    return float(field), float(prm.eps_w * ((55 - w) / 55) + 6 * (w / 55))


class BandedNewton:
    def __init__(self, x, prm, jac_rule=0):
        nv = len(x)
        cells = np.stack([np.arange(nv - 1), np.arange(1, nv)], axis=1)
        self.disc = solver.Discretisation(x, cells, 7, jac_rule=jac_rule)
        self.prm = prm
        self.bd, self.bv = solver.bc_1d(nv, 7, prm.V)
        n = self.disc.ndof
        r, c = self.disc.rows, self.disc.cols
        isbc = np.zeros(n, bool)
        isbc[self.bd] = True
        self.keep = (~isbc[r]).astype(np.float64)
        self.flat = (KL + r - c) * n + c
        self.diag_bc = KL * n + self.bd
        self.n = n

    def residual(self, u, un, flux):
        return solver.apply_bc_residual(self.disc.residual(u, un, self.prm, flux), u, self.bd, self.bv)

    def solve(self, u, b):
        Je = forms.element_jacobian(self.disc.gather(u), self.disc.g, self.disc.vol, self.prm, self.disc.ruleJ)
        ab = np.bincount(self.flat, weights=Je.ravel() * self.keep, minlength=(2 * KL + 1) * self.n)
        ab[self.diag_bc] = 1.0
        return solve_banded((KL, KL), ab.reshape(2 * KL + 1, self.n), b, overwrite_ab=True, check_finite=False)

    def newton(self, u, un, flux, rtol=1e-4, atol=1e-4, maxit=50):
        x = u.copy()
        b = self.residual(x, un, flux)
        r0 = r = float(np.linalg.norm(b))
        k = 0
        conv = r < atol
        while not conv and k < maxit:
            x = x - self.solve(x, b)
            k += 1
            b = self.residual(x, un, flux)
            r = float(np.linalg.norm(b))
            conv = (r / r0 < rtol) or (r < atol)
        return x, k, conv


def march_to(V, t_end=0.2, dt=1.0e-5, dt2=4.0e-3, n1=5, grow=1.5, jac_rule=1, cation="K"):
    """Backward-Euler march of the default configuration from u = 0 to ``t_end``: ``n1`` steps of ``dt`` (the
    reference's step), then the step grows by ``grow`` per step up to ``dt2`` and lands exactly on ``t_end``.
    Returns (field_OHP, eps_rel_OHP, steps, Newton iterations)."""
    m = meshio.load_mesh("1D_variable_50um_mesh_5990")
    x = m.x[:, 0]
    nv = len(x)
    prm = params.params_1d(voltage_multiplier=V, cation=cation, time_step=dt)
    bn = BandedNewton(x, prm, jac_rule)
    u = np.zeros(bn.n)
    un = np.tile(np.array([1.0] * 6 + [0.0]), nv)
    t, dt_cur, n, tot = 0.0, dt, 0, 0
    while t_end - t > 1e-15:
        n += 1
        if n > n1:
            dt_cur = min(dt2, dt_cur * grow, t_end - t)
            bn.prm = prm = params.params_1d(voltage_multiplier=V, cation=cation, time_step=dt_cur)
        t += dt_cur
        u, k, conv = bn.newton(u, un, prm.jflux)
        if not conv:
            raise RuntimeError(f"Newton failed in step {n}")
        tot += k
        un = u.copy()
    f, e = ohp_metrics(x, u.reshape(nv, 7), prm)
    return f, e, n, tot


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--V", type=float, default=-2.5)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--every", type=int, default=250)
    ap.add_argument("--cation", default="K")
    ap.add_argument("--dt", type=float, default=1.0e-5, help="time step in s (the reference: 1e-5)")
    ap.add_argument("--dt2", type=float, default=0.0, help="feasibility mode: after --n1 steps of --dt continue with this step")
    ap.add_argument("--n1", type=int, default=100)
    ap.add_argument("--grow", type=float, default=1.25, help="feasibility mode: dt grows by this factor per step up to --dt2")
    ap.add_argument("--t_end", type=float, default=0.2, help="feasibility mode: stop exactly at this physical time")
    ap.add_argument("--jac_rule", type=int, default=0, help="1: consistent Jacobian (same step solutions, fewer iterations)")
    ap.add_argument("--out", default=None)
    ap.add_argument("--check", type=int, default=0, help="compare the first N steps with oracle.solver.march_1d")
    a = ap.parse_args()
    m = meshio.load_mesh("1D_variable_50um_mesh_5990")
    x = m.x[:, 0]
    prm = params.params_1d(voltage_multiplier=a.V, cation=a.cation, time_step=a.dt)
    bn = BandedNewton(x, prm, a.jac_rule)
    nv = len(x)
    if a.check:
        hist, its, _ = solver.march_1d(x, prm, a.check)
    u = np.zeros(bn.n)
    un = np.tile(np.array([1.0] * 6 + [0.0]), nv)
    out = open(a.out, "w") if a.out else sys.stdout
    t0 = time.time()
    tot = 0
    t_phys, dt_cur = 0.0, a.dt
    for n in range(1, a.steps + 1):
        if a.dt2 > 0 and n > a.n1:
            # feasibility mode: grow the step gently to dt2 and land exactly on t_end
            dt_cur = min(a.dt2, dt_cur * a.grow, a.t_end - t_phys)
            if dt_cur <= 1e-15:
                break
            prm = params.params_1d(voltage_multiplier=a.V, cation=a.cation, time_step=dt_cur)
            bn.prm = prm
        t_phys += dt_cur
        last = a.dt2 > 0 and abs(t_phys - a.t_end) < 1e-12
        u, k, conv = bn.newton(u, un, prm.jflux)
        if not conv:
            raise RuntimeError(f"Newton failed in step {n}")
        tot += k
        if a.check and n <= a.check:
            d = np.abs(u.reshape(nv, 7) - hist[n]).max()
            print(f"step {n}: its {k} (oracle {its[n - 1]}), max |diff| vs oracle.march_1d {d:.2e}", file=sys.stderr)
        un = u.copy()
        if n % a.every == 0 or n == a.steps or n in (1, 10, 100) or last:
            f, e = ohp_metrics(x, u.reshape(nv, 7), prm)
            gE, ge = STERN.get(a.V, (float("nan"), float("nan")))
            out.write(json.dumps({"V": a.V, "step": n, "t_s": t_phys, "field_OHP": f, "eps_rel_OHP": e,
                                  "field_rel_dev": f / gE - 1, "eps_rel_dev": e / ge - 1, "newton_total": tot,
                                  "wall_s": round(time.time() - t0, 1)}) + "\n")
            out.flush()


if __name__ == "__main__":
    main()
