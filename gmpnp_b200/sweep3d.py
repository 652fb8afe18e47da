"""Batched 3D pore geometry x voltage sweeps (BASELINE.json config 4).

The reference runs one (geometry, voltage) point per process (``solveEDL`` per CLI call, 3D/MPNP_CO2ER_pore.py:1237).
Here all voltage points of one pore mesh form one batch (they share the mesh, the BSR pattern and the gather
lists), every batch is marched to its steady state with the voltage ramp of ``PoreProblem.steady``, and a sweep
shards across GPUs by (mesh, voltage) point with no data-path collective -- the per-point summaries are gathered
once at the end (``sweep.gather_results``).

Failure handling: a point whose damped Newton iteration fails (the discrete problem breaks down beyond
|V| ~ 3 V_T on the reference meshes) is PARKED -- its wall voltage is set to 0 and its state reset
to the bulk state, so it costs almost nothing in the remaining steps -- and reported with its status; it never
aborts the batch.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import meshio, params as _params
from ._lib import NewtonOpts
from .solver3d import PoreProblem, bulk_state

# the 11 pore meshes present in utilities/ (SURVEY App. E): (file stem, L [m], R [m]).  L_50_R_2.5 is byte-identical
# to L_100_R_5 and L_50_R_7.5 is unreachable from the reference CLI (int() truncation), both are still listed.
CONFIG4_MESHES = (("L_100_R_5", 100e-9, 5e-9), ("L_10_R_5", 10e-9, 5e-9), ("L_25_R_5", 25e-9, 5e-9),
                  ("L_50_R_1", 50e-9, 1e-9), ("L_50_R_2", 50e-9, 2e-9), ("L_50_R_2.5", 50e-9, 2.5e-9),
                  ("L_50_R_4", 50e-9, 4e-9), ("L_50_R_5", 50e-9, 5e-9), ("L_50_R_7.5", 50e-9, 7.5e-9),
                  ("L_50_R_10", 50e-9, 10e-9), ("L_80_R_5", 80e-9, 5e-9))


@dataclass
class PorePoint:
    mesh: str
    L: float
    R: float
    V: float
    index: int = 0


def config4_points(n_voltages: int = 256, vmax: float = -12.5, meshes=CONFIG4_MESHES):
    """All present pore meshes x V_k = vmax (k+1)/n (SURVEY 8d cfg 4)."""
    pts = []
    for name, L, R in meshes:
        for k in range(n_voltages):
            pts.append(PorePoint(name, L, R, vmax * (k + 1) / n_voltages, len(pts)))
    return pts


class Sweep3D:
    """The points of one rank, grouped by mesh; ``solve`` returns one summary row per point:
    [status, pseudo-time steps, Newton iterations, median OH, median HCO3, median CO32, median cation,
    CO2 entry value, max cation]."""

    NCOL = 9

    def __init__(self, points, device: int = 0, utilities_dir=None, dv_max: float = 0.5, tol: float = 1e-8,
                 max_steps: int = 40, opts: NewtonOpts | None = None, **param_kw):
        self.points = list(points)
        self.device = int(device)
        self.dv_max, self.tol, self.max_steps = dv_max, tol, max_steps
        self.opts = opts or NewtonOpts.sweep_3d()
        self.utilities_dir, self.param_kw = utilities_dir, param_kw
        self.by_mesh = {}
        for i, p in enumerate(self.points):
            self.by_mesh.setdefault((p.mesh, p.L, p.R), []).append(i)

    def solve(self):
        res = np.zeros((len(self.points), self.NCOL))
        for (name, L, R), idx in self.by_mesh.items():
            mesh = meshio.load_mesh(name, self.utilities_dir)
            plist = [_params.params_3d(L=L, R=R, voltage_multiplier=self.points[i].V, utilities_dir=self.utilities_dir,
                                       **self.param_kw) for i in idx]
            pp = PoreProblem(mesh, L, R, plist, device=self.device)
            res[idx] = self._steady_with_parking(pp)
            pp.solver.close()
        return res

    def _steady_with_parking(self, pp: PoreProblem):
        s = pp.solver
        B = s.batch
        dev = pp.device
        un = bulk_state(B, s.n, dev)
        u = un.clone()
        co2 = [float(p.extras["eq_scaled"][0]) for p in pp.plist]
        packed = np.stack([p.pack() for p in pp.plist])
        Vt = np.array([p.V for p in pp.plist], dtype=np.float64)
        n_ramp = max(1, int(np.ceil(np.abs(Vt).max() / self.dv_max - 1e-12)))
        status = np.zeros(B, dtype=np.int64)
        its = np.zeros(B, dtype=np.int64)
        steps = np.zeros(B, dtype=np.int64)
        parked = np.zeros(B, dtype=bool)
        done = np.zeros(B, dtype=bool)
        bulk = bulk_state(1, s.n, dev)[0]
        for step in range(self.max_steps):
            Vk = np.where(parked, 0.0, Vt * min(1.0, (step + 1) / n_ramp))
            packed[:, _params.P_V] = Vk
            s.set_params(packed)
            s.set_dirichlet(pp.dirichlet_values(co2, V=Vk))
            out = s.newton(u, un, self.opts)
            st = out["status"].cpu().numpy()
            k = out["iters"].cpu().numpy()
            newly = (st != 0) & ~parked
            for b in np.nonzero(newly)[0]:
                status[b] = st[b]
                parked[b] = True
                u[b].copy_(bulk)
                un[b].copy_(bulk)
                co2[b] = float(pp.plist[b].extras["eq_scaled"][0])
            live = ~parked & ~done
            its[live] += k[live]
            steps[live] += 1
            med = [s.median(u, c).cpu().numpy() for c in (1, 2, 3, 7)]
            for b in np.nonzero(~parked)[0]:
                co2[b] = _params.sechenov_co2_scaled(pp.plist[b], med[0][b], med[1][b], med[2][b], med[3][b])
            inc = ((u - un).abs().amax(dim=(1, 2)) / u.abs().amax(dim=(1, 2)).clamp(min=1.0)).cpu().numpy()
            un.copy_(u)
            if step + 1 >= n_ramp:
                done |= (inc <= self.tol) & ~parked
            if (done | parked).all():
                break
        status[~parked & ~done] = 1                                  # max_steps reached
        med = [s.median(u, c).cpu().numpy() for c in (1, 2, 3, 7)]
        ucat = u[:, :, 7].amax(dim=1).cpu().numpy()
        out = np.zeros((B, self.NCOL))
        out[:, 0], out[:, 1], out[:, 2] = status, steps, its
        for j in range(4):
            out[:, 3 + j] = med[j]
        out[:, 7], out[:, 8] = np.array(co2), ucat
        return out


def shard(points, rank: int, world: int):
    """Static shard by point; a rank's points of one mesh still form one batch."""
    return points[rank::world]
