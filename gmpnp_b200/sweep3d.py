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
                 max_steps: int = 40, opts: NewtonOpts | None = None, retry: bool = True, **param_kw):
        self.points = list(points)
        self.device = int(device)
        self.dv_max, self.tol, self.max_steps = dv_max, tol, max_steps
        self.opts = opts or NewtonOpts.sweep_3d()
        self.retry = bool(retry)
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
            bad = [k for k, i in enumerate(idx) if res[i, 0] != 0]
            if self.retry and bad:
                # parked points once more on their own: half the voltage increment, twice the pseudo-time steps and
                # the parity linear solver GMRES(100)/1e-10 (failure handling: never abort the batch)
                pp = PoreProblem(mesh, L, R, [plist[k] for k in bad], device=self.device)
                keep = (self.dv_max, self.max_steps, self.opts)
                self.dv_max, self.max_steps, self.opts = 0.5 * self.dv_max, 2 * self.max_steps, NewtonOpts.reference_3d()
                again = self._steady_with_parking(pp)
                self.dv_max, self.max_steps, self.opts = keep
                pp.solver.close()
                for k, row in zip(bad, again):
                    if row[0] == 0:
                        res[idx[k]] = row
        return res

    def _steady_with_parking(self, pp: PoreProblem):
        """One batch through ``gmpnp_steady_3d``: the wall voltages are ramped together, every point stops marching
        when its own increment is below ``tol``; a point whose Newton solve fails is PARKED by the library (its
        status is reported, the rest of the batch goes on).  Status 1 also marks points that did not reach ``tol``
        within ``max_steps``."""
        s = pp.solver
        out = pp.steady(opts=self.opts, tol=self.tol, max_steps=self.max_steps, dv_max=self.dv_max,
                        raise_on_failure=False)
        u = out["u"]
        status = out["status"].astype(np.int64).copy()
        status[(status == 0) & ~out["converged"]] = 1                 # max_steps reached
        med = [s.median(u, c).cpu().numpy() for c in (1, 2, 3, 7)]
        ucat = u[:, :, 7].amax(dim=1).cpu().numpy()
        res = np.zeros((s.batch, self.NCOL))
        res[:, 0], res[:, 1], res[:, 2] = status, out["steps_per_problem"], out["iters"].sum(axis=0)
        for j in range(4):
            res[:, 3 + j] = med[j]
        res[:, 7], res[:, 8] = out["co2_entry"], ucat
        return res


def shard(points, rank: int, world: int):
    """Static shard by point; a rank's points of one mesh still form one batch."""
    return points[rank::world]
