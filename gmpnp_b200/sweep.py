"""Batched parameter sweeps of the 1D steady problem (BASELINE.json config 2).

The reference runs one sweep point per process ("submit multiple jobs on a computational
cluster", README.md:37; one ``solve_EDL`` call per CLI invocation, 1D/MPNP_CO2ER_EDL.py:1105).
Here a sweep is a batch: every point is an independent steady problem with its own voltage
continuation path, all points that share a mesh go through ONE kernel launch, and the sweep
shards across GPUs by point with no data-path collective (results are gathered at the end).
"""
from __future__ import annotations

import numpy as np
import torch

from . import meshio, params as _params
from ._lib import NewtonOpts
from .solver1d import Solver1D, bulk_state, pack_1d, NC

from .sweep_points import (CONFIG2_CATIONS, CONFIG2_CONCS, CONFIG2_LN, CONFIG2_NV, CONFIG2_VMAX,  # noqa: F401
                           SweepPoint, config2_points, shard, voltage_paths)


N_SUMMARY = 10      # per-point summary row: status, Newton iterations, u(OHP)[7], projected field at the OHP


def gather_results(local: "torch.Tensor", local_index: "torch.Tensor", n_total: int, world: int):
    """The one collective of a sharded sweep: all ranks contribute their per-point summary rows
    (``local`` [n_local, k]) and the global point indices they own; every rank gets the assembled
    [n_total, k] table.  Works on the gloo (CPU tensors) and nccl (CUDA tensors) backends."""
    import torch.distributed as dist
    if world == 1 or not dist.is_initialized():
        out = torch.zeros(n_total, local.shape[1], dtype=local.dtype, device=local.device)
        out[local_index.long()] = local
        return out
    n_max = (n_total + world - 1) // world
    k = local.shape[1]
    pad = torch.full((n_max, k + 1), -1.0, dtype=torch.float64, device=local.device)
    pad[: local.shape[0], :k] = local.to(torch.float64)
    pad[: local.shape[0], k] = local_index.to(torch.float64)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    out = torch.zeros(n_total, k, dtype=torch.float64, device=local.device)
    for b in bufs:
        valid = b[:, k] >= 0
        out[b[valid, k].long()] = b[valid, :k]
    return out


class Sweep1D:
    """All sweep points of one rank, grouped by mesh into one :class:`Solver1D` each."""

    def __init__(self, points, device: int = 0, utilities_dir=None, dv_max: float = 0.5,
                 xtol: float = 1e-10, xtol_path: float = 1e-1, maxit: int = 50, jac_rule: int = 1,
                 pivot: int = 0, xtol_floor: float = 1e-8, partitions: int | None = None):
        """``pivot``: partial pivoting inside the 7x7 blocks of the block-Thomas elimination.  The sweep default is
        0: with the Poisson row equilibrated the in-block pivot is the diagonal in 99.95 % of the steps, and a
        Newton iteration that converges (increment criterion, residual evaluated independently of the linear
        solve) is correct whatever the pivoting of its linear solves; points that do NOT converge are re-run
        with pivoting and halved voltage increments by :meth:`retry_failed`.

        ``xtol`` = 1e-10: measured on a B200 (profiles/r02_floor_diagnostic.log), the Newton increments of the
        200 um-mesh problems do not contract below 4e-12 ... 2e-11 of max|u| whatever the pivoting or the Jacobian
        rule -- that is the fp64 round-off floor of the residual evaluation of these ill-conditioned problems (the
        converged states of this library and of the SuperLU oracle differ by 4e-10 there, of 1e-12 ... 1e-15 on the
        short meshes) -- so 1e-12 is not a reachable increment tolerance for 18 % of config 2.  With quadratic
        convergence an accepted increment of 1e-10 leaves an error far below that.  ``xtol_floor`` (gmpnp.h) is the
        safety net: an increment that stalls between ``xtol`` and ``xtol_floor`` ends with status 4
        (GMPNP_STAGNATED), never with 0; :meth:`summary` counts those points and reports their largest final
        relative increment."""
        self.points = list(points)
        self.device = torch.device("cuda", int(device))
        self.dv_max, self.xtol, self.xtol_path, self.maxit = dv_max, xtol, xtol_path, maxit
        self.jac_rule, self.pivot, self.xtol_floor = jac_rule, pivot, xtol_floor
        # sweeps per problem (gmpnp_newton_opts.partitions): the two-sided elimination is the throughput form; when the
        # points of this rank leave the GPU idle (strong scaling: a shard of the sweep), the partitioned elimination
        # shortens the critical path of the long problems at the price of ~2x the work
        self.partitions = partitions          # None: automatic per mesh group (see _auto_partitions); int or callable(n, batch)
        self.groups = []
        by_mesh = {}
        for i, p in enumerate(self.points):
            by_mesh.setdefault(p.L_n, []).append(i)
        pcache = {}
        # longest mesh first and, inside a mesh, longest continuation path first (LPT order: the
        # block scheduler hands out blocks in index order, so the critical-path problems start at t=0);
        # neighbours in a warp still get similar path lengths
        for L_n, idx in sorted(by_mesh.items(), key=lambda kv: -meshio.load_mesh(_params.mesh_name_1d(kv[0]), utilities_dir).num_vertices):
            idx = sorted(idx, key=lambda i: (-abs(self.points[i].V), self.points[i].conc, self.points[i].cation))
            mesh = meshio.load_mesh(_params.mesh_name_1d(L_n), utilities_dir)
            plist = []
            for i in idx:
                p = self.points[i]
                key = (p.cation, p.conc, L_n)
                if key not in pcache:
                    pcache[key] = pack_1d(_params.params_1d(concentration_elec=p.conc, cation=p.cation, L_n=L_n,
                                                            voltage_multiplier=0.0, utilities_dir=utilities_dir))
                rec = pcache[key].copy()
                rec[_params.P_V] = p.V
                plist.append(rec)
            packed = np.stack(plist)
            solver = Solver1D(mesh.x[:, 0], batch=len(idx), device=self.device.index)
            path = voltage_paths(np.array([self.points[i].V for i in idx]), dv_max)
            self.groups.append(dict(L_n=L_n, idx=np.array(idx), solver=solver, packed=packed, path=path,
                                    stream=torch.cuda.Stream(self.device)))
        self.n_points = len(self.points)
        self.extra_launches = 0       # kernels launched by sub-solvers (polish / retry), not counted by the groups

    def opts(self, pivot=None):
        o = NewtonOpts.steady(xtol=self.xtol, maxit=self.maxit, xtol_path=self.xtol_path, jac_rule=self.jac_rule,
                              xtol_floor=self.xtol_floor)
        o.pivot = self.pivot if pivot is None else pivot
        o.partitions = 2
        return o

    def _auto_partitions(self, n_nodes: int, batch: int) -> int:
        """Sweeps per problem for one mesh group.  The critical path of a launch is (Newton iterations of its longest
        path) x (rows per sweep); the partitioned elimination shortens it 2.5x (8 sweeps) at ~2x the work, which pays
        only when this rank's points leave the GPU idle.  Measured on a B200 (tools/part_time.py, config-2 shards):
        960 points 109.8 -> 61.6 ms, 1920 points 121.0 -> 102.7 ms, 3840 points 135.4 -> 199.7 ms; mixing (8 for the long
        meshes only) is slower than 8 everywhere."""
        if self.partitions is not None:
            return int(self.partitions(n_nodes, batch)) if callable(self.partitions) else int(self.partitions)
        return 8 if self.n_points <= 2400 else 2

    def group_opts(self, g, pivot=None):
        o = self.opts(pivot)
        o.partitions = self._auto_partitions(g["solver"].n, g["solver"].batch)
        return o

    # -- device-resident solve (inputs already in HBM) -----------------------------------------
    def upload(self):
        for g in self.groups:
            g["solver"].set_params(g["packed"])
            g["d_path"] = torch.as_tensor(g["path"], device=self.device)
            g["u"] = torch.empty(g["solver"].batch, g["solver"].n, NC, dtype=torch.float64, device=self.device)
            g["d_idx"] = torch.as_tensor(g["idx"], dtype=torch.long, device=self.device)
        self.d_gindex = torch.as_tensor([p.index for p in self.points], dtype=torch.long, device=self.device)

    def solve_resident(self, host_out=None):
        """One pass of the hot path over the whole batch: every point from the bulk state to its
        converged steady solution.  One launch per mesh, on concurrent streams.  ``host_out`` (one pinned host
        tensor per mesh group, shaped like the group's ``u``): the solution profiles are copied back on the
        group's own stream as soon as its launch finishes, overlapping the other meshes' compute."""
        cur = torch.cuda.current_stream(self.device)
        outs = []
        for k, g in enumerate(self.groups):
            st = g["stream"]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                u = g["u"]
                u.fill_(1.0)
                u[:, :, NC - 1] = 0.0
                outs.append(g["solver"].steady(u, g["d_path"], self.group_opts(g)))
                if host_out is not None:
                    host_out[k].copy_(u, non_blocking=True)
        for g in self.groups:
            cur.wait_stream(g["stream"])
        self.last = outs
        return outs

    def summary(self, outs):
        """Counts per status over the sweep and the largest final relative increment per status class."""
        st = np.concatenate([o["status"].cpu().numpy() for o in outs])
        dx = np.concatenate([o["dx"].cpu().numpy() for o in outs])
        res = {"converged": int((st == 0).sum()), "stagnated_at_floor": int((st == 4).sum()),
               "failed": int(((st != 0) & (st != 4)).sum())}
        res["max_final_dx_converged"] = float(dx[st == 0].max()) if (st == 0).any() else None
        res["max_final_dx_stagnated"] = float(dx[st == 4].max()) if (st == 4).any() else None
        return res

    def finish(self, outs, polish: bool = False):
        """End of a sweep pass: failed points are retried with pivoting and halved voltage increments; with
        ``polish`` the points that stalled above ``xtol`` (status 4) take a few pivoted iterations more (measured: it
        does not help, the floor is not the pivoting).  One device->host read of two counters when there is nothing
        to do."""
        st = torch.cat([o["status"] for o in outs])
        n4, nbad = torch.stack([(st == 4).sum(), ((st != 0) & (st != 4)).sum()]).tolist()
        polished = self.polish_stagnated(outs) if (n4 and polish) else 0
        retried = self.retry_failed(outs) if nbad else 0
        return dict(polished=int(polished), retried=int(retried))

    def results_device(self, outs):
        """Per-point summary rows on the device, [n_points, N_SUMMARY] in the order of ``self.points``, and the
        global sweep-point index of every row: the payload of the sharded sweep's one collective."""
        rows = torch.empty(self.n_points, N_SUMMARY, dtype=torch.float64, device=self.device)
        for g, out in zip(self.groups, outs):
            f = g["solver"].field_ohp(g["u"])
            blk = torch.cat([out["status"].to(torch.float64)[:, None],
                             out["iters"].sum(dim=1).to(torch.float64)[:, None], g["u"][:, 0, :], f[:, None]], dim=1)
            rows[g["d_idx"]] = blk
        return rows, self.d_gindex

    def polish_stagnated(self, outs):
        """Points that stalled at the round-off floor of the pivot-free elimination (status 4) take a few more Newton
        iterations with in-block pivoting from their current state, at their target voltage; their status becomes 0
        if the strict criterion is then met.  Returns the number of points polished."""
        n = 0
        for g, out in zip(self.groups, outs):
            status = out["status"].cpu().numpy()
            idx = np.nonzero(status == 4)[0]
            if not len(idx):
                continue
            n += len(idx)
            Vs = np.array([self.points[g["idx"][b]].V for b in idx])
            sub = Solver1D(g["solver"].x, batch=len(idx), device=self.device.index)
            sub.set_params(g["packed"][idx])
            sel = torch.as_tensor(idx, device=self.device)
            u = g["u"][sel].contiguous()
            o = self.opts(pivot=1)
            o.xtol_floor = 0.0
            o.maxit = 8
            res = sub.steady(u, Vs[:, None], o)
            ok = res["status"] == 0
            g["u"][sel[ok]] = u[ok]
            out["status"][sel[ok]] = 0
            out["dx"][sel[ok]] = res["dx"][ok]
            out["iters"][sel, -1] += res["iters"][:, 0]
            self.extra_launches += sub.launch_count()
            sub.close()
        return n

    def retry_failed(self, outs, max_rounds: int = 3):
        """Points whose Newton failed (status not 0 and not 4) are re-run from bulk with a halved voltage step
        (failure handling: per-problem status, never abort the batch)."""
        n_retry = 0
        for g, out in zip(self.groups, outs):
            status = out["status"].cpu().numpy()
            bad = np.nonzero((status != 0) & (status != 4))[0]
            dv = self.dv_max
            rounds = 0
            while len(bad) and rounds < max_rounds:
                dv *= 0.5
                rounds += 1
                n_retry += len(bad)
                Vs = np.array([self.points[g["idx"][b]].V for b in bad])
                sub = Solver1D(g["solver"].x, batch=len(bad), device=self.device.index)
                sub.set_params(g["packed"][bad])
                u = bulk_state(len(bad), sub.n, self.device)
                o = sub.steady(u, voltage_paths(Vs, dv), self.opts(pivot=1))
                st = o["status"].cpu().numpy()
                ok = st == 0
                g["u"][torch.as_tensor(bad[ok], device=self.device)] = u[torch.as_tensor(np.nonzero(ok)[0], device=self.device)]
                out["status"][torch.as_tensor(bad[ok], device=self.device)] = 0
                bad = bad[~ok]
                self.extra_launches += sub.launch_count()
                sub.close()
        return n_retry

    def newton_iterations(self, outs):
        """Total Newton iterations and the algorithmic HBM bytes they imply (1072 B per node per
        iteration for the fused assemble+eliminate sweep pair)."""
        its, nbytes = 0, 0
        for g, out in zip(self.groups, outs):
            k = int(out["iters"].sum().item())
            its += k
            nbytes += k * 1072 * g["solver"].n
        return its, nbytes

    def results(self, outs):
        """Gather per-point summaries on the host: status, OHP nodal values, OHP field."""
        res = np.zeros((self.n_points, 10))
        for g, out in zip(self.groups, outs):
            s = g["solver"]
            f = s.field(g["u"])[:, 0].cpu().numpy()
            u0 = g["u"][:, 0, :].cpu().numpy()
            st = out["status"].cpu().numpy()
            it = out["iters"].sum(dim=1).cpu().numpy()
            res[g["idx"], 0] = st
            res[g["idx"], 1] = it
            res[g["idx"], 2:9] = u0
            res[g["idx"], 9] = f
        return res

    def solve_checkpointed(self, directory: str, profiles: bool = False):
        """Long sweeps (the reference keeps everything in RAM and writes once at the end, so a crash loses the
        run): one mesh group at a time, each group's per-point results flushed to ``directory/group_<k>.npz`` as soon
        as it has converged (atomic rename); a re-run skips the groups whose file exists and returns the same table.
        Columns: status, Newton iterations, u(OHP)[7], projected field at the OHP; ``profiles`` also stores u."""
        import os
        os.makedirs(directory, exist_ok=True)
        if not all("u" in g for g in self.groups):
            self.upload()
        table = np.zeros((self.n_points, N_SUMMARY))
        solved = 0
        for k, g in enumerate(self.groups):
            path = os.path.join(directory, f"group_{k}.npz")
            if os.path.exists(path):
                d = np.load(path)
                if np.array_equal(d["point_index"], np.array([self.points[i].index for i in g["idx"]])):
                    table[g["idx"]] = d["summary"]
                    continue
            u = g["u"]
            u.fill_(1.0)
            u[:, :, NC - 1] = 0.0
            out = g["solver"].steady(u, g["d_path"], self.group_opts(g))
            f = g["solver"].field(u)[:, 0]
            rows = torch.cat([out["status"].to(torch.float64)[:, None], out["iters"].sum(dim=1).to(torch.float64)[:, None],
                              u[:, 0, :], f[:, None]], dim=1).cpu().numpy()
            table[g["idx"]] = rows
            data = dict(summary=rows, point_index=np.array([self.points[i].index for i in g["idx"]]), L_n=g["L_n"])
            if profiles:
                data["u"] = u.cpu().numpy()
            tmp = path + ".tmp.npz"
            np.savez_compressed(tmp, **data)
            os.replace(tmp, path)
            solved += 1
        return table, solved

    def close(self):
        for g in self.groups:
            g["solver"].close()
