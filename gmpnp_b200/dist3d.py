"""Mesh-partitioned 3D pore mode: one problem spans the GPUs of a box (BASELINE config 5).

What replaces what: the reference solves every Newton step of 3D/MPNP_CO2ER_pore.py:789-799 with a serial MUMPS
LU (3D:792) and has no distributed path at all.  Here the refined pore mesh is cut into z-slabs
(:mod:`gmpnp_b200.partition`), every rank assembles the block rows of the vertices it owns from its local tets
(no exchange of matrix entries), and the linear solve is right-preconditioned restarted GMRES with

* BSR SpMV on the local rows after a HALO EXCHANGE of the ghost vertices' 9-vectors
  (``torch.distributed`` point-to-point over NCCL/NVLink; interface size ~ a pore cross-section),
* a per-node 9x9 block-Jacobi preconditioner (communication-free) plus the additive z-slab coarse correction of
  the single-mesh path (16 slabs x 9 components): its 144 x 144 Galerkin matrix is summed over the ranks once per
  Newton iteration, its restricted residual (144 doubles) once per application,
* CGS2 orthogonalisation with ONE all-reduce per Gram-Schmidt pass (the norm of the new direction is fused
  into the second pass), i.e. two small all-reduces per iteration; the Hessenberg/Givens recurrences run
  redundantly on every rank's host from the reduced dot products.

All per-rank arithmetic goes through the C-ABI (include/gmpnp.h: assemble / spmv / bjacobi / vec_multi_dot /
vec_lincomb); PyTorch provides buffers and the collectives.  The same driver runs several parts inside ONE
process (:class:`LocalComm`) -- that is how the path is tested on a single GPU and how ranks are emulated
without launching kernels that wait on one another.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import marking, partition as _part, params as _params
from ._lib import check, ptr
from .solver3d import NC, Solver3D


# ------------------------------------------------------------------------------------------------
# communication
# ------------------------------------------------------------------------------------------------
class LocalComm:
    """All parts live in this process: the halo exchange is a direct copy, the all-reduce a local sum."""

    def __init__(self, parts):
        self.parts = list(parts)
        self.world = self.parts[0].world
        assert sorted(p.rank for p in self.parts) == list(range(self.world)), "LocalComm needs every part"
        self.halo_bytes = 0

    def halo(self, xs):
        """xs[k]: [n_local_k, ncomp] tensor of part k; fills the ghost rows from their owners."""
        by_rank = {p.rank: (p, x) for p, x in zip(self.parts, xs)}
        for p, x in zip(self.parts, xs):
            for nbr, ridx in p.recv.items():
                q, xq = by_rank[nbr]
                sidx = q.send[p.rank]
                x[torch.as_tensor(ridx, device=x.device)] = xq[torch.as_tensor(sidx, device=x.device)]
                self.halo_bytes += len(ridx) * x.shape[-1] * 8

    def allreduce_sum(self, ts):
        """ts[k]: this part's partial sums (same shape for all parts) -> the total (one tensor)."""
        out = ts[0].clone()
        for t in ts[1:]:
            out += t
        return out

    def gather_values(self, vs):
        """vs[k]: 1-D tensor of part k's OWNED values -> all values of the mesh (any order), on every part."""
        return torch.cat([v.reshape(-1) for v in vs])


class TorchComm:
    """One part per process: halo exchange = batched isend/irecv, reductions = all_reduce
    (NCCL for CUDA tensors, gloo for the CPU tests)."""

    def __init__(self, part, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.parts = [part]
        self.world = part.world
        self.group = group
        self.halo_bytes = 0
        self._idx = {}

    def _index(self, arr, device):
        key = (id(arr), str(device))
        if key not in self._idx:
            self._idx[key] = torch.as_tensor(np.asarray(arr), dtype=torch.int64, device=device)
        return self._idx[key]

    def halo(self, xs):
        dist = self.dist
        (p,), (x,) = self.parts, xs
        nvtx = x.is_cuda
        if nvtx:
            torch.cuda.nvtx.range_push("gmpnp:halo_exchange")
        ops, recvs = [], []
        for nbr in sorted(set(p.send) | set(p.recv)):
            if nbr in p.send:
                sb = x[self._index(p.send[nbr], x.device)].contiguous()
                ops.append(dist.P2POp(dist.isend, sb, nbr, group=self.group))
            if nbr in p.recv:
                rb = torch.empty(len(p.recv[nbr]), x.shape[-1], dtype=x.dtype, device=x.device)
                recvs.append((nbr, rb))
                ops.append(dist.P2POp(dist.irecv, rb, nbr, group=self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for nbr, rb in recvs:
            x[self._index(p.recv[nbr], x.device)] = rb
            self.halo_bytes += rb.numel() * 8
        if nvtx:
            torch.cuda.nvtx.range_pop()

    def allreduce_sum(self, ts):
        t = ts[0].clone()
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def gather_values(self, vs):
        """All-gather of the ranks' owned values (ragged: sizes first, then padded buffers)."""
        v = vs[0].reshape(-1).contiguous()
        if self.world == 1:
            return v
        dist = self.dist
        n = torch.tensor([v.numel()], dtype=torch.int64, device=v.device)
        sizes = [torch.zeros_like(n) for _ in range(self.world)]
        dist.all_gather(sizes, n, group=self.group)
        sizes = [int(t.item()) for t in sizes]
        buf = torch.zeros(max(sizes), dtype=v.dtype, device=v.device)
        buf[: v.numel()] = v
        outs = [torch.empty_like(buf) for _ in range(self.world)]
        dist.all_gather(outs, buf, group=self.group)
        return torch.cat([o[:k] for o, k in zip(outs, sizes)])


# ------------------------------------------------------------------------------------------------
# the partitioned problem
# ------------------------------------------------------------------------------------------------
class PartitionedPore:
    """One pore problem (one parameter point) on a partitioned mesh.  ``parts`` are the parts handled by THIS
    process (all of them with :class:`LocalComm`, exactly one with :class:`TorchComm`)."""

    NZ = 16          # z-slabs of the coarse space (NZ in csrc/pore3d.cu)

    def __init__(self, mesh, L: float, R: float, prm, parts, comm, device: int = 0, dirichlet=None, coarse: bool = True,
                 intended_bcs: bool = False):
        """``intended_bcs``: the wall-flux / pore-exit Robin integrals of 3D:474-499 (see ``PoreProblem``).  The
        lumped wall weights are per-vertex data and are simply restricted to a part; an exit facet is given to every
        part that owns at least one of its vertices (its tet is local there, so all three vertices and the blocks
        it touches exist locally; the owned rows receive every facet that touches them, ghost rows are unused)."""
        self.mesh, self.L, self.R, self.prm = mesh, L, R, prm
        self.parts, self.comm = list(parts), comm
        self.device = torch.device("cuda", int(device))
        self.dofs, self.kind, self.info = dirichlet if dirichlet is not None else marking.dirichlet_sets(mesh, L, R)
        self.solvers, self.dir_sel = [], []
        self.coarse = bool(coarse)
        self.intended_bcs = bool(intended_bcs)
        if self.intended_bcs:
            wall_w, exit_f, exit_a = marking.facet_terms(mesh, L, R)
        z = np.asarray(mesh.x)[:, 2]
        zmin, zmax = float(z.min()), float(z.max())
        for p in self.parts:
            ld, sel = _part.local_dirichlet(p, self.dofs)
            s = Solver3D(p.local_mesh(), ld, batch=1, device=device)
            s.set_params([prm])
            # GLOBAL z-slab ids of the local vertices (the handle's own default is relative to the local slab)
            agg = np.clip(((p.x[:, 2] - zmin) / (zmax - zmin if zmax > zmin else 1.0) * self.NZ).astype(np.int32),
                          0, self.NZ - 1)
            agg = np.ascontiguousarray(agg, dtype=np.int32)
            check(s.lib.gmpnp_set_aggregates_3d(s._h, agg.ctypes.data_as(C.POINTER(C.c_int))), s._h)
            if self.intended_bcs:
                lf, sel_f = _part.local_facets(p, exit_f, mesh.x.shape[0])
                s.set_facet_terms(wall_w[p.glob], lf, np.asarray(exit_a)[sel_f],
                                  prm.extras["J_wall"][None, :], prm.extras["k_exit"][None, :])
            self.solvers.append(s)
            self.dir_sel.append(sel)
        self.lib = self.solvers[0].lib
        self.n_own = [p.n_own for p in self.parts]
        self.J = [None] * len(self.parts)
        self.stats = dict(spmv=0, halo=0, allreduce=0, gmres_iters=0)
        self.overlap, self.overlap_min_rows = None, 1 << 40
        self.set_dirichlet(float(prm.extras["eq_scaled"][0]))

    # -- helpers ----------------------------------------------------------------------------
    def set_dirichlet(self, co2_scaled: float, V=None):
        eq = self.prm.extras["eq_scaled"]
        vals = marking.dirichlet_values(self.kind, self.prm.V if V is None else V, co2_scaled, eq[1], eq[2])
        for s, sel in zip(self.solvers, self.dir_sel):
            s.set_dirichlet(vals[sel][None, :])

    def zeros(self):
        return [torch.zeros(1, p.n_local, NC, dtype=torch.float64, device=self.device) for p in self.parts]

    def from_global(self, xg: np.ndarray):
        return [torch.as_tensor(_part.scatter_to_part(p, xg)[None], dtype=torch.float64, device=self.device).contiguous()
                for p in self.parts]

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def dot_owned(self, Vs, nvec: int, ws, extra_self: bool = False):
        """Reduced dot products [V_k . w for k < nvec] (+ w . w if ``extra_self``) over the owned rows of all
        parts: two-stage device reductions per part, ONE all-reduce."""
        outs = []
        for s, p, V, w in zip(self.solvers, self.parts, Vs, ws):
            n = p.n_own * NC
            out = torch.empty(nvec + (1 if extra_self else 0), dtype=torch.float64, device=self.device)
            if nvec:
                check(self.lib.gmpnp_vec_multi_dot(s._h, ptr(V), V.stride(0), nvec, ptr(w), n, ptr(out), self._stream()), s._h)
            if extra_self:
                check(self.lib.gmpnp_vec_multi_dot(s._h, ptr(w), w.numel(), 1, ptr(w), n,
                                                   C.c_void_p(out.data_ptr() + 8 * nvec), self._stream()), s._h)
            outs.append(out)
        self.stats["allreduce"] += 1
        return self.comm.allreduce_sum(outs)

    def lincomb(self, Vs, nvec: int, coef: torch.Tensor, outs, beta: float = 0.0, ys=None):
        """outs[k] = beta * ys[k] + sum_j coef[j] V_j on the owned rows."""
        for k, (s, p, V, o) in enumerate(zip(self.solvers, self.parts, Vs, outs)):
            y = None if ys is None else ys[k]
            check(self.lib.gmpnp_vec_lincomb(s._h, ptr(V), V.stride(0), nvec, ptr(coef), float(beta), ptr(y), ptr(o),
                                             p.n_own * NC, self._stream()), s._h)

    # -- operators ---------------------------------------------------------------------------
    def assemble(self, us, uns, want_J=True):
        """Halo-exchange u, assemble every part; returns (F list, ||F||_2 over owned rows)."""
        self.comm.halo([u[0] for u in us])
        self.stats["halo"] += 1
        Fs = []
        for k, (s, u, un) in enumerate(zip(self.solvers, us, uns)):
            F, J = s.assemble(u, un, want_F=True, want_J=want_J)
            if want_J:
                self.J[k] = J
                check(self.lib.gmpnp_bjacobi_setup_3d(s._h, ptr(J), self._stream()), s._h)
            Fs.append(F)
        if want_J and self.coarse:
            nco = self.NZ * NC
            Acs = []
            for s, p, J in zip(self.solvers, self.parts, self.J):
                Ac = torch.empty(nco * nco, dtype=torch.float64, device=self.device)
                check(self.lib.gmpnp_coarse_accumulate_3d(s._h, ptr(J), p.n_own, ptr(Ac), self._stream()), s._h)
                Acs.append(Ac)
            Ac = self.comm.allreduce_sum(Acs)
            self.stats["allreduce"] += 1
            for s in self.solvers:
                check(self.lib.gmpnp_coarse_invert_3d(s._h, ptr(Ac), self._stream()), s._h)
        nrm2 = self.dot_owned([F.view(1, -1) for F in Fs], 1, [F.view(-1) for F in Fs])
        return Fs, math.sqrt(float(nrm2[0]))

    def spmv(self, xs, overlap=None):
        """y = J x on the owned rows (ghost rows of y are not computed).  With ``overlap`` the interior rows (no ghost
        column) are multiplied on a side stream while the halo exchange of x is in flight and the boundary rows
        follow it.  Measured on 2 B200s over NVLink (x3 refinement, 342 k owned rows, 480 KB halo per rank): the
        exchange costs ~40 us next to a 780 us SpMV, and the two extra launches + stream hand-offs of the overlapped
        form cost ~120 us, so the default is OFF (``overlap_min_rows`` can enable it for parts whose interface is a
        larger fraction of the work)."""
        if overlap is None:
            overlap = self.overlap if self.overlap is not None else min(self.n_own) >= self.overlap_min_rows
        self.stats["halo"] += 1
        self.stats["spmv"] += 1
        main = torch.cuda.current_stream(self.device)
        out = [torch.empty_like(x) for x in xs]
        if overlap:
            if not hasattr(self, "_side"):
                self._side = torch.cuda.Stream(self.device)
            side = self._side
            side.wait_stream(main)
            with torch.cuda.stream(side):
                for s, p, J, x, y in zip(self.solvers, self.parts, self.J, xs, out):
                    y.record_stream(side)
                    check(self.lib.gmpnp_spmv_rows_3d(s._h, ptr(J), ptr(x), ptr(y), 0, p.n_int,
                                                      C.c_void_p(side.cuda_stream)), s._h)
        self.comm.halo([x[0] for x in xs])
        for s, p, J, x, y in zip(self.solvers, self.parts, self.J, xs, out):
            check(self.lib.gmpnp_spmv_rows_3d(s._h, ptr(J), ptr(x), ptr(y), p.n_int if overlap else 0, p.n_own,
                                              self._stream()), s._h)
        if overlap:
            main.wait_stream(side)
        return out

    def precond(self, rs, zs):
        """z = D^-1 r + P A_c^-1 P^T r on the owned rows (one 144-double all-reduce when the coarse space is on)."""
        for s, p, r, z in zip(self.solvers, self.parts, rs, zs):
            check(self.lib.gmpnp_bjacobi_apply_3d(s._h, ptr(r), ptr(z), p.n_own, self._stream()), s._h)
        if self.coarse:
            rcs = []
            for s, p, r in zip(self.solvers, self.parts, rs):
                rc = torch.empty(self.NZ * NC, dtype=torch.float64, device=self.device)
                check(self.lib.gmpnp_coarse_restrict_3d(s._h, ptr(r), p.n_own, ptr(rc), self._stream()), s._h)
                rcs.append(rc)
            rc = self.comm.allreduce_sum(rcs)
            self.stats["allreduce"] += 1
            for s, p, z in zip(self.solvers, self.parts, zs):
                check(self.lib.gmpnp_coarse_prolong_3d(s._h, ptr(rc), ptr(z), p.n_own, self._stream()), s._h)

    # -- GMRES(m), right-preconditioned, CGS2 ------------------------------------------------------
    def gmres(self, bs, m: int = 50, maxit: int = 500, rtol: float = 1e-10, callback=None):
        """Solve J x = b (owned rows).  Returns (xs, iterations, relative residual estimate)."""
        dev = self.device
        K = len(self.parts)
        nl = [p.n_local * NC for p in self.parts]
        Vb = [torch.zeros(m + 1, nl[k], dtype=torch.float64, device=dev) for k in range(K)]
        xs = self.zeros()
        zs = self.zeros()
        ws = [torch.zeros(nl[k], dtype=torch.float64, device=dev) for k in range(K)]
        r = [b.clone().view(-1) for b in bs]
        beta0 = math.sqrt(float(self.dot_owned([t.view(1, -1) for t in r], 1, r)[0]))
        if not beta0 > 0.0:
            return xs, 0, 0.0
        total, rel = 0, 1.0
        one = torch.ones(1, dtype=torch.float64, device=dev)
        while total < maxit:
            beta = math.sqrt(float(self.dot_owned([t.view(1, -1) for t in r], 1, r)[0]))
            rel = beta / beta0
            if rel <= rtol:
                break
            self.lincomb([t.view(1, -1) for t in r], 1, one / beta, [V[0] for V in Vb])
            H = np.zeros((m + 1, m))
            cs, sn = np.zeros(m), np.zeros(m)
            g = np.zeros(m + 1)
            g[0] = beta
            jd = 0
            for j in range(m):
                # w = J M^-1 v_j
                self.precond([V[j] for V in Vb], [z.view(-1) for z in zs])
                wl = self.spmv(zs)
                for k in range(K):
                    ws[k] = wl[k].view(-1)
                # CGS2: pass 1, then pass 2 fused with ||w||^2
                d1 = self.dot_owned(Vb, j + 1, ws)
                self.lincomb(Vb, j + 1, -d1, ws, beta=1.0, ys=ws)
                d2n = self.dot_owned(Vb, j + 1, ws, extra_self=True)
                d2 = d2n[: j + 1]
                self.lincomb(Vb, j + 1, -d2, ws, beta=1.0, ys=ws)
                # one device->host read per iteration: h = d1 + d2, ||w'||^2 and sum d2^2 (Pythagoras for ||w''||)
                host = torch.cat([d1 + d2, d2n[j + 1:j + 2], (d2 * d2).sum().view(1)]).cpu().numpy()
                hv, ww, s2 = host[: j + 1], float(host[j + 1]), float(host[j + 2])
                nrm2 = ww - s2
                if not nrm2 > 1e-28 * max(ww, 1e-300):                       # cancellation: recompute directly
                    nrm2 = float(self.dot_owned([t.view(1, -1) for t in ws], 1, ws)[0])
                hn = math.sqrt(max(nrm2, 0.0))
                H[: j + 1, j] = hv
                H[j + 1, j] = hn
                for i in range(j):
                    t = cs[i] * H[i, j] + sn[i] * H[i + 1, j]
                    H[i + 1, j] = -sn[i] * H[i, j] + cs[i] * H[i + 1, j]
                    H[i, j] = t
                d = math.hypot(H[j, j], H[j + 1, j])
                cs[j], sn[j] = (H[j, j] / d, H[j + 1, j] / d) if d > 0 else (1.0, 0.0)
                H[j, j], H[j + 1, j] = d, 0.0
                g[j + 1] = -sn[j] * g[j]
                g[j] = cs[j] * g[j]
                jd = j + 1
                total += 1
                rel = abs(g[j + 1]) / beta0
                if callback:
                    callback(total, rel)
                if rel <= rtol or not hn > 0.0 or total >= maxit:
                    break
                self.lincomb([w.view(1, -1) for w in ws], 1, one / hn, [V[j + 1] for V in Vb])
            # x += M^-1 (V y)
            y = np.linalg.solve(np.triu(H[:jd, :jd]), g[:jd])
            coef = torch.as_tensor(y, dtype=torch.float64, device=dev)
            self.lincomb(Vb, jd, coef, ws)
            self.precond(ws, [z.view(-1) for z in zs])
            for k in range(K):
                n = self.parts[k].n_own * NC
                xs[k].view(-1)[:n] += zs[k].view(-1)[:n]
            # true residual
            Jx = self.spmv(xs)
            for k in range(K):
                n = self.parts[k].n_own * NC
                r[k] = bs[k].view(-1).clone()
                r[k][:n] -= Jx[k].view(-1)[:n]
        self.stats["gmres_iters"] += total
        return xs, total, rel

    # -- one damped Newton solve (dolfin semantics) ----------------------------------
    def newton(self, us, uns, rtol=1e-4, atol=1e-4, maxit=50, relax=0.9, lin_rtol=1e-10, lin_restart=50, lin_maxit=2000):
        Fs, r0 = self.assemble(us, uns, want_J=False)
        r, k, lin = r0, 0, 0
        conv = r0 < atol
        while not conv and k < maxit:
            Fs, _ = self.assemble(us, uns, want_J=True)
            dx, its, rel = self.gmres(Fs, m=lin_restart, maxit=lin_maxit, rtol=lin_rtol)
            lin += its
            for kk, p in enumerate(self.parts):
                n = p.n_own * NC
                us[kk].view(-1)[:n] -= relax * dx[kk].view(-1)[:n]
            k += 1
            Fs, r = self.assemble(us, uns, want_J=False)
            conv = (r / r0 < rtol) or (r < atol)
            if not math.isfinite(r):
                break
        return dict(iters=k, r0=r0, r=r, lin_iters=lin, converged=bool(conv))

    # -- the reference's loop 3D:782-858 on the partitioned mesh -----------------------------------------
    def median(self, us, comp: int) -> float:
        """np.median of component ``comp`` over ALL vertices of the mesh (3D:817-820): the owned values of every part are
        gathered on every rank (V doubles), sorted on the device, and the middle value (odd V) or the mean of the two
        middle values (even V) is returned -- an order statistic, so it is exact and identical on every rank."""
        vals = self.comm.gather_values([u[0, : p.n_own, comp] for u, p in zip(us, self.parts)])
        srt, _ = torch.sort(vals)
        n = srt.numel()
        assert n == self.mesh.x.shape[0], (n, self.mesh.x.shape[0])
        mid = srt[n // 2] if n % 2 else 0.5 * (srt[n // 2 - 1] + srt[n // 2])
        return float(mid)

    def march(self, n_steps: int, us=None, uns=None, callback=None, **newton_kw):
        """``for n in range(tot_num_steps)`` of 3D/MPNP_CO2ER_pore.py:782-858 for ONE problem on the partitioned mesh:
        u = 0, u_n = (1, .., 1, 0); per step the CO2 entry Dirichlet value (bc4 rebuild, 3D:835-838), one damped Newton
        solve (3D:789-799), the Sechenov update from the nodal medians of OH-, HCO3-, CO3-- and the cation
        (3D:817-833), u_n <- u (3D:856).  Like dolfin, a Newton solve that does not converge raises."""
        if us is None:
            us = self.zeros()
        if uns is None:
            uns = self.zeros()
            for x in uns:
                x[:, :, : NC - 1] = 1.0
        co2 = float(self.prm.extras["eq_scaled"][0])
        iters, lin_iters, co2_hist = [], [], []
        for step in range(n_steps):
            self.set_dirichlet(co2)
            out = self.newton(us, uns, **newton_kw)
            if not out["converged"]:
                raise RuntimeError(f"Newton solver did not converge at step {step}: r = {out['r']:.3e} after {out['iters']} iterations")
            iters.append(out["iters"]); lin_iters.append(out["lin_iters"]); co2_hist.append(co2)
            self.comm.halo([u[0] for u in us])                       # ghosts of the accepted state (u_n must be complete)
            med = [self.median(us, c) for c in (1, 2, 3, 7)]
            co2 = float(_params.sechenov_co2_scaled(self.prm, *med))
            for x, xn in zip(us, uns):
                xn.copy_(x)
            if callback is not None:
                callback(step, us, out)
        return dict(us=us, uns=uns, iters=iters, lin_iters=lin_iters, co2_entry=co2_hist, co2_next=co2)

    def close(self):
        for s in self.solvers:
            s.close()
