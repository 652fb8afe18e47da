"""Host-side mirror of the 3D pore hot path (batched problems that share one tet mesh).

Maps on the C-ABI (include/gmpnp.h) and on the reference call sites:

* ``assemble``  -> FFC kernels + SystemAssembler for the forms 3D/MPNP_CO2ER_pore.py:503-769
* ``spmv``      -> the matrix-vector product inside the linear solve (MUMPS in the reference, 3D:792;
                   restarted GMRES + block-Jacobi/coarse preconditioner here)
* ``newton``    -> ``solve(F == 0, u, bcs, {newton, relaxation 0.9})`` 3D:789-799
* ``march``     -> the pseudo-time loop 3D:782-858 with the Sechenov median update 3D:817-838
* ``steady``    -> steady equations (kappa = 0) with voltage continuation and the Sechenov fixed point
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, marking, params as _params
from ._lib import NewtonOpts, check, ptr

NC = 9


def bulk_state(batch: int, n: int, device) -> torch.Tensor:
    """u = (1,...,1,0): the reference's initial u_n (3D:427-432)."""
    u = torch.ones(batch, n, NC, dtype=torch.float64, device=device)
    u[:, :, NC - 1] = 0.0
    return u


class Solver3D:
    def __init__(self, mesh, dir_dofs: np.ndarray, batch: int = 1, device: int = 0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.GmpnpError("CUDA device required: the GMPNP hot path has no CPU fallback")
        self.device = torch.device("cuda", int(device))
        self.mesh = mesh
        xyz = np.ascontiguousarray(mesh.x, dtype=np.float64)
        tets = np.ascontiguousarray(mesh.cells, dtype=np.int32)
        self.dir_dofs = np.ascontiguousarray(dir_dofs, dtype=np.int32)
        self.n = int(xyz.shape[0])
        self.n_tet = int(tets.shape[0])
        self.batch = int(batch)
        self._h = C.c_void_p()
        check(self.lib.gmpnp_create_3d(C.byref(self._h), self.device.index,
                                       xyz.ctypes.data_as(C.POINTER(C.c_double)), self.n,
                                       tets.ctypes.data_as(C.POINTER(C.c_int)), self.n_tet,
                                       self.dir_dofs.ctypes.data_as(C.POINTER(C.c_int)), len(self.dir_dofs), 8,
                                       self.batch), self._h)
        nb = C.c_int()
        check(self.lib.gmpnp_pattern_3d(self._h, C.byref(nb), None, None), self._h)
        self.n_blocks = nb.value

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.gmpnp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def pattern(self):
        rp = np.zeros(self.n + 1, dtype=np.int32)
        ci = np.zeros(self.n_blocks, dtype=np.int32)
        check(self.lib.gmpnp_pattern_3d(self._h, None, rp.ctypes.data_as(C.POINTER(C.c_int)),
                                        ci.ctypes.data_as(C.POINTER(C.c_int))), self._h)
        return rp, ci

    def set_params(self, plist):
        P = np.ascontiguousarray(plist, dtype=np.float64) if isinstance(plist, np.ndarray) \
            else np.stack([p.pack() for p in plist])
        assert P.shape == (self.batch, _params.NPAR)
        check(self.lib.gmpnp_set_params(self._h, P.ctypes.data_as(C.POINTER(C.c_double)), self.batch), self._h)

    def set_dirichlet(self, vals: np.ndarray):
        vals = np.ascontiguousarray(vals, dtype=np.float64).reshape(self.batch, len(self.dir_dofs))
        check(self.lib.gmpnp_set_dirichlet_3d(self._h, vals.ctypes.data_as(C.POINTER(C.c_double)), self.batch),
              self._h)

    def set_facet_terms(self, wall_w=None, exit_facets=None, exit_area=None, jwall=None, kexit=None):
        """Switch the intended boundary integrals on (3D:474-499; arrays from ``marking.facet_terms`` and the
        per-problem ``J_wall`` / ``k_exit`` of ``params_3d``) or, with no arguments, back off (as executed)."""
        if wall_w is None:
            check(self.lib.gmpnp_set_facet_terms_3d(self._h, None, None, None, 0, None, None, self.batch), self._h)
            return
        ww = np.ascontiguousarray(wall_w, dtype=np.float64)
        ef = np.ascontiguousarray(exit_facets, dtype=np.int32).reshape(-1, 3)
        ea = np.ascontiguousarray(exit_area, dtype=np.float64)
        jw = np.ascontiguousarray(jwall, dtype=np.float64).reshape(self.batch, 8)
        ke = np.ascontiguousarray(kexit, dtype=np.float64).reshape(self.batch, 8)
        pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int)
        check(self.lib.gmpnp_set_facet_terms_3d(self._h, ww.ctypes.data_as(pd), ef.ctypes.data_as(pi),
                                                ea.ctypes.data_as(pd), len(ea), jw.ctypes.data_as(pd),
                                                ke.ctypes.data_as(pd), self.batch), self._h)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _chk(self, t):
        assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
        assert tuple(t.shape) == (self.batch, self.n, NC), tuple(t.shape)

    def assemble(self, u, un, want_F=True, want_J=True):
        self._chk(u); self._chk(un)
        F = torch.empty(self.batch, self.n, NC, dtype=torch.float64, device=self.device) if want_F else None
        J = torch.empty(self.batch, self.n_blocks, NC, NC, dtype=torch.float64, device=self.device) if want_J else None
        check(self.lib.gmpnp_assemble_3d(self._h, ptr(u), ptr(un), ptr(F), ptr(J), self._stream()), self._h)
        return F, J

    def spmv(self, J, x):
        self._chk(x)
        y = torch.empty_like(x)
        check(self.lib.gmpnp_spmv_3d(self._h, ptr(J), ptr(x), ptr(y), self._stream()), self._h)
        return y

    def newton(self, u, un, opts: NewtonOpts | None = None):
        opts = opts or NewtonOpts.reference_3d()
        self._chk(u); self._chk(un)
        dev = self.device
        it = torch.zeros(self.batch, dtype=torch.int32, device=dev)
        st = torch.zeros(self.batch, dtype=torch.int32, device=dev)
        li = torch.zeros(self.batch, dtype=torch.int32, device=dev)
        r0 = torch.zeros(self.batch, dtype=torch.float64, device=dev)
        r = torch.zeros(self.batch, dtype=torch.float64, device=dev)
        check(self.lib.gmpnp_newton_3d(self._h, ptr(u), ptr(un), C.byref(opts), ptr(it), ptr(r0), ptr(r), ptr(li),
                                       ptr(st), self._stream()), self._h)
        return dict(iters=it, r0=r0, r=r, lin_iters=li, status=st)

    def set_march_data(self, kind, tab, sech):
        """Per-DOF Dirichlet kinds (0..4, ``marking.dirichlet_sets``), per-problem value table [batch, 4] =
        (wall potential, initial CO2 entry value, CO entry, H2 entry) and Sechenov records [batch, 8]
        (``PoreProblem.march_data``) for the library-side loops ``march`` / ``steady``."""
        kind = np.ascontiguousarray(kind, dtype=np.int8)
        tab = np.ascontiguousarray(tab, dtype=np.float64).reshape(self.batch, 4)
        sech = np.ascontiguousarray(sech, dtype=np.float64).reshape(self.batch, 8)
        assert kind.shape == (len(self.dir_dofs),)
        check(self.lib.gmpnp_set_march_data_3d(self._h, kind.ctypes.data_as(C.POINTER(C.c_byte)),
                                               tab.ctypes.data_as(C.POINTER(C.c_double)),
                                               sech.ctypes.data_as(C.POINTER(C.c_double)), self.batch), self._h)

    def march(self, u, un, n_steps: int, opts: NewtonOpts | None = None, history: bool = False):
        """The reference's loop 3D:782-858 inside the library: per step the Dirichlet values (incl. the current CO2
        entry value), one damped Newton solve, the Sechenov update from the nodal medians, u_n <- u.  A problem whose
        Newton solve fails stops (status), the others go on."""
        opts = opts or NewtonOpts.reference_3d()
        self._chk(u); self._chk(un)
        dev, B = self.device, self.batch
        hist = torch.zeros(B, n_steps, self.n, NC, dtype=torch.float64, device=dev) if history else None
        it = torch.zeros(B, n_steps, dtype=torch.int32, device=dev)
        li = torch.zeros(B, n_steps, dtype=torch.int32, device=dev)
        co2 = torch.zeros(B, n_steps, dtype=torch.float64, device=dev)
        steps = torch.zeros(B, dtype=torch.int32, device=dev)
        st = torch.zeros(B, dtype=torch.int32, device=dev)
        check(self.lib.gmpnp_march_3d(self._h, ptr(u), ptr(un), int(n_steps), C.byref(opts), ptr(hist), ptr(it), ptr(li),
                                      ptr(co2), ptr(steps), ptr(st), self._stream()), self._h)
        return dict(history=hist, iters=it, lin_iters=li, co2_entry=co2, steps=steps, status=st)

    def steady(self, u, un, opts: NewtonOpts | None = None, tol: float = 1e-10, max_steps: int = 200, n_ramp: int = 1):
        """Pseudo-time march to the steady state inside the library, wall voltage ramped over ``n_ramp`` steps; stops
        when the last relative increment of every problem still alive is <= tol (after the ramp)."""
        opts = opts or NewtonOpts.reference_3d()
        self._chk(u); self._chk(un)
        dev, B = self.device, self.batch
        it = torch.zeros(B, max_steps, dtype=torch.int32, device=dev)
        inc = torch.zeros(max_steps, B, dtype=torch.float64, device=dev)
        co2 = torch.zeros(B, dtype=torch.float64, device=dev)
        steps = torch.zeros(B, dtype=torch.int32, device=dev)
        st = torch.zeros(B, dtype=torch.int32, device=dev)
        cv = torch.zeros(B, dtype=torch.int32, device=dev)
        run = C.c_int(0)
        check(self.lib.gmpnp_steady_3d(self._h, ptr(u), ptr(un), C.byref(opts), float(tol), int(max_steps), int(n_ramp),
                                       ptr(it), ptr(inc), ptr(co2), ptr(steps), ptr(st), ptr(cv), C.byref(run),
                                       self._stream()), self._h)
        k = run.value
        return dict(iters=it[:, :k], increments=inc[:k], co2_entry=co2, steps=steps, status=st, converged=cv, steps_run=k)

    def median(self, u, comp: int):
        self._chk(u)
        med = torch.empty(self.batch, dtype=torch.float64, device=self.device)
        check(self.lib.gmpnp_median_3d(self._h, ptr(u), int(comp), ptr(med), self._stream()), self._h)
        return med

    def grad_project(self, u, n_iter: int = 40):
        """P1 L2 projection of grad(u_i) for all 9 components: [batch, n, 9, 3] (3D:884-909)."""
        self._chk(u)
        g = torch.empty(self.batch, self.n, NC, 3, dtype=torch.float64, device=self.device)
        check(self.lib.gmpnp_grad_project_3d(self._h, ptr(u), ptr(g), int(n_iter), self._stream()), self._h)
        return g

    def launch_count(self) -> int:
        return int(self.lib.gmpnp_launch_count(self._h))


class PoreProblem:
    """One pore geometry + a batch of parameter points: Dirichlet sets per the reference's marking
    (3D:335-379, 460-467) and the drivers built on :class:`Solver3D`."""

    def __init__(self, mesh, L: float, R: float, plist, device: int = 0, intended_bcs: bool = False,
                 rxn_diff: bool = False):
        """``intended_bcs``: add the wall-flux and pore-exit Robin integrals the reference's author wrote but Python
        discards (3D:474-499, 560-750; SURVEY finding 3 / App. H).  Default False = parity with the script as executed.
        ``rxn_diff``: the parameter points describe ``3D/rxn_diff_CO2ER_pore.py`` (z = nu = 0, V = 0; see
        ``gmpnp_b200/rxn_diff3d.py``): the cation starts at its bulk value instead of 0 and the Sechenov update takes
        the electroneutral cation estimate of RD3:589-592."""
        self.rxn_diff = bool(rxn_diff)
        self.mesh, self.L, self.R = mesh, L, R
        self.plist = list(plist)
        self.dofs, self.kind, self.info = marking.dirichlet_sets(mesh, L, R)
        self.solver = Solver3D(mesh, self.dofs, batch=len(self.plist), device=device)
        self.solver.set_params(self.plist)
        self.device = self.solver.device
        self.intended_bcs = bool(intended_bcs)
        if self.intended_bcs:
            ww, ef, ea = marking.facet_terms(mesh, L, R)
            self.solver.set_facet_terms(ww, ef, ea, np.stack([p.extras["J_wall"] for p in self.plist]),
                                        np.stack([p.extras["k_exit"] for p in self.plist]))

    def dirichlet_values(self, co2_scaled, V=None):
        vals = []
        for b, p in enumerate(self.plist):
            eq = p.extras["eq_scaled"]
            v = p.V if V is None else V[b]
            vals.append(marking.dirichlet_values(self.kind, v, co2_scaled[b], eq[1], eq[2]))
        return np.stack(vals)

    def _sechenov_update(self, u):
        """CO2 entry value of every problem from the nodal MEDIANS (3D:817-838; RD3:575-601)."""
        s = self.solver
        if not self.rxn_diff:
            med = [s.median(u, c).cpu().numpy() for c in (1, 2, 3, 7)]
            return [_params.sechenov_co2_scaled(p, med[0][b], med[1][b], med[2][b], med[3][b])
                    for b, p in enumerate(self.plist)]
        med = [s.median(u, c).cpu().numpy() for c in (0, 1, 2, 3)]
        out = []
        for b, p in enumerate(self.plist):
            c0 = p.c0
            cat = med[2][b] * c0[2] + 2 * med[3][b] * c0[3] + med[1][b] * c0[1] - med[0][b] * c0[0]   # RD3:589-592
            out.append(_params.sechenov_co2_scaled(p, med[1][b], med[2][b], med[3][b], cat / c0[7]))
        return out

    def march_data(self, V=None):
        """Arrays for ``Solver3D.set_march_data``: value table and Sechenov records (CO2_conc, 3D:70-93, with the
        constant factors folded: co2_scaled = A * 10^-(sum_k coef_k * median_k))."""
        import math
        tab, sech = [], []
        for b, p in enumerate(self.plist):
            e = p.extras
            eq = e["eq_scaled"]
            tab.append([p.V if V is None else V[b], float(eq[0]), float(eq[1]), float(eq[2])])
            h = e["h_sechenov"]
            temp = e["temp"]
            lnK = 93.4517 * (100 / temp) - 60.2409 + 23.3585 * math.log(temp / 100)
            h_co2 = h["CO2_0"] + h["CO2_T"] * (temp - 298.15)
            A = e["fugacity_CO2"] * math.exp(lnK) * 1000 / p.c0[4]
            cat = p.species[-1]
            c = p.c0
            if not self.rxn_diff:
                sech.append([A, (h["OH"] + h_co2) * c[1] / 1000, (h["HCO3"] + h_co2) * c[2] / 1000,
                             (h["CO32"] + h_co2) * c[3] / 1000, (h[cat] + h_co2) * c[7] / 1000, 0.0, 0.0, 0.0])
            else:           # electroneutral cation c_HCO3 + 2 c_CO32 + c_OH - c_H (RD3:589-592) folded into the ion terms
                hc = h[cat] + h_co2
                sech.append([A, (h["OH"] + h_co2 + hc) * c[1] / 1000, (h["HCO3"] + h_co2 + hc) * c[2] / 1000,
                             (h["CO32"] + h_co2 + 2 * hc) * c[3] / 1000, 0.0, 1.0, -hc * c[0] / 1000, 0.0])
        return self.kind.astype(np.int8), np.array(tab), np.array(sech)

    def march(self, n_steps: int, opts: NewtonOpts | None = None, history=True):
        """The reference's loop (3D:782-858): u = 0, u_n = (1,..,1,0); per step one damped Newton solve,
        then the CO2 entry Dirichlet value is re-evaluated from the nodal MEDIANS (3D:817-838).  The whole loop runs
        inside the library (``gmpnp_march_3d``); like dolfin, a failed Newton solve raises."""
        opts = opts or NewtonOpts.reference_3d()
        s = self.solver
        B = s.batch
        s.set_params(self.plist)
        s.set_march_data(*self.march_data())
        u = torch.zeros(B, s.n, NC, dtype=torch.float64, device=self.device)
        if self.rxn_diff:
            u[:, :, 7] = 1.0                               # passenger cation: stays at its bulk value
        un = bulk_state(B, s.n, self.device)
        first = un.cpu().numpy().copy() if history else None
        out = s.march(u, un, n_steps, opts, history=history)
        st = out["status"].cpu().numpy()
        if (st != 0).any():
            raise RuntimeError(f"Newton solver did not converge: status {st.tolist()}")   # dolfin raises too
        hist = None
        if history:
            hist = np.concatenate([first[None], out["history"].permute(1, 0, 2, 3).cpu().numpy()])
        return dict(u=u, history=hist, iters=out["iters"].cpu().numpy().T, lin_iters=out["lin_iters"].cpu().numpy().T,
                    co2_entry=out["co2_entry"].cpu().numpy().T)

    def steady(self, opts: NewtonOpts | None = None, tol: float = 1e-10, max_steps: int = 200,
               u0=None, dv_max: float | None = None, raise_on_failure: bool = True):
        """Steady state as the limit of the reference's pseudo-time march, optionally with a VOLTAGE RAMP, inside the
        library (``gmpnp_steady_3d``).

        With the as-executed boundary conditions (3D:460-467, no facet integrals -- SURVEY finding 3) every ionic
        species is pure-Neumann, so the time-independent equations are singular: the total amount of, e.g., the
        cation is fixed only by the initial state, which backward Euler conserves exactly.  The steady state is
        therefore computed the way the reference reaches it (3D:782-858: backward Euler with dt_scaled = 73.84,
        Sechenov median update per step) and marched until max|u - u_n| <= tol * max(1, max|u|) for every problem.
        Each step is one damped Newton solve (relaxation 0.9, 3D:796) from the previous state.

        ``dv_max`` (in V_T): the reference applies the full wall voltage in the first step, from which its Newton
        iteration diverges beyond |V| ~ 1.5 V_T on L_50_R_5; with ``dv_max`` the wall voltage of every problem is
        ramped proportionally over ceil(max|V| / dv_max) pseudo-time steps (voltage continuation, BASELINE
        north_star), and the convergence test only starts once the ramp is complete.  A problem whose Newton solve
        fails is parked (status != 0, ``converged`` False); with ``raise_on_failure`` that raises like dolfin."""
        opts = opts or NewtonOpts.reference_3d()
        s = self.solver
        B = s.batch
        s.set_params(self.plist)
        s.set_march_data(*self.march_data())
        un = bulk_state(B, s.n, self.device) if u0 is None else u0.clone()
        u = un.clone()
        Vt = np.array([p.V for p in self.plist], dtype=np.float64)
        n_ramp = 1 if not dv_max else max(1, int(np.ceil(np.abs(Vt).max() / dv_max - 1e-12)))
        out = s.steady(u, un, opts, tol=tol, max_steps=max_steps, n_ramp=n_ramp)
        st = out["status"].cpu().numpy()
        if raise_on_failure and (st != 0).any():
            raise RuntimeError(f"Newton solver did not converge in the pseudo-time march: status {st.tolist()}")
        inc = out["increments"].cpu().numpy()                       # [steps, B]
        conv = (st == 0) & (out["converged"].cpu().numpy() != 0)
        return dict(u=u, iters=out["iters"].cpu().numpy().T, increments=inc.max(axis=1), increments_per_problem=inc,
                    co2_entry=out["co2_entry"].cpu().numpy(), steps=out["steps_run"], status=st, converged=conv,
                    steps_per_problem=out["steps"].cpu().numpy())
