"""Global assembly, Dirichlet handling and Newton drivers of the oracle (test infrastructure).

* assembly / scatter: what dolfin's SystemAssembler does for ``solve(F == 0, u, bcs)``
  (1D/MPNP_CO2ER_EDL.py:737-742, 3D/MPNP_CO2ER_pore.py:789-799);
* Dirichlet sets: 1D:237-254, 350-355; 3D:460-467 (list order, later BC wins);
* Newton: dolfin 2019.1 ``NewtonSolver`` semantics as written out in SURVEY App. C;
* march: the reference's pseudo-time loop 1D:633-796 / 3D:782-858 (incl. the H_OHP ladder
  1D:766-793 and the Sechenov median update 3D:817-838);
* steady: the time term dropped, voltage continuation, tight increment criterion.

DOF numbering: node-major, ``dof = ncomp * vertex + component``.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import forms, quadrature


class Discretisation:
    """Mesh + fixed sparsity data for one (mesh, ncomp) pair."""

    def __init__(self, x, cells, ncomp, jac_rule: int = 0, facet_terms=None):
        """``jac_rule`` 0: FFC's rule pair (Jacobian integrated with the degree-4 rule, SURVEY App. B -- the
        reference's iteration path); 1: the Jacobian uses the residual's rule (exact derivative of the discrete
        residual: quadratic convergence, same converged solution)."""
        self.x = np.asarray(x, dtype=np.float64)
        if self.x.ndim == 1:
            self.x = self.x[:, None]
        self.cells = np.asarray(cells, dtype=np.int64)
        self.ncomp = ncomp
        self.dim = self.x.shape[1]
        self.nv = self.x.shape[0]
        self.ndof = ncomp * self.nv
        self.g, self.vol = forms.geometry(self.x, self.cells)
        self.ruleF, self.ruleJ = quadrature.rules_for_dim(self.dim)
        if jac_rule == 1:
            self.ruleJ = self.ruleF
        # intended 3D boundary integrals (3D/MPNP_CO2ER_pore.py:474-499, dead code as executed; live in
        # 3D/rxn_diff_CO2ER_pore.py:480-511): facet_terms = (wall_w[nv], exit_facets[nf,3], exit_area[nf], J_wall[8], k_exit[8])
        self.facet = None
        if facet_terms is not None:
            wall_w, ef, ea, jwall, kexit = facet_terms
            ef = np.asarray(ef, dtype=np.int64).reshape(-1, 3)
            r = np.repeat(ef, 3, axis=1).ravel()
            c = np.tile(ef, (1, 3)).ravel()
            w = np.where(np.repeat(np.arange(3), 3)[None, :] == np.tile(np.arange(3), 3)[None, :], 1.0 / 6.0, 1.0 / 12.0)
            E = sp.coo_matrix(((w * np.asarray(ea)[:, None]).ravel(), (r, c)), shape=(self.nv, self.nv)).tocsr()
            self.facet = (np.asarray(wall_w, float), E, np.asarray(jwall, float), np.asarray(kexit, float))
        dofs = (self.cells[:, :, None] * ncomp + np.arange(ncomp)[None, None, :])   # [c, a, i]
        self.cell_dofs = dofs
        nloc = dofs.shape[1] * ncomp
        d = dofs.reshape(len(self.cells), nloc)
        self.rows = np.repeat(d, nloc, axis=1).ravel()
        self.cols = np.tile(d, (1, nloc)).ravel()

    def gather(self, u):
        return u.reshape(self.nv, self.ncomp)[self.cells]          # [c, a, i]

    def residual(self, u, un, prm, point_flux=None):
        Fe = forms.element_residual(self.gather(u), self.gather(un), self.g, self.vol, prm, self.ruleF)
        F = np.zeros(self.ndof, dtype=Fe.dtype)
        np.add.at(F, self.cell_dofs.ravel(), Fe.ravel())
        if point_flux is not None:          # 1D `J_i * v_i * ds`: point evaluation at BOTH ends (1D:553, 738)
            ns = self.ncomp - 1
            for node in (0, self.nv - 1):
                F[node * self.ncomp: node * self.ncomp + ns] += point_flux[:ns]
        if self.facet is not None:
            wall_w, E, jwall, kexit = self.facet
            ns = self.ncomp - 1
            U = u.reshape(self.nv, self.ncomp)
            Ff = F.reshape(self.nv, self.ncomp)
            Ff[:, :ns] += wall_w[:, None] * jwall[None, :ns] + (E @ (U[:, :ns] - 1.0)) * kexit[None, :ns]
        return F

    def jacobian(self, u, prm):
        Je = forms.element_jacobian(self.gather(u), self.g, self.vol, prm, self.ruleJ)
        A = sp.coo_matrix((Je.ravel(), (self.rows, self.cols)), shape=(self.ndof, self.ndof))
        if self.facet is not None:
            _, E, _, kexit = self.facet
            Ec = E.tocoo()
            ns = self.ncomp - 1
            rr = (Ec.row[:, None] * self.ncomp + np.arange(ns)[None, :]).ravel()
            cc = (Ec.col[:, None] * self.ncomp + np.arange(ns)[None, :]).ravel()
            vv = (Ec.data[:, None] * kexit[None, :ns]).ravel()
            A = A + sp.coo_matrix((vv, (rr, cc)), shape=(self.ndof, self.ndof))
        return A.tocsr()


def apply_bc_residual(b, x, bc_dofs, bc_vals):
    """b[dof] = x[dof] - g; bcs given in application order (later wins automatically)."""
    b[bc_dofs] = x[bc_dofs] - bc_vals
    return b


def apply_bc_matrix(A, bc_dofs):
    """Dirichlet rows -> identity rows."""
    A = A.tolil(copy=False) if False else A
    mask = np.zeros(A.shape[0], dtype=bool)
    mask[bc_dofs] = True
    keep = sp.diags((~mask).astype(np.float64))
    return (keep @ A + sp.diags(mask.astype(np.float64))).tocsc()


def bc_1d(nv, ncomp, V):
    """1D Dirichlet set: bulk (1,...,1,0) on all components at x=1, potential V at x=0
    (1D:350-355).  Returns (dofs, values) in application order."""
    last = (nv - 1) * ncomp
    dofs = list(range(last, last + ncomp)) + [ncomp - 1]
    vals = [1.0] * (ncomp - 1) + [0.0] + [V]
    return np.array(dofs, dtype=np.int64), np.array(vals, dtype=np.float64)


def newton(disc, prm, u, un, bc_dofs, bc_vals, point_flux=None, rtol=1e-4, atol=1e-4, maxit=50,
           relax=1.0, criterion="residual", xtol=1e-12, history=None, xtol_floor=0.0, info=None):
    """dolfin NewtonSolver.  Returns (u, iterations, converged, r0, r).

    criterion 'residual': stop when ||b||/||b0|| < rtol or ||b|| < atol (reference).
    criterion 'increment': stop when ||dx||_inf <= xtol * max(1, ||x||_inf)  (steady mode).
    ``xtol_floor`` > 0 mirrors the opt-in rule of include/gmpnp.h (gmpnp_newton_opts.xtol_floor): an increment that
    stopped contracting between xtol and xtol_floor ends the iteration with ``info['stagnated'] = True`` and
    converged = False.  ``info`` (dict, optional) also receives 'dx_rel' (last relative increment) and the wall time
    spent in 'assembly' (residual + Jacobian) and 'lu' (sparse LU factorisation + solve)."""
    import time as _time
    x = u.copy()
    t0 = _time.perf_counter()
    b = apply_bc_residual(disc.residual(x, un, prm, point_flux), x, bc_dofs, bc_vals)
    t_asm, t_lu = _time.perf_counter() - t0, 0.0
    r0 = float(np.linalg.norm(b))
    r = r0
    k = 0
    conv = (r < atol) if criterion == "residual" else False
    stagnated = False
    dx_prev, dx_rel = np.inf, np.inf
    while not conv and k < maxit:
        t0 = _time.perf_counter()
        A = apply_bc_matrix(disc.jacobian(x, prm), bc_dofs)
        t1 = _time.perf_counter()
        dx = spla.splu(A).solve(b)
        t2 = _time.perf_counter()
        x = x - relax * dx
        k += 1
        b = apply_bc_residual(disc.residual(x, un, prm, point_flux), x, bc_dofs, bc_vals)
        t3 = _time.perf_counter()
        t_asm += (t1 - t0) + (t3 - t2)
        t_lu += t2 - t1
        r = float(np.linalg.norm(b))
        dxmax = float(np.abs(dx).max())
        if history is not None:
            history.append((k, r, dxmax))
        if criterion == "residual":
            conv = (r / r0 < rtol) or (r < atol)
        else:
            scale = max(1.0, float(np.abs(x).max()))
            conv = dxmax <= xtol * scale
            if not conv and xtol_floor > 0 and k >= 3 and dxmax <= xtol_floor * scale and dxmax >= 0.25 * dx_prev:
                stagnated = True
            dx_prev, dx_rel = dxmax, dxmax / scale
            if stagnated:
                break
        if not np.isfinite(r):
            break
    if info is not None:
        info["stagnated"] = stagnated
        info["dx_rel"] = dx_rel
        info["assembly"] = info.get("assembly", 0.0) + t_asm
        info["lu"] = info.get("lu", 0.0) + t_lu
    return x, k, conv, r0, r


def h_ohp_ladder(frac, H_OHP_frac, H_OHP):
    """Proton-current controller of 1D:770-782."""
    if H_OHP_frac < 0:
        frac = frac / 1.1
    elif H_OHP_frac < (H_OHP - 0.05):
        frac = frac / 1.05
    elif H_OHP_frac < (H_OHP - 0.025):
        frac = frac / 1.01
    elif (H_OHP_frac > H_OHP and H_OHP_frac <= (H_OHP + 0.4) and frac <= 1.0):
        frac = frac * 1.04
    elif H_OHP_frac > (H_OHP + 0.4) and frac <= 1.0:
        frac = frac * 1.15
    return frac


def march_1d(x, prm, n_steps, H_OHP=None, rtol=1e-4, atol=1e-4, maxit=50):
    """The reference's 1D pseudo-time loop (1D:633-796): u starts at 0, u_n at (1,..,1,0)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    nv = len(x)
    cells = np.stack([np.arange(nv - 1), np.arange(1, nv)], axis=1)
    ncomp = prm.ns + 1
    disc = Discretisation(x, cells, ncomp)
    bc_dofs, bc_vals = bc_1d(nv, ncomp, prm.V)
    u = np.zeros(disc.ndof)
    un = np.tile(np.array([1.0] * prm.ns + [0.0]), nv)
    hist = [un.reshape(nv, ncomp).copy()]
    its = []
    flux = prm.jflux.copy()
    frac = prm.extras.get("current_H_frac", 0.0)
    for _ in range(n_steps):
        u, k, conv, r0, r = newton(disc, prm, u, un, bc_dofs, bc_vals, point_flux=flux,
                                   rtol=rtol, atol=atol, maxit=maxit)
        if not conv:
            raise RuntimeError("Newton solver did not converge")
        its.append(k)
        hist.append(u.reshape(nv, ncomp).copy())
        if H_OHP is not None:
            frac = h_ohp_ladder(frac, u[0], H_OHP)
            e = prm.extras
            flux[1] = -1.0 * e["J_OH_prefactor"] * e["current_OHP_ss"] * (1 - frac)
            flux[0] = e["J_H_prefactor"] * e["current_OHP_ss"] * frac
        un = u.copy()
    return np.array(hist), its, frac


def steady_1d(x, prm, V_path, u0=None, xtol=1e-12, maxit=50, xtol_path=0.0, jac_rule=0, info=None):
    """Steady equations (kappa = 0) with voltage continuation along ``V_path``.
    Starts from the bulk state unless ``u0`` is given.  Returns (u[nv, ncomp] at the last V,
    list of Newton counts)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    nv = len(x)
    cells = np.stack([np.arange(nv - 1), np.arange(1, nv)], axis=1)
    ncomp = prm.ns + 1
    disc = Discretisation(x, cells, ncomp, jac_rule=jac_rule)
    ps = prm.with_(kappa=0.0)
    u = np.tile(np.array([1.0] * prm.ns + [0.0]), nv) if u0 is None else np.asarray(u0, float).reshape(-1).copy()
    its = []
    V_path = list(V_path)
    for s_, V in enumerate(V_path):
        pv = ps.with_(V=float(V))
        bc_dofs, bc_vals = bc_1d(nv, ncomp, float(V))
        tol = xtol if (s_ + 1 == len(V_path) or not xtol_path > 0) else xtol_path
        u, k, conv, r0, r = newton(disc, pv, u, u, bc_dofs, bc_vals, point_flux=pv.jflux,
                                   criterion="increment", xtol=tol, maxit=maxit, info=info)
        if not conv:
            raise RuntimeError(f"steady Newton failed at V={V} after {k} its (r={r})")
        its.append(k)
    return u.reshape(nv, ncomp), its


def p1_gradient_projection_1d(x, f):
    """L2 projection of df/dx onto P1 (dolfin ``project(grad(f), W)``, 1D:802-803):
    solve M g = b with the consistent P1 mass matrix."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    h = np.diff(x)
    n = len(x)
    slope = np.diff(f) / h
    b = np.zeros(n)
    b[:-1] += 0.5 * h * slope
    b[1:] += 0.5 * h * slope
    main = np.zeros(n)
    main[:-1] += h / 3
    main[1:] += h / 3
    M = sp.diags([h / 6, main, h / 6], [-1, 0, 1]).tocsc()
    return spla.splu(M).solve(b)


# ---------------------------------------------------------------------------------------------
# 3D pore
# ---------------------------------------------------------------------------------------------

def p1_gradient_projection_3d(mesh_x, mesh_cells, f):
    """dolfin ``project(grad(f), W)`` for P1 nodal fields f[nv, k] (3D/MPNP_CO2ER_pore.py:884-909): solve
    M g = b with the consistent P1 mass matrix, b_v = sum_{t ni v} vol_t/4 grad(f)_t.  Returns g[nv, k, 3]."""
    x = np.asarray(mesh_x, dtype=np.float64)
    cells = np.asarray(mesh_cells, dtype=np.int64)
    g, vol = forms.geometry(x, cells)                    # g[c, a, d] = grad lambda_a
    nv = x.shape[0]
    f = np.asarray(f, dtype=np.float64).reshape(nv, -1)
    gradf = np.einsum("cad,cak->ckd", g, f[cells])      # constant per cell
    b = np.zeros((nv, f.shape[1], 3))
    for a in range(4):
        np.add.at(b, cells[:, a], 0.25 * vol[:, None, None] * gradf)
    rows = np.repeat(cells, 4, axis=1).ravel()
    cols = np.tile(cells, (1, 4)).ravel()
    w = np.where(np.repeat(np.arange(4), 4)[None, :] == np.tile(np.arange(4), 4)[None, :], 0.1, 0.05) * vol[:, None]
    M = sp.coo_matrix((w.ravel(), (rows, cols)), shape=(nv, nv)).tocsc()
    lu = spla.splu(M)
    return lu.solve(b.reshape(nv, -1)).reshape(nv, f.shape[1], 3)


def march_3d(mesh_x, mesh_cells, prm, bc_dofs, bc_kind, n_steps, rtol=1e-4, atol=1e-4, maxit=50, relax=0.9,
             sechenov=None, facet_terms=None):
    """The reference's 3D loop (3D:782-858): u starts at 0, u_n at (1,..,1,0); damped Newton
    (relaxation 0.9, 3D:796); after every step the CO2 entry value is re-evaluated from the MEDIANS of
    the nodal OH/HCO3/CO32/cation values (3D:817-838) via ``sechenov(med_OH, med_HCO3, med_CO32, med_cat)``.
    Returns (history [n_steps+1, nv, 9], Newton counts, list of CO2 entry values used)."""
    ncomp = prm.ns + 1
    disc = Discretisation(mesh_x, mesh_cells, ncomp, facet_terms=facet_terms)
    nv = disc.nv
    eq = prm.extras["eq_scaled"]
    co2 = float(eq[0])
    table = lambda c: np.array([0.0, prm.V, c, eq[1], eq[2]])
    u = np.zeros(disc.ndof)
    un = np.tile(np.array([1.0] * prm.ns + [0.0]), nv)
    hist = [un.reshape(nv, ncomp).copy()]
    its, co2s = [], []
    for _ in range(n_steps):
        vals = table(co2)[bc_kind.astype(np.int64)]
        co2s.append(co2)
        u, k, conv, r0, r = newton(disc, prm, u, un, bc_dofs, vals, rtol=rtol, atol=atol, maxit=maxit, relax=relax)
        if not conv:
            raise RuntimeError("Newton solver did not converge")
        its.append(k)
        U = u.reshape(nv, ncomp)
        hist.append(U.copy())
        if sechenov is not None:
            co2 = sechenov(np.median(U[:, 1]), np.median(U[:, 2]), np.median(U[:, 3]), np.median(U[:, prm.ns - 1]))
        un = u.copy()
    return np.array(hist), its, co2s


def steady_march_3d(mesh_x, mesh_cells, prm, bc_dofs, bc_kind, tol=1e-8, max_steps=40, n_ramp=1, sechenov=None,
                    rtol=1e-4, atol=1e-4, maxit=50, relax=0.9):
    """Steady state as the limit of the reference's pseudo-time march (3D:782-858) with the wall voltage ramped linearly
    over the first ``n_ramp`` steps (voltage continuation) -- the algorithm of ``gmpnp_steady_3d`` (include/gmpnp.h):
    starts from the bulk state; per step the Dirichlet values with the current CO2 entry value, one damped Newton solve,
    the Sechenov update from the nodal medians, u_n <- u; stops at the first step >= n_ramp whose relative increment
    max|u - u_n| / max(1, max|u|) is <= tol.  Returns (u[nv, ncomp], Newton counts, increments, CO2 entry value)."""
    ncomp = prm.ns + 1
    disc = Discretisation(mesh_x, mesh_cells, ncomp)
    nv = disc.nv
    eq = prm.extras["eq_scaled"]
    co2 = float(eq[0])
    un = np.tile(np.array([1.0] * prm.ns + [0.0]), nv)
    u = un.copy()
    its, incs = [], []
    for s_ in range(max_steps):
        V = prm.V * min(1.0, (s_ + 1) / n_ramp)
        vals = np.array([0.0, V, co2, eq[1], eq[2]])[bc_kind.astype(np.int64)]
        u, k, conv, r0, r = newton(disc, prm.with_(V=V), u, un, bc_dofs, vals, rtol=rtol, atol=atol, maxit=maxit, relax=relax)
        if not conv:
            raise RuntimeError(f"Newton solver did not converge in pseudo-time step {s_}")
        its.append(k)
        U = u.reshape(nv, ncomp)
        if sechenov is not None:
            co2 = sechenov(np.median(U[:, 1]), np.median(U[:, 2]), np.median(U[:, 3]), np.median(U[:, prm.ns - 1]))
        inc = float(np.abs(u - un).max() / max(1.0, np.abs(u).max()))
        incs.append(inc)
        un = u.copy()
        if s_ + 1 >= n_ramp and inc <= tol:
            break
    return u.reshape(nv, ncomp), its, incs, co2


def steady_3d(mesh_x, mesh_cells, prm, bc_dofs, bc_kind, V_path, co2_entry, u0=None, xtol=1e-12, maxit=50):
    """Steady 3D equations (kappa = 0) with voltage continuation, full Newton steps, fixed CO2 entry value."""
    ncomp = prm.ns + 1
    disc = Discretisation(mesh_x, mesh_cells, ncomp)
    nv = disc.nv
    eq = prm.extras["eq_scaled"]
    ps = prm.with_(kappa=0.0)
    u = np.tile(np.array([1.0] * prm.ns + [0.0]), nv) if u0 is None else np.asarray(u0, float).reshape(-1).copy()
    its = []
    for V in V_path:
        vals = np.array([0.0, float(V), co2_entry, eq[1], eq[2]])[bc_kind.astype(np.int64)]
        u, k, conv, r0, r = newton(disc, ps.with_(V=float(V)), u, u, bc_dofs, vals, criterion="increment",
                                   xtol=xtol, maxit=maxit)
        if not conv:
            raise RuntimeError(f"steady Newton failed at V={V} after {k} its (r={r})")
        its.append(k)
    return u.reshape(nv, ncomp), its
