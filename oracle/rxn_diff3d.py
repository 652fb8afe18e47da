"""CPU restatement of ``3D/rxn_diff_CO2ER_pore.py`` (test infrastructure; parity unpinned -- FEniCS is absent).

An INDEPENDENT 7-species oracle (H, OH, HCO3, CO32, CO2, CO, H2; no potential, no cation, no steric term), written
directly from the reference's forms so that it can check the product's embedding of this model into the 9-component
GMPNP kernels (``gmpnp_b200/rxn_diff3d.py``: nu = 0, z = 0, passenger cation and potential) as well as the
9-component oracle (``oracle/solver.py``) with the same switches.

Follows, line by line:
* forms ``F_H .. F_H2``                      RD3:513-548 (time term, ``dot(grad u, grad v)``, ``- R_i v``, facet terms);
* reaction sources ``R_H .. R_CO2``          RD3:480-511 (CO, H2: none);
* wall fluxes ``J_*_wall * v * ds(2)``       RD3:422-431; pore-exit Robin ``J_pore_exit_i * v_i * ds(3)`` RD3:434-447;
* Dirichlet gases at the pore entry         RD3:408-412 (marker 1);
* ``solve(F == 0, u, bcs, newton/mumps, rtol = atol = 1e-4, maxit 50, relaxation 0.9)``  RD3:563-571, with dolfin's
  NewtonSolver semantics;
* the loop and the Sechenov update from the nodal MEDIANS with an electroneutral cation  RD3:557-611.

Quadrature: every ``dx`` integrand is a polynomial of degree <= 3 on affine P1 tets, and FFC selects rules that are
exact for the estimated degree (3 for F and for J: no rational term here), so the exact monomial integrals
``int l_a l_b = vol (1 + d_ab)/20``, ``int l_a l_b l_c = vol a!b!c!-pattern / 120``, ``int l_a = vol/4`` are used
-- independent of the quadrature tables in ``oracle/quadrature.py``.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

NS = 7
H, OH, HCO3, CO32, CO2, CO, H2 = range(NS)


def _geometry(x, cells):
    X = x[cells]                                           # [c, 4, 3]
    E = np.concatenate([X, np.ones((len(cells), 4, 1))], axis=2)          # rows (x, y, z, 1)
    inv = np.linalg.inv(E)                                 # columns: coefficients of lambda_a
    g = np.transpose(inv[:, :3, :], (0, 2, 1))             # [c, a, d] = grad lambda_a
    vol = np.abs(np.linalg.det(E)) / 6.0
    return g, vol


def _triple():
    """T[a, b, c] = int l_a l_b l_c / vol on a tet."""
    T = np.empty((4, 4, 4))
    for a in range(4):
        for b in range(4):
            for c in range(4):
                k = len({a, b, c})
                T[a, b, c] = {1: 6.0, 2: 2.0, 3: 1.0}[k] / 120.0
    return T


class RxnDiff3D:
    """Assembly for one mesh.  ``prm`` is the 8-species ``ProblemParams`` of ``params_3d`` (the reference computes the
    same groups in RD3:115-324); only its first seven species are used."""

    def __init__(self, x, cells, prm, wall_w, exit_facets, exit_area):
        self.x = np.asarray(x, float)
        self.cells = np.asarray(cells, np.int64)
        self.nv = self.x.shape[0]
        self.ndof = NS * self.nv
        self.g, self.vol = _geometry(self.x, self.cells)
        self.T = _triple()
        self.M = (np.ones((4, 4)) + np.eye(4)) / 20.0
        self.K = np.einsum("cad,cbd->cab", self.g, self.g) * self.vol[:, None, None]
        self.kappa = float(prm.kappa)                      # 1 / del_t
        c0, s, k = prm.c0, prm.extras["scale_R"], prm.rate
        self.c0, self.s = c0, s
        # products of rate constants and bulk concentrations as they stand in RD3:480-511
        self.k_w2 = k["kw2"] * c0[H] * c0[OH]
        self.k_w1 = k["kw1"]
        self.k_a1 = k["ka1"] * c0[OH] * c0[HCO3]
        self.k_a2 = k["ka2"] * c0[CO32]
        self.k_b1 = k["kb1"] * c0[CO2] * c0[OH]
        self.k_b2 = k["kb2"] * c0[HCO3]
        self.jwall = np.asarray(prm.extras["J_wall"], float)[:NS]
        self.kexit = np.asarray(prm.extras["k_exit"], float)[:NS]
        self.wall_w = np.asarray(wall_w, float)
        ef = np.asarray(exit_facets, np.int64).reshape(-1, 3)
        w = (np.ones((3, 3)) + np.eye(3)) / 12.0
        self.E = sp.coo_matrix(((w[None] * np.asarray(exit_area)[:, None, None]).ravel(),
                                (np.repeat(ef, 3, axis=1).ravel(), np.tile(ef, (1, 3)).ravel())),
                               shape=(self.nv, self.nv)).tocsr()
        d = (self.cells[:, :, None] * NS + np.arange(NS)[None, None, :]).reshape(len(self.cells), 4 * NS)
        self.cell_dofs = d
        self.rows = np.repeat(d, 4 * NS, axis=1).ravel()
        self.cols = np.tile(d, (1, 4 * NS)).ravel()

    # element-level pieces: U[c, a, i] nodal values
    def _P(self, f, g):
        return np.einsum("abd,cb,cd->ca", self.T, f, g) * self.vol[:, None]

    def _L(self, f):
        return np.einsum("ab,cb->ca", self.M, f) * self.vol[:, None]

    def residual(self, u, un):
        U = u.reshape(self.nv, NS)[self.cells]
        Un = un.reshape(self.nv, NS)[self.cells]
        Fe = np.zeros_like(U)
        for i in range(NS):
            Fe[:, :, i] = self.kappa * self._L(U[:, :, i] - Un[:, :, i]) + np.einsum("cab,cb->ca", self.K, U[:, :, i])
        C = 0.25 * self.vol[:, None] * np.ones((1, 4))
        Pw = self._P(U[:, :, H], U[:, :, OH])
        Pa = self._P(U[:, :, OH], U[:, :, HCO3])
        Pb = self._P(U[:, :, CO2], U[:, :, OH])
        L3, Lh = self._L(U[:, :, CO32]), self._L(U[:, :, HCO3])
        s = self.s
        # "- R_i * v_i * dx" with R_i = - scale_R_i * (...)
        Fe[:, :, H] += s[H] * (self.k_w2 * Pw - self.k_w1 * C)
        Fe[:, :, OH] += s[OH] * (self.k_w2 * Pw + self.k_a1 * Pa + self.k_b1 * Pb - self.k_w1 * C - self.k_a2 * L3
                                 - self.k_b2 * Lh)
        Fe[:, :, HCO3] += s[HCO3] * (self.k_a1 * Pa + self.k_b2 * Lh - self.k_a2 * L3 - self.k_b1 * Pb)
        Fe[:, :, CO32] += s[CO32] * (self.k_a2 * L3 - self.k_a1 * Pa)
        Fe[:, :, CO2] += s[CO2] * (self.k_b1 * Pb - self.k_b2 * Lh)
        F = np.zeros(self.ndof)
        np.add.at(F, self.cell_dofs.ravel(), Fe.reshape(len(self.cells), -1).ravel())
        Ff = F.reshape(self.nv, NS)
        Ug = u.reshape(self.nv, NS)
        Ff += self.wall_w[:, None] * self.jwall[None, :] + (self.E @ (Ug - 1.0)) * self.kexit[None, :]
        return F

    def jacobian(self, u):
        U = u.reshape(self.nv, NS)[self.cells]
        nc = len(self.cells)
        Je = np.zeros((nc, 4, NS, 4, NS))
        lin = self.kappa * self.M[None] * self.vol[:, None, None] + self.K
        for i in range(NS):
            Je[:, :, i, :, i] += lin
        Mv = self.M[None] * self.vol[:, None, None]

        def dP(g):                                         # d P(f, g)_a / d f_b
            return np.einsum("abd,cd->cab", self.T, g) * self.vol[:, None, None]

        s = self.s
        dPw_H, dPw_OH = dP(U[:, :, OH]), dP(U[:, :, H])
        dPa_OH, dPa_HCO3 = dP(U[:, :, HCO3]), dP(U[:, :, OH])
        dPb_CO2, dPb_OH = dP(U[:, :, OH]), dP(U[:, :, CO2])

        def add(row, col, block):
            Je[:, :, row, :, col] += block

        add(H, H, s[H] * self.k_w2 * dPw_H); add(H, OH, s[H] * self.k_w2 * dPw_OH)
        add(OH, H, s[OH] * self.k_w2 * dPw_H)
        add(OH, OH, s[OH] * (self.k_w2 * dPw_OH + self.k_a1 * dPa_OH + self.k_b1 * dPb_OH))
        add(OH, HCO3, s[OH] * (self.k_a1 * dPa_HCO3 - self.k_b2 * Mv))
        add(OH, CO2, s[OH] * self.k_b1 * dPb_CO2)
        add(OH, CO32, -s[OH] * self.k_a2 * Mv)
        add(HCO3, OH, s[HCO3] * (self.k_a1 * dPa_OH - self.k_b1 * dPb_OH))
        add(HCO3, HCO3, s[HCO3] * (self.k_a1 * dPa_HCO3 + self.k_b2 * Mv))
        add(HCO3, CO32, -s[HCO3] * self.k_a2 * Mv)
        add(HCO3, CO2, -s[HCO3] * self.k_b1 * dPb_CO2)
        add(CO32, CO32, s[CO32] * self.k_a2 * Mv)
        add(CO32, OH, -s[CO32] * self.k_a1 * dPa_OH)
        add(CO32, HCO3, -s[CO32] * self.k_a1 * dPa_HCO3)
        add(CO2, CO2, s[CO2] * self.k_b1 * dPb_CO2)
        add(CO2, OH, s[CO2] * self.k_b1 * dPb_OH)
        add(CO2, HCO3, -s[CO2] * self.k_b2 * Mv)
        A = sp.coo_matrix((Je.ravel(), (self.rows, self.cols)), shape=(self.ndof, self.ndof))
        Ec = self.E.tocoo()
        rr = (Ec.row[:, None] * NS + np.arange(NS)[None, :]).ravel()
        vv = (Ec.data[:, None] * self.kexit[None, :]).ravel()
        cc = (Ec.col[:, None] * NS + np.arange(NS)[None, :]).ravel()
        return (A + sp.coo_matrix((vv, (rr, cc)), shape=(self.ndof, self.ndof))).tocsr()


def newton(disc, u, un, bc_dofs, bc_vals, rtol=1e-4, atol=1e-4, maxit=50, relax=0.9):
    """dolfin NewtonSolver, residual criterion.  Returns (u, k, converged, r0, r)."""
    x = u.copy()
    mask = np.zeros(disc.ndof, bool)
    mask[bc_dofs] = True
    keep = sp.diags((~mask).astype(float))
    ident = sp.diags(mask.astype(float))

    def res(x):
        b = disc.residual(x, un)
        b[bc_dofs] = x[bc_dofs] - bc_vals
        return b

    b = res(x)
    r0 = r = float(np.linalg.norm(b))
    k = 0
    conv = r < atol
    while not conv and k < maxit:
        A = (keep @ disc.jacobian(x) + ident).tocsc()
        x = x - relax * spla.splu(A).solve(b)
        k += 1
        b = res(x)
        r = float(np.linalg.norm(b))
        conv = (r / r0 < rtol) or (r < atol)
    return x, k, conv, r0, r


def march(x, cells, prm, entry_verts, wall_w, exit_facets, exit_area, n_steps, sechenov=None):
    """RD3:557-611.  ``entry_verts``: vertices of the facets that carry marker 1; ``sechenov(med_OH, med_HCO3,
    med_CO32, med_cat)`` returns the scaled CO2 entry value, med_cat being the electroneutral estimate
    (HCO3 + 2 CO32 + OH - H in mol/m3, RD3:589-592) divided by the cation's bulk concentration.
    Returns (history [n_steps + 1, nv, 7], Newton counts, CO2 entry values used)."""
    disc = RxnDiff3D(x, cells, prm, wall_w, exit_facets, exit_area)
    nv = disc.nv
    ev = np.asarray(entry_verts, np.int64)
    bc_dofs = np.concatenate([ev * NS + CO2, ev * NS + CO, ev * NS + H2])
    eq = prm.extras["eq_scaled"]
    co2 = float(eq[0])
    u = np.zeros(disc.ndof)                                # u = Function(V), RD3:377
    un = np.ones(disc.ndof)                                # u_n = interpolate(u_0, V), RD3:381-386
    hist, its, co2s = [un.reshape(nv, NS).copy()], [], []
    c0 = prm.c0
    for _ in range(n_steps):
        vals = np.concatenate([np.full(len(ev), co2), np.full(len(ev), eq[1]), np.full(len(ev), eq[2])])
        co2s.append(co2)
        u, k, conv, r0, r = newton(disc, u, un, bc_dofs, vals)
        if not conv:
            raise RuntimeError("Newton solver did not converge")
        its.append(k)
        U = u.reshape(nv, NS)
        hist.append(U.copy())
        if sechenov is not None:
            m = [float(np.median(U[:, i])) for i in (H, OH, HCO3, CO32)]
            cat = m[2] * c0[HCO3] + 2 * m[3] * c0[CO32] + m[1] * c0[OH] - m[0] * c0[H]
            co2 = sechenov(m[1], m[2], m[3], cat / c0[7])
        un = u.copy()
    return np.array(hist), its, co2s
