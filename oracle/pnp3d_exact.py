"""Second, independent CPU restatement of the 3D pore forms (test infrastructure; parity unpinned -- FEniCS is absent).

The 9-component problem of ``3D/MPNP_CO2ER_pore.py`` WITHOUT the Bikerman steric term (all ``scale_vol`` = 0, i.e. the
Poisson-Nernst-Planck limit of the forms 3D:503-769), written directly from the reference's forms with EXACT monomial
integrals on affine P1 tetrahedra instead of quadrature tables:

    int l_a = vol/4,   int l_a l_b = vol (1 + d_ab)/20,   int l_a l_b l_c = vol {6, 2, 1}/120 (1, 2, 3 distinct indices)

Every ``dx`` integrand of that limit is a polynomial of degree <= 3 (time term 2, diffusion 0, migration ``z_i u_i
grad(p).grad(v)`` 1, bilinear reactions 3, ``eps_r(u) grad(p).grad(v)`` 1, space charge 2), and FFC selects rules that
are exact for the estimated degree, so whatever tetrahedron scheme FIAT uses, dolfin's assembled F and J for nu = 0 equal
these integrals to round-off.  This pins -- independently of ``oracle/forms.py`` and ``oracle/quadrature.py`` -- every
term of the 3D residual and Jacobian except the rational steric term, whose value depends on the quadrature rule (FIAT's
degree-3 / degree-4 schemes, restated in ``oracle/quadrature.py``; [upstream]).

Follows: species forms 3D:534-750 (volume terms, as executed), reaction sources 3D:505-532, Poisson form 3D:752-767;
species order H, OH, HCO3, CO32, CO2, CO, H2, cation, potential last (3D:138, 407).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

NS, NC = 8, 9
H, OH, HCO3, CO32, CO2, CO, H2, CAT, P = range(9)


def _geometry(x, cells):
    X = x[cells]
    E = np.concatenate([X, np.ones((len(cells), 4, 1))], axis=2)
    inv = np.linalg.inv(E)
    g = np.transpose(inv[:, :3, :], (0, 2, 1))             # [c, a, d] = grad lambda_a
    vol = np.abs(np.linalg.det(E)) / 6.0
    return g, vol


def _triple():
    T = np.empty((4, 4, 4))
    for a in range(4):
        for b in range(4):
            for c in range(4):
                T[a, b, c] = {1: 6.0, 2: 2.0, 3: 1.0}[len({a, b, c})] / 120.0
    return T


class Pnp3DExact:
    def __init__(self, x, cells, prm):
        self.x = np.asarray(x, float)
        self.cells = np.asarray(cells, np.int64)
        self.nv = self.x.shape[0]
        self.ndof = NC * self.nv
        self.g, self.vol = _geometry(self.x, self.cells)
        self.T = _triple()
        self.M = (np.ones((4, 4)) + np.eye(4)) / 20.0
        self.K = np.einsum("cad,cbd->cab", self.g, self.g) * self.vol[:, None, None]
        self.kappa = float(prm.kappa)                      # 1 / del_t (3D:534)
        c0, k = prm.c0, prm.rate
        self.c0, self.z, self.q = c0, prm.z, float(prm.q)
        self.s = prm.extras["scale_R"]
        self.k_w2 = k["kw2"] * c0[H] * c0[OH]
        self.k_w1 = k["kw1"]
        self.k_a1 = k["ka1"] * c0[OH] * c0[HCO3]
        self.k_a2 = k["ka2"] * c0[CO32]
        self.k_b1 = k["kb1"] * c0[CO2] * c0[OH]
        self.k_b2 = k["kb2"] * c0[HCO3]
        self.eps_w = float(prm.eps_w)
        # w = (n_cat c0_cat u_cat + n_H c0_H u_H) 1e-3 ; eps_r = eps_w (55 - w)/55 + 6 w/55   (3D:752-761)
        self.wH = prm.n_water_H * c0[H] * 1.0e-3
        self.wC = prm.n_water_cat * c0[CAT] * 1.0e-3
        d = (self.cells[:, :, None] * NC + np.arange(NC)[None, None, :]).reshape(len(self.cells), 4 * NC)
        self.cell_dofs = d
        self.rows = np.repeat(d, 4 * NC, axis=1).ravel()
        self.cols = np.tile(d, (1, 4 * NC)).ravel()

    def _P(self, f, g):
        return np.einsum("abd,cb,cd->ca", self.T, f, g) * self.vol[:, None]

    def _L(self, f):
        return np.einsum("ab,cb->ca", self.M, f) * self.vol[:, None]

    def residual(self, u, un):
        U = u.reshape(self.nv, NC)[self.cells]
        Un = un.reshape(self.nv, NC)[self.cells]
        Fe = np.zeros_like(U)
        Kp = np.einsum("cab,cb->ca", self.K, U[:, :, P])               # int grad(p).grad(l_a)
        for i in range(NS):
            Fe[:, :, i] = (self.kappa * self._L(U[:, :, i] - Un[:, :, i]) + np.einsum("cab,cb->ca", self.K, U[:, :, i])
                           + self.z[i] * U[:, :, i].mean(axis=1)[:, None] * Kp)
        C = 0.25 * self.vol[:, None] * np.ones((1, 4))
        Pw = self._P(U[:, :, H], U[:, :, OH])
        Pa = self._P(U[:, :, OH], U[:, :, HCO3])
        Pb = self._P(U[:, :, CO2], U[:, :, OH])
        L3, Lh = self._L(U[:, :, CO32]), self._L(U[:, :, HCO3])
        s = self.s
        Fe[:, :, H] += s[H] * (self.k_w2 * Pw - self.k_w1 * C)
        Fe[:, :, OH] += s[OH] * (self.k_w2 * Pw + self.k_a1 * Pa + self.k_b1 * Pb - self.k_w1 * C - self.k_a2 * L3
                                 - self.k_b2 * Lh)
        Fe[:, :, HCO3] += s[HCO3] * (self.k_a1 * Pa + self.k_b2 * Lh - self.k_a2 * L3 - self.k_b1 * Pb)
        Fe[:, :, CO32] += s[CO32] * (self.k_a2 * L3 - self.k_a1 * Pa)
        Fe[:, :, CO2] += s[CO2] * (self.k_b1 * Pb - self.k_b2 * Lh)
        # Poisson (3D:752-767)
        w = self.wC * U[:, :, CAT].mean(axis=1) + self.wH * U[:, :, H].mean(axis=1)
        eps = self.eps_w * (55.0 - w) / 55.0 + 6.0 * w / 55.0
        rho = sum(self.z[j] * self.c0[j] * U[:, :, j] for j in (H, OH, HCO3, CO32, CAT))
        Fe[:, :, P] = -eps[:, None] * Kp + self.q * self._L(rho)
        F = np.zeros(self.ndof)
        np.add.at(F, self.cell_dofs.ravel(), Fe.reshape(len(self.cells), -1).ravel())
        return F

    def jacobian(self, u):
        U = u.reshape(self.nv, NC)[self.cells]
        nc = len(self.cells)
        Je = np.zeros((nc, 4, NC, 4, NC))
        Mv = self.M[None] * self.vol[:, None, None]
        Kp = np.einsum("cab,cb->ca", self.K, U[:, :, P])
        lin = self.kappa * Mv + self.K
        for i in range(NS):
            Je[:, :, i, :, i] += lin
            if self.z[i] != 0.0:
                Je[:, :, i, :, i] += self.z[i] * 0.25 * Kp[:, :, None] * np.ones((1, 1, 4))
                Je[:, :, i, :, P] += self.z[i] * U[:, :, i].mean(axis=1)[:, None, None] * self.K

        def dP(g):
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
        # Poisson row
        w = self.wC * U[:, :, CAT].mean(axis=1) + self.wH * U[:, :, H].mean(axis=1)
        eps = self.eps_w * (55.0 - w) / 55.0 + 6.0 * w / 55.0
        add(P, P, -eps[:, None, None] * self.K)
        deps = (6.0 - self.eps_w) / 55.0
        for j, wj in ((H, self.wH), (CAT, self.wC)):
            add(P, j, -(deps * wj * 0.25) * Kp[:, :, None] * np.ones((1, 1, 4)))
        for j in (H, OH, HCO3, CO32, CAT):
            add(P, j, self.q * self.z[j] * self.c0[j] * Mv)
        return sp.coo_matrix((Je.ravel(), (self.rows, self.cols)), shape=(self.ndof, self.ndof)).tocsr()
