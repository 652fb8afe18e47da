"""Element residual and Jacobian of the GMPNP weak forms (oracle; test infrastructure).

Restates, for P1 elements on intervals and tetrahedra, the forms the reference builds in UFL:

* reactions R_i           1D/MPNP_CO2ER_EDL.py:383-410  == 3D/MPNP_CO2ER_pore.py:505-532
* Poisson F_p             1D:412-427                    == 3D:752-767
* species F_i (MPNP)      1D:457-595                    == 3D:534-750 (volume terms only, as executed)
* Jacobian                the Gateaux derivative dolfin's ``solve(F == 0, u, bcs)`` forms implicitly
                          (1D:737-742, 3D:789-799); entries written out in SURVEY App. A.2.

Unknown layout per cell: ``U[c, a, i]`` with a = local vertex, i = component
(species in the reference's order, potential last).  ``prm`` is a
``gmpnp_b200.params.ProblemParams``-like object (z, nu, c0, s, rate, q, eps_w,
n_water_H, n_water_cat, kappa).  All functions are dtype-generic so that the
complex-step derivative test can run through them.
"""
import numpy as np


def geometry(x, cells):
    """P1 geometry: gradients of the barycentric functions ``g[c, a, d]`` and cell volumes."""
    X = x[cells]                       # [nc, nv, dim]
    nc, nv, dim = X.shape
    if dim == 1:
        h = X[:, 1, 0] - X[:, 0, 0]
        g = np.zeros((nc, 2, 1))
        g[:, 0, 0] = -1.0 / h
        g[:, 1, 0] = 1.0 / h
        return g, np.abs(h)
    Jm = X[:, 1:, :] - X[:, :1, :]     # rows = edge vectors
    det = np.linalg.det(Jm)
    Jinv = np.linalg.inv(Jm)           # grad lambda_{a}, a>=1 = columns of Jinv
    g = np.zeros((nc, nv, dim))
    g[:, 1:, :] = np.transpose(Jinv, (0, 2, 1))
    g[:, 0, :] = -g[:, 1:, :].sum(axis=1)
    return g, np.abs(det) / {2: 2.0, 3: 6.0}[dim]


def _coeffs(prm):
    c0 = np.asarray(prm.c0, dtype=np.float64)
    r = prm.rate
    k = dict(kW=r["kw2"] * c0[0] * c0[1], kA=r["ka1"] * c0[1] * c0[2], kB=r["kb1"] * c0[4] * c0[1],
             kA2=r["ka2"] * c0[3], kB2=r["kb2"] * c0[2], kw1=r["kw1"])
    epsH = prm.n_water_H * c0[0] * 1.0e-3
    epsC = prm.n_water_cat * c0[-1] * 1.0e-3
    return k, epsH, epsC


def minus_R(uq, prm):
    """-R_i(u) at quadrature points for i = H, OH, HCO3, CO32, CO2 (1D:383-410).  uq[..., ncomp]."""
    k, _, _ = _coeffs(prm)
    s = prm.s
    w = k["kW"] * uq[..., 0] * uq[..., 1]
    a = k["kA"] * uq[..., 1] * uq[..., 2]
    b = k["kB"] * uq[..., 4] * uq[..., 1]
    a2 = k["kA2"] * uq[..., 3]
    b2 = k["kB2"] * uq[..., 2]
    out = [s[0] * (w - k["kw1"]),
           s[1] * (w + a + b - k["kw1"] - a2 - b2),
           s[2] * (a + b2 - a2 - b),
           s[3] * (a2 - a),
           s[4] * (b - b2)]
    return np.stack(out, axis=-1)


def minus_dR(uq, prm):
    """d(-R_i)/du_j at quadrature points, shape [..., 5, ncomp]."""
    k, _, _ = _coeffs(prm)
    s = prm.s
    ncomp = uq.shape[-1]
    d = np.zeros(uq.shape[:-1] + (5, ncomp), dtype=uq.dtype)
    uH, uOH, uHCO3, uCO2 = uq[..., 0], uq[..., 1], uq[..., 2], uq[..., 4]
    # w = kW uH uOH ; a = kA uOH uHCO3 ; b = kB uCO2 uOH ; a2 = kA2 uCO32 ; b2 = kB2 uHCO3
    dw = {0: k["kW"] * uOH, 1: k["kW"] * uH}
    da = {1: k["kA"] * uHCO3, 2: k["kA"] * uOH}
    db = {4: k["kB"] * uOH, 1: k["kB"] * uCO2}
    da2 = {3: k["kA2"]}
    db2 = {2: k["kB2"]}

    def acc(row, scale, terms):
        for sign, t in terms:
            for j, v in t.items():
                d[..., row, j] = d[..., row, j] + sign * scale * v
    acc(0, s[0], [(1, dw)])
    acc(1, s[1], [(1, dw), (1, da), (1, db), (-1, da2), (-1, db2)])
    acc(2, s[2], [(1, da), (1, db2), (-1, da2), (-1, db)])
    acc(3, s[3], [(1, da2), (-1, da)])
    acc(4, s[4], [(1, db), (-1, db2)])
    return d


def eps_r(uq, prm):
    """Concentration dependent relative permittivity (1D:413-420)."""
    _, epsH, epsC = _coeffs(prm)
    w = epsC * uq[..., -2] + epsH * uq[..., 0]
    return prm.eps_w * ((55 - w) / 55) + 6 * (w / 55)


def element_residual(U, Un, g, vol, prm, rule):
    """F_e[c, a, i] with quadrature ``rule`` = (lam[q, nv], w[q])."""
    lam, wq = rule
    ns = U.shape[2] - 1
    z = np.asarray(prm.z, dtype=np.float64)
    nu = np.asarray(prm.nu, dtype=np.float64)
    zc0 = z * np.asarray(prm.c0)
    uq = np.einsum("qa,cai->cqi", lam, U)
    unq = np.einsum("qa,cai->cqi", lam, Un)
    gu = np.einsum("cai,cad->cid", U, g)                     # grad of every component
    gp = gu[:, ns, :]
    W = wq[None, :] * vol[:, None]                           # [c, q]
    S = np.einsum("cqi,i->cq", uq[..., :ns], nu)
    Dq = 1.0 / (1.0 - S)
    G = np.einsum("cid,i->cd", gu[:, :ns, :], nu)
    F = np.zeros(U.shape, dtype=uq.dtype)
    # time + diffusion + migration + steric
    F[:, :, :ns] += prm.kappa * np.einsum("cq,cqi,qa->cai", W, uq[..., :ns] - unq[..., :ns], lam)
    F[:, :, :ns] += vol[:, None, None] * np.einsum("cid,cad->cai", gu[:, :ns, :], g)
    gpa = np.einsum("cd,cad->ca", gp, g)
    F[:, :, :ns] += gpa[:, :, None] * (z[None, None, :] * np.einsum("cq,cqi->ci", W, uq[..., :ns])[:, None, :])
    Ga = np.einsum("cd,cad->ca", G, g)
    F[:, :, :ns] += Ga[:, :, None] * np.einsum("cq,cqi->ci", W, uq[..., :ns] * Dq[..., None])[:, None, :]
    # reactions
    F[:, :, :5] += np.einsum("cq,cqi,qa->cai", W, minus_R(uq, prm), lam)
    # Poisson
    F[:, :, ns] += -gpa * np.einsum("cq,cq->c", W, eps_r(uq, prm))[:, None]
    rho = np.einsum("cqi,i->cq", uq[..., :ns], zc0) * prm.q
    F[:, :, ns] += np.einsum("cq,cq,qa->ca", W, rho, lam)
    return F


def element_jacobian(U, g, vol, prm, rule):
    """J_e[c, a, i, b, j] = dF_e[c,a,i]/dU[c,b,j] integrated with ``rule`` (App. A.2)."""
    lam, wq = rule
    nc, nv, ncomp = U.shape
    ns = ncomp - 1
    z = np.asarray(prm.z, dtype=np.float64)
    nu = np.asarray(prm.nu, dtype=np.float64)
    zc0 = z * np.asarray(prm.c0)
    _, epsH, epsC = _coeffs(prm)
    uq = np.einsum("qa,cai->cqi", lam, U)
    gu = np.einsum("cai,cad->cid", U, g)
    gp = gu[:, ns, :]
    W = wq[None, :] * vol[:, None]
    S = np.einsum("cqi,i->cq", uq[..., :ns], nu)
    Dq = 1.0 / (1.0 - S)
    G = np.einsum("cid,i->cd", gu[:, :ns, :], nu)
    K = np.einsum("cad,cbd->cab", g, g) * vol[:, None, None]      # int grad a . grad b
    M = np.einsum("cq,qa,qb->cab", W, lam, lam)                   # int phi_a phi_b
    m = np.einsum("cq,qb->cb", W, lam)                            # int phi_b
    gpa = np.einsum("cd,cad->ca", gp, g)
    Ga = np.einsum("cd,cad->ca", G, g)
    J = np.zeros((nc, nv, ncomp, nv, ncomp), dtype=uq.dtype)
    eye = np.eye(ns)
    # delta_ij [ kappa M + K + z_i (gp.grad a) m_b ]
    J[:, :, :ns, :, :ns] += eye[None, None, :, None, :] * (prm.kappa * M + K)[:, :, None, :, None]
    J[:, :, :ns, :, :ns] += (eye * z[:, None])[None, None, :, None, :] * (gpa[:, :, None] * m[:, None, :])[:, :, None, :, None]
    # reactions
    J[:, :, :5, :, :] += np.einsum("cq,cqij,qa,qb->caibj", W, minus_dR(uq, prm), lam, lam)
    # steric:  (G.grad a) int [delta_ij phi_b D + u_i nu_j phi_b D^2]  +  nu_j K_ab/vol int u_i D
    mD = np.einsum("cq,cq,qb->cb", W, Dq, lam)
    J[:, :, :ns, :, :ns] += eye[None, None, :, None, :] * (Ga[:, :, None] * mD[:, None, :])[:, :, None, :, None]
    mUD2 = np.einsum("cq,cqi,cq,qb->cib", W, uq[..., :ns], Dq * Dq, lam)
    J[:, :, :ns, :, :ns] += np.einsum("ca,cib,j->caibj", Ga, mUD2, nu)
    iUD = np.einsum("cq,cqi,cq->ci", W, uq[..., :ns], Dq)
    J[:, :, :ns, :, :ns] += np.einsum("cab,ci,j->caibj", K / vol[:, None, None], iUD, nu)
    # dF_i/dp
    iU = np.einsum("cq,cqi->ci", W, uq[..., :ns])
    J[:, :, :ns, :, ns] += np.einsum("cab,ci,i->caib", K / vol[:, None, None], iU, z)
    # dF_p/du_j and dF_p/dp
    deps = np.zeros(ns)
    deps[0] = (6.0 - prm.eps_w) / 55.0 * epsH
    deps[ns - 1] = (6.0 - prm.eps_w) / 55.0 * epsC
    J[:, :, ns, :, :ns] += -np.einsum("ca,cb,j->cabj", gpa, m, deps)
    J[:, :, ns, :, :ns] += prm.q * np.einsum("cab,j->cabj", M, zc0)
    iE = np.einsum("cq,cq->c", W, eps_r(uq, prm))
    J[:, :, ns, :, ns] += -(K / vol[:, None, None]) * iE[:, None, None]
    return J
