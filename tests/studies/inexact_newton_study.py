"""CPU study (test infrastructure; drives the oracle): how accurate must the linear solves of the damped 3D Newton
iteration be?  The reference solves every Newton system directly (MUMPS, 3D:792) but damps the step by 0.9 (3D:796),
so the nonlinear residual contracts by ~0.1 per iteration whatever the linear accuracy; the CUDA path uses GMRES to
1e-8 (``NewtonOpts.sweep_3d``), 87 iterations per Newton step on config 3 (profiles/r01_pore3d_steady_launches.md).

For a list of forcing terms eta the reference march (3D:782-858, oracle ``Discretisation`` + App. C Newton) is run with
GMRES(40) preconditioned exactly like the CUDA path (block-Jacobi on the 9x9 diagonal blocks + additive Galerkin coarse
space on 16 z-slabs x 9 components) and stopped at ||b - A dx|| <= eta ||b||.  Reported per eta: Newton counts per
pseudo-time step, total GMRES iterations, and the distance of every step's solution from the direct-solve march.

    python tests/studies/inexact_newton_study.py [--mesh L_10_R_5 --L 10e-9 --R 5e-9] [--steps 3]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from gmpnp_b200 import marking, meshio, params  # noqa: E402
from oracle import solver  # noqa: E402

NZ, NC = 16, 9


class Precond:
    """z = D^-1 r + P (P^T A P)^-1 P^T r  (csrc/pore3d.cu: bjacobi_invert / coarse_setup / precond_apply)."""

    def __init__(self, A, z_coord):
        n = A.shape[0]
        nv = n // NC
        Ab = A.tobsr(blocksize=(NC, NC))
        Ab.sort_indices()
        D = np.zeros((nv, NC, NC))
        for i in range(nv):
            cols = Ab.indices[Ab.indptr[i]:Ab.indptr[i + 1]]
            D[i] = Ab.data[Ab.indptr[i] + int(np.searchsorted(cols, i))]
        self.Dinv = np.linalg.inv(D)
        zmin, zmax = z_coord.min(), z_coord.max()
        slab = np.clip(((z_coord - zmin) / (zmax - zmin) * NZ).astype(int), 0, NZ - 1)
        rows = np.arange(n)
        cols = np.repeat(slab, NC) * NC + np.tile(np.arange(NC), nv)
        self.P = sp.csr_matrix((np.ones(n), (rows, cols)), shape=(n, NZ * NC))
        Ac = (self.P.T @ A @ self.P).toarray()
        empty = np.abs(Ac).sum(axis=1) == 0
        Ac[empty, empty] = 1.0
        self.Acinv = np.linalg.inv(Ac)
        self.nv = nv

    def __call__(self, r):
        z = np.einsum("vij,vj->vi", self.Dinv, r.reshape(self.nv, NC)).ravel()
        return z + self.P @ (self.Acinv @ (self.P.T @ r))


def newton_inexact(disc, prm, u, un, bc_dofs, bc_vals, z_coord, eta, relax=0.9, rtol=1e-4, atol=1e-4, maxit=50):
    x = u.copy()
    b = solver.apply_bc_residual(disc.residual(x, un, prm), x, bc_dofs, bc_vals)
    r0 = r = float(np.linalg.norm(b))
    k, lin = 0, 0
    conv = r < atol
    while not conv and k < maxit:
        A = solver.apply_bc_matrix(disc.jacobian(x, prm), bc_dofs)
        if eta is None:
            dx = spla.splu(A).solve(b)
        else:
            M = Precond(A.tocsr(), z_coord)
            cnt = [0]

            def cb(_):
                cnt[0] += 1

            # right preconditioning as in the CUDA path: solve (A M) y = b, dx = M y
            op = spla.LinearOperator(A.shape, matvec=lambda y: A @ M(y))
            y, info = spla.gmres(op, b, rtol=eta, atol=0.0, restart=40, maxiter=100, callback=cb, callback_type="pr_norm")
            dx = M(y)
            lin += cnt[0]
        x = x - relax * dx
        k += 1
        b = solver.apply_bc_residual(disc.residual(x, un, prm), x, bc_dofs, bc_vals)
        r = float(np.linalg.norm(b))
        conv = (r / r0 < rtol) or (r < atol)
    return x, k, conv, lin


def march(mesh, prm, dofs, kind, n_steps, eta):
    disc = solver.Discretisation(mesh.x, mesh.cells, NC)
    nv = disc.nv
    eq = prm.extras["eq_scaled"]
    co2 = float(eq[0])
    u = np.zeros(disc.ndof)
    un = np.tile(np.array([1.0] * 8 + [0.0]), nv)
    its, lins, states = [], [], []
    for _ in range(n_steps):
        vals = np.array([0.0, prm.V, co2, eq[1], eq[2]])[kind.astype(np.int64)]
        u, k, conv, lin = newton_inexact(disc, prm, u, un, dofs, vals, mesh.x[:, 2], eta)
        if not conv:
            raise RuntimeError(f"Newton did not converge (eta={eta})")
        its.append(k); lins.append(lin)
        U = u.reshape(nv, NC)
        states.append(U.copy())
        co2 = params.sechenov_co2_scaled(prm, np.median(U[:, 1]), np.median(U[:, 2]), np.median(U[:, 3]), np.median(U[:, 7]))
        un = u.copy()
    return its, lins, states


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", default="L_10_R_5")
    ap.add_argument("--L", type=float, default=10e-9)
    ap.add_argument("--R", type=float, default=5e-9)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--etas", default="1e-8,1e-4,1e-2,5e-2,1e-1")
    a = ap.parse_args()
    mesh = meshio.load_mesh(a.mesh)
    prm = params.params_3d(L=a.L, R=a.R)
    dofs, kind, _ = marking.dirichlet_sets(mesh, a.L, a.R)
    dofs = dofs.astype(np.int64)
    t0 = time.time()
    its0, _, ref = march(mesh, prm, dofs, kind, a.steps, None)
    print(json.dumps({"eta": "direct", "newton": its0, "wall_s": round(time.time() - t0, 1)}), flush=True)
    for eta in [float(e) for e in a.etas.split(",")]:
        t0 = time.time()
        try:
            its, lins, st = march(mesh, prm, dofs, kind, a.steps, eta)
        except RuntimeError as e:
            print(json.dumps({"eta": eta, "failed": str(e)}), flush=True)
            continue
        dist = [float(max(np.linalg.norm(s[:, c] - r[:, c]) / max(np.linalg.norm(r[:, c]), 1e-300) for c in range(NC)))
                for s, r in zip(st, ref)]
        print(json.dumps({"eta": eta, "newton": its, "gmres_total": lins, "gmres_per_newton": round(sum(lins) / sum(its), 1),
                          "max_rel_l2_vs_direct_per_step": dist, "wall_s": round(time.time() - t0, 1)}), flush=True)


if __name__ == "__main__":
    main()
