"""3D diagnostics: kernel timings (assembly, SpMV) with achieved GB/s vs the algorithmic bytes of SURVEY 8d,
Newton/GMRES iteration counts for config 3, pseudo-time convergence.

    python tools/prof_3d.py [--mesh L_50_R_5] [--batch 1] [--steps 3] [--refine 0]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gmpnp_b200 import meshio, params, solver3d  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mesh", default="L_50_R_5")
ap.add_argument("--L", type=float, default=50e-9)
ap.add_argument("--R", type=float, default=5e-9)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--refine", type=int, default=0)
ap.add_argument("--steady", action="store_true")
a = ap.parse_args()

mesh = meshio.load_mesh(a.mesh)
for _ in range(a.refine):
    mesh = meshio.red_refine(mesh, project_radius=a.R / a.L)
Vs = -0.5 - 0.75 * np.arange(a.batch) / max(1, a.batch - 1)          # the bench's voltage range [-0.5, -1.25] V_T
plist = [params.params_3d(L=a.L, R=a.R, voltage_multiplier=float(V)) for V in Vs]
t0 = time.time()
pp = solver3d.PoreProblem(mesh, a.L, a.R, plist)
s = pp.solver
print(f"mesh {a.mesh} refine {a.refine}: V={s.n} T={s.n_tet} blocks={s.n_blocks} batch={a.batch} setup {time.time()-t0:.2f}s",
      flush=True)
dev = s.device
u = solver3d.bulk_state(a.batch, s.n, dev)
u += 0.01 * torch.rand_like(u)
un = solver3d.bulk_state(a.batch, s.n, dev)
s.set_dirichlet(pp.dirichlet_values([float(p.extras["eq_scaled"][0]) for p in plist]))


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


V, T, nb, B = s.n, s.n_tet, s.n_blocks, a.batch
F, J = s.assemble(u, un)
x = torch.rand_like(u)
ms_asm = timeit(lambda: s.assemble(u, un))
ms_res = timeit(lambda: s.assemble(u, un, want_J=False))
ms_spmv = timeit(lambda: s.spmv(J, x), reps=50)
b_asm = B * (8 * 81 * nb + 8 * 9 * V + 2 * 8 * 9 * V + 8 * 3 * V + 4 * 4 * T + 4 * 16 * T)
b_spmv = B * (8 * 81 * nb + 4 * nb + 4 * (V + 1) + 2 * 8 * 9 * V)
print(f"assemble J+F: {ms_asm:.3f} ms  -> {b_asm / ms_asm / 1e6:.1f} GB/s algorithmic ({b_asm / 1e6:.1f} MB)")
print(f"residual only: {ms_res:.3f} ms")
print(f"BSR SpMV: {ms_spmv:.4f} ms -> {b_spmv / ms_spmv / 1e6:.1f} GB/s algorithmic ({b_spmv / 1e6:.1f} MB)", flush=True)
if a.steady:
    t0 = time.time()
    out = pp.steady(tol=1e-9, max_steps=80)
    torch.cuda.synchronize()
    print(f"pseudo-time steady: {out['steps']} steps in {time.time()-t0:.2f}s, Newton its/step {out['iters'][:, 0].tolist()}")
    print("increments", np.array2string(out["increments"], precision=2))
elif a.steps:
    t0 = time.time()
    s.set_params(plist)
    s.set_march_data(*pp.march_data())
    u0 = torch.zeros(a.batch, s.n, 9, dtype=torch.float64, device=dev)
    out = s.march(u0, solver3d.bulk_state(a.batch, s.n, dev), a.steps)          # per-problem status, never raises
    torch.cuda.synchronize()
    dt = time.time() - t0
    st = out["status"].cpu().numpy()
    print(f"march {a.steps} steps: {dt:.2f}s, converged {int((st == 0).sum())}/{a.batch}, Newton its (problem 0) "
          f"{out['iters'][0].tolist()}, GMRES its (problem 0) {out['lin_iters'][0].tolist()}, launches {s.launch_count()}")
