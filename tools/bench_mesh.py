#!/usr/bin/env python
"""BASELINE config 5: one synthetic, uniformly refined pore (L_10_R_5 red-refined = the aspect of the missing
L_100_R_50 mesh) partitioned across the GPUs of a box.  Run under torchrun for N > 1:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_mesh.py --refine 2 --gmres-iters 30

Times, per rank on the device and as the max over ranks: assembly (J + F), BSR SpMV including the halo exchange,
and GMRES iterations (SpMV + block-Jacobi + CGS2 with two all-reduces).  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gmpnp_b200 import marking, meshio, params, partition  # noqa: E402
from gmpnp_b200.dist3d import LocalComm, PartitionedPore, TorchComm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mesh", default="L_10_R_5")
ap.add_argument("--L", type=float, default=100e-9)
ap.add_argument("--R", type=float, default=50e-9)
ap.add_argument("--refine", type=int, default=2)
ap.add_argument("--gmres-iters", type=int, default=30)
ap.add_argument("--reps", type=int, default=30)
ap.add_argument("--overlap", type=int, default=-1, help="1/0: force the interior-rows/halo overlap on/off (-1: auto)")
ap.add_argument("--newton", type=int, default=0, help="also time this many damped Newton iterations of the first "
                "reference time step (assembly + preconditioner setup + GMRES to --lin-rtol)")
ap.add_argument("--lin-rtol", type=float, default=1e-8)
ap.add_argument("--coarse", type=int, default=1)
ap.add_argument("--emulate", type=int, default=0, help="emulate this many ranks inside one process (LocalComm)")
a = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)

t0 = time.time()
mesh = meshio.load_mesh(a.mesh)
for _ in range(a.refine):
    mesh = meshio.red_refine(mesh, project_radius=a.R / a.L)
nv, nt = mesh.x.shape[0], mesh.cells.shape[0]
prm = params.params_3d(L=a.L, R=a.R)
dirichlet = marking.dirichlet_sets(mesh, a.L, a.R)
if a.emulate > 1:
    parts = partition.partition_z(mesh, a.emulate)
    comm = LocalComm(parts)
    nparts = a.emulate
else:
    parts = partition.partition_z(mesh, world, ranks=[rank])
    comm = TorchComm(parts[0])
    nparts = world
pp = PartitionedPore(mesh, a.L, a.R, prm, parts, comm, device=local, dirichlet=dirichlet, coarse=bool(a.coarse))
if a.overlap >= 0:
    pp.overlap = bool(a.overlap)
setup_s = time.time() - t0


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, reps):
    for _ in range(5):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / reps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    return ms


rng = np.random.default_rng(0)
ug = np.ones((nv, 9)); ug[:, 8] = 0.0
ug += 0.01 * rng.random((nv, 9))
ung = np.ones((nv, 9)); ung[:, 8] = 0.0
us, uns = pp.from_global(ug), pp.from_global(ung)
Fs, nrm = pp.assemble(us, uns)
ms_asm = timed(lambda: pp.assemble(us, uns), max(2, a.reps // 3))
xs = pp.from_global(rng.normal(size=(nv, 9)))
hb0 = comm.halo_bytes
ms_spmv = timed(lambda: pp.spmv(xs), a.reps)
halo_per_spmv = sum(p.halo_doubles() * 8 for p in pp.parts)      # bytes this process receives per exchange
# SpMV without the exchange (local kernel only)
ms_spmv_local = timed(lambda: [s.spmv(J, x) for s, J, x in zip(pp.solvers, pp.J, xs)], a.reps)
# GMRES iterations: fixed count, one restart cycle
pp.gmres(Fs, m=a.gmres_iters, maxit=a.gmres_iters, rtol=1e-30)        # warm-up (allocations, NCCL channels)
st0 = dict(pp.stats)
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
_, its, rel = pp.gmres(Fs, m=a.gmres_iters, maxit=a.gmres_iters, rtol=1e-30)
e1.record()
barrier()
ms_gmres = e0.elapsed_time(e1)
if world > 1:
    t = torch.tensor([ms_gmres], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_gmres = float(t[0])

newton = None
if a.newton > 0:
    u0 = pp.from_global(ung * 0.0)
    barrier()
    t0 = time.time()
    out = pp.newton(u0, uns, maxit=a.newton, lin_rtol=a.lin_rtol, lin_restart=100, lin_maxit=3000)
    barrier()
    newton = dict(out, wall_s=time.time() - t0)

nb_local = sum(s.n_blocks for s in pp.solvers)
own_rows = sum(p.n_own for p in pp.parts)
ghost = sum(p.n_ghost for p in pp.parts)
cnt = torch.tensor([nb_local, own_rows, ghost, halo_per_spmv], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
nb_tot, own_tot, ghost_tot, halo_tot = [float(v) for v in cnt.tolist()]
if rank == 0:
    # algorithmic bytes: SpMV 8*81*nb + 4*nb + 4*(V+1) + 2*8*9*V ; assembly 8*81*nb + 8*9*V*3 + 8*3*V + 4*4*T + 4*16*T
    b_spmv = 8 * 81 * nb_tot + 4 * nb_tot + 4 * (own_tot + nparts) + 2 * 8 * 9 * own_tot
    b_asm = 8 * 81 * nb_tot + 3 * 8 * 9 * own_tot + 8 * 3 * own_tot + (4 * 4 + 4 * 16) * nt
    peak = 6551.7
    pk = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk)).get("hbm_gbs", peak))
    gpus = max(world, 1)
    line = {"workload": f"config5: {a.mesh} red-refined x{a.refine}, L={a.L:g}, R={a.R:g}, as-executed BCs",
            "n_gpus": world, "parts": nparts, "emulated_in_one_process": bool(a.emulate > 1),
            "vertices": nv, "tets": nt, "dofs": 9 * nv, "bsr_blocks_local_total": int(nb_tot),
            "jacobian_GB": 8 * 81 * nb_tot / 1e9, "ghost_vertices_total": int(ghost_tot),
            "halo_bytes_per_spmv_total": int(halo_tot), "overlap": a.overlap, "setup_s": setup_s,
            "assemble_ms": ms_asm, "assemble_GBs_aggregate": b_asm / ms_asm / 1e6,
            "spmv_ms": ms_spmv, "spmv_local_kernel_ms": ms_spmv_local, "spmv_GBs_aggregate": b_spmv / ms_spmv / 1e6,
            "spmv_frac_of_hbm_peak_per_gpu": b_spmv / ms_spmv / 1e6 / (peak * gpus),
            "gmres_iters": its, "gmres_ms_per_iter": ms_gmres / max(its, 1),
            "allreduces_per_iter": (pp.stats["allreduce"] - st0["allreduce"]) / max(its, 1),
            "coarse_space": bool(a.coarse), "newton": newton, "peak_GBs": peak}
    print(json.dumps(line), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
