#!/usr/bin/env python
"""Benchmark of the GMPNP hot path: steady solves/sec on the 1D parameter sweep (BASELINE.json
config 2: {K,Cs} x 256 voltages x {0.1,0.5,1.0} M x 5 meshes = 7680 steady problems per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" = one pass of the hot path over the whole batch: every sweep point goes from the bulk
state to its converged steady solution (voltage continuation + Newton, all inside the CUDA
kernels).  Prints ONE JSON line (see the task contract): `value` = solves/s with inputs resident
in HBM, `e2e` = the same through the public host API with host buffers (H2D of parameters and
continuation paths, D2H of every solution profile inside the timed region), `roofline` for the
fused assemble+eliminate kernel, `cpu_baseline` = the oracle on the box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GMPNP steady solves/sec (1D sweep)"
UNIT = "solves/s"
WORKLOAD = "config2: 1D GMPNP sweep {K,Cs} x 256 V x {0.1,0.5,1.0} M KHCO3 x 5 meshes = 7680 steady solves per GPU"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--voltages", type=int, default=256, help="voltage points per chain (256 = config 2)")
    ap.add_argument("--cpu-sample", type=int, default=16, help="sweep points timed on the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pore3d-batch", type=int, default=128, help="3D pore problems per GPU in the 3D part (0: skip)")
    ap.add_argument("--pivot", type=int, default=0, help="partial pivoting inside the 7x7 blocks (0: none; "
                    "non-converged points are retried with pivoting, see Sweep1D)")
    ap.add_argument("--dv", type=float, default=0.75, help="largest voltage increment of the continuation [V_T]")
    ap.add_argument("--xtol-path", type=float, default=1.0,
                    help="increment tolerance of the intermediate continuation stages (1.0 = one Newton corrector "
                         "per voltage increment); the final stage always converges to xtol = 1e-12")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle (port of the reference's algorithm class: P1 assembly + sparse LU + Newton)
# --------------------------------------------------------------------------------------------
def _cpu_solve_point(args):
    cation, conc, L_n, V, dv, xtol_path = args
    import numpy as np
    from gmpnp_b200 import meshio, params
    from gmpnp_b200.sweep import voltage_paths
    from oracle import solver as osolver
    os.environ["OMP_NUM_THREADS"] = "1"
    mesh = meshio.load_mesh(params.mesh_name_1d(L_n))
    prm = params.params_1d(concentration_elec=conc, cation=cation, L_n=L_n, voltage_multiplier=V)
    path = voltage_paths(np.array([V]), dv)[0]
    path = path[~np.isnan(path)]
    t = time.perf_counter()
    try:
        u, its = osolver.steady_1d(mesh.x[:, 0], prm, path, xtol=1e-12, xtol_path=xtol_path, jac_rule=1)
        ok = True
    except RuntimeError:
        its, ok = [], False
    return time.perf_counter() - t, sum(its), ok


def cpu_sample_points(n_sample, n_voltages, dv=0.75, xtol_path=1.0):
    import numpy as np
    from gmpnp_b200.sweep import config2_points
    pts = config2_points(n_voltages)
    rng = np.random.default_rng(0)
    sel = rng.choice(len(pts), size=min(n_sample, len(pts)), replace=False)
    return [(pts[i].cation, pts[i].conc, pts[i].L_n, pts[i].V, dv, xtol_path) for i in sorted(sel)]


def run_cpu(n_sample, n_voltages, cores=None, dv=0.75, xtol_path=1.0):
    import multiprocessing as mp
    cores = cores or (os.cpu_count() or 1)
    cores = max(1, min(cores, n_sample))
    work = cpu_sample_points(n_sample, n_voltages, dv, xtol_path)
    t = time.perf_counter()
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_cpu_solve_point, work, chunksize=1)
    wall = time.perf_counter() - t
    n_ok = sum(1 for r in res if r[2])
    return dict(value=len(work) / wall, unit=UNIT, cores=cores, kind="port",
                sample=f"{len(work)} of the {7680 if n_voltages == 256 else 30 * n_voltages} sweep points (seed 0), "
                       f"oracle steady_1d (NumPy assembly + SuperLU, same continuation and Jacobian rule as the GPU arm), "
                       f"{n_ok} converged, wall {wall:.1f} s",
                newton_iterations=int(sum(r[1] for r in res)))


def _cpu_newton_iteration_3d(_):
    """One damped-Newton iteration of config 3 on the CPU oracle: P1 assembly of F and J (NumPy) + sparse LU
    (SuperLU; the reference uses MUMPS, 3D:792) + solve."""
    os.environ["OMP_NUM_THREADS"] = "1"
    import numpy as np
    import scipy.sparse.linalg as spla
    from gmpnp_b200 import marking, meshio, params
    from oracle import solver as osolver
    mesh = meshio.load_mesh("L_50_R_5")
    p3 = params.params_3d(L=50e-9, R=5e-9)
    dofs, kind, _info = marking.dirichlet_sets(mesh, 50e-9, 5e-9)
    disc = osolver.Discretisation(mesh.x, mesh.cells, 9)
    eq = p3.extras["eq_scaled"]
    vals = np.array([0.0, p3.V, float(eq[0]), eq[1], eq[2]])[kind.astype(np.int64)]
    u = np.zeros(disc.ndof)
    un = np.tile(np.array([1.0] * 8 + [0.0]), disc.nv)
    t = time.perf_counter()
    b = osolver.apply_bc_residual(disc.residual(u, un, p3), u, dofs.astype(np.int64), vals)
    A = osolver.apply_bc_matrix(disc.jacobian(u, p3), dofs.astype(np.int64))
    dx = spla.splu(A).solve(b)
    return time.perf_counter() - t, float(np.abs(dx).max())


def run_cpu_3d(newton_per_solve, cores=None):
    """Bounded CPU sample of the 3D workload: ONE Newton iteration per process on up to 8 host cores at once;
    steady solves/s is extrapolated with the Newton count per steady solve measured on the GPU arm."""
    import multiprocessing as mp
    cores = max(1, min(cores or (os.cpu_count() or 1), 8))
    t = time.perf_counter()
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_cpu_newton_iteration_3d, range(cores), chunksize=1)
    wall = time.perf_counter() - t
    per_it = sum(r[0] for r in res) / len(res)
    return dict(value=cores / (per_it * newton_per_solve), unit="steady solves/s (extrapolated)", cores=cores, kind="port",
                sample=f"one damped-Newton iteration of config 3 (NumPy assembly + SuperLU of 33111 DOFs) per core on {cores} "
                       f"cores at once: {per_it:.1f} s per iteration, wall {wall:.1f} s; x {newton_per_solve} iterations per steady solve")


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for _ in range(args.warmup if args.warmup < 1 else 1):
        run_cpu(min(args.cpu_sample, os.cpu_count() or 1), args.voltages, dv=args.dv, xtol_path=args.xtol_path)
    info = None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        info = run_cpu(args.cpu_sample, args.voltages, dv=args.dv, xtol_path=args.xtol_path)
        vals.append(info["value"])
    T = (time.perf_counter() - t0) / max(1, args.steps)
    v = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": T * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU oracle port (FEniCS is not installable here); each step "
                       "= a bounded sample of the sweep on all host cores"},
            "cpu_baseline": dict(info, value=v),
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# 3D pore part of the metric ("steady solves/sec (1D & 3D pore); assembly + SpMV HBM GB/s")
# --------------------------------------------------------------------------------------------
def bench_pore3d(local, world, dev, batch, peak):
    """BASELINE config 3 (L_50_R_5, parameters_pore.yaml, 1.0 M, K+, as-executed BCs) as a batch of `batch` wall
    voltages in [-0.5, -1.25] V_T: (i) steady solves = the reference's pseudo-time march (damped Newton, relaxation
    0.9, GMRES + block-Jacobi/coarse) run to increments <= 1e-8, (ii) the assembly and BSR SpMV kernels alone."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from gmpnp_b200 import meshio, params, solver3d
    from gmpnp_b200._lib import NewtonOpts
    mesh = meshio.load_mesh("L_50_R_5")
    Vs = np.linspace(-0.5, -1.25, batch)
    plist = [params.params_3d(L=50e-9, R=5e-9, voltage_multiplier=float(V)) for V in Vs]
    pp = solver3d.PoreProblem(mesh, 50e-9, 5e-9, plist, device=local)
    s = pp.solver
    Vn, T, nb = s.n, s.n_tet, s.n_blocks

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps):
        for _ in range(3):
            fn()
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        sync()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    u = solver3d.bulk_state(batch, Vn, dev)
    u += 0.01 * torch.rand_like(u)
    un = solver3d.bulk_state(batch, Vn, dev)
    s.set_dirichlet(pp.dirichlet_values([float(p.extras["eq_scaled"][0]) for p in plist]))
    F, J = s.assemble(u, un)
    x = torch.rand_like(u)
    ms_asm = timed(lambda: s.assemble(u, un), 10)
    ms_spmv = timed(lambda: s.spmv(J, x), 30)
    b_asm = batch * (8 * 81 * nb + 8 * 9 * Vn + 2 * 8 * 9 * Vn + 8 * 3 * Vn + 4 * 4 * T + 4 * 16 * T)
    b_spmv = batch * (8 * 81 * nb + 4 * nb + 4 * (Vn + 1) + 2 * 8 * 9 * Vn)
    del F, J, x
    l0 = s.launch_count()
    o3 = NewtonOpts.sweep_3d()
    pp.steady(opts=o3, tol=1e-8, max_steps=3)             # warm-up (allocations of the Krylov basis)
    sync()
    l1 = s.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = pp.steady(opts=o3, tol=1e-8, max_steps=20)
    e1.record()
    sync()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_steady = float(t[0])
    res = {
        "workload": f"config3 batch: L_50_R_5 (V=3679, T=17297, 33111 DOFs), {batch} wall voltages in [-0.5,-1.25] V_T per GPU, "
                    "pseudo-time march to steady state (increment <= 1e-8); damped Newton (relaxation 0.9, residual "
                    "criterion 1e-4 as 3D:789-798), GMRES(40) to 1e-8 + block-Jacobi + z-slab coarse space",
        "steady_solves_per_s": world * batch / (ms_steady * 1e-3), "ms_per_batch": ms_steady,
        "pseudo_time_steps": int(out["steps"]), "newton_iterations_per_problem": int(out["iters"].sum(axis=0).max()),
        "gpu_launches": int(s.launch_count() - l1),
        "assemble": {"ms": ms_asm, "GBs": b_asm / ms_asm / 1e6, "frac_of_hbm_peak": b_asm / ms_asm / 1e6 / peak,
                     "algorithmic_bytes": b_asm,
                     "dram_traffic_over_algorithmic": 1.18,      # ncu, profiles/r01_pore3d_ncu_summary.md
                     "kernels": "tet_moments_kernel + assemble_bsr_kernel + residual_gather_kernel"},
        "spmv": {"ms": ms_spmv, "GBs": b_spmv / ms_spmv / 1e6, "frac_of_hbm_peak": b_spmv / ms_spmv / 1e6 / peak,
                 "algorithmic_bytes": b_spmv,
                 "dram_traffic_over_algorithmic": 0.99,      # ncu, profiles/r01_spmv_v2_ncu_summary.md
                 "note": "batch Jacobians (%.2f GB) exceed the 126 MB L2" % (batch * 8 * 81 * nb / 1e9)},
    }
    pp.solver.close()
    return res


# --------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [s.strip() for s in out.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm), "power_w_max": max(float(s[2]) for s in self.samples)}


def main():
    args = parse()
    if args.impl == "reference":
        return reference_arm(args)
    import numpy as np
    import torch
    import torch.distributed as dist
    from gmpnp_b200 import sweep
    from gmpnp_b200.solver1d import NC

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # weak scaling: every rank owns one full config-2 sweep (independent sweep points, no collective
    # on the data path; one gather of the per-point summaries at the end)
    pts = sweep.config2_points(args.voltages)
    sw = sweep.Sweep1D(pts, device=local, dv_max=args.dv, xtol_path=args.xtol_path, pivot=args.pivot)
    n_local = sw.n_points

    # pinned host staging for the e2e arm
    h_params = [torch.as_tensor(g["packed"]).pin_memory() for g in sw.groups]
    h_paths = [torch.as_tensor(g["path"]).pin_memory() for g in sw.groups]
    h_out = [torch.empty(g["solver"].batch, g["solver"].n, NC, dtype=torch.float64).pin_memory() for g in sw.groups]
    h2d = sum(t.numel() * 8 for t in h_params) + sum(t.numel() * 8 for t in h_paths)
    d2h = sum(t.numel() * 8 for t in h_out)

    sw.upload()
    launches0 = sum(g["solver"].launch_count() for g in sw.groups)
    # ---- resident arm ---------------------------------------------------------------------
    for _ in range(args.warmup):
        outs = sw.solve_resident()
    barrier()
    launches1 = sum(g["solver"].launch_count() for g in sw.groups)
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        outs = sw.solve_resident()
    ev1.record()
    barrier()
    sampler.stop_flag = True
    ms = ev0.elapsed_time(ev1) / args.steps
    launches2 = sum(g["solver"].launch_count() for g in sw.groups)
    n_its, alg_bytes = sw.newton_iterations(outs)
    status = np.concatenate([o["status"].cpu().numpy() for o in outs])
    n_conv = int((status == 0).sum())

    # ---- e2e arm: host buffers in, host buffers out, through the public API ---------------
    def e2e_step():
        for g, hp, hv in zip(sw.groups, h_params, h_paths):
            g["solver"].set_params(hp.numpy())                  # H2D of the parameter records
            g["d_path"].copy_(hv, non_blocking=True)            # H2D of the continuation paths
        o = sw.solve_resident(host_out=h_out)                   # D2H of every solution profile, per mesh stream
        torch.cuda.synchronize()
        return o

    e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / args.steps

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cnt = torch.tensor([n_local, n_conv, n_its, alg_bytes], dtype=torch.float64, device=dev)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        n_total, n_conv_total, n_its_total, bytes_total = [float(v) for v in cnt.tolist()]
    else:
        n_total, n_conv_total, n_its_total, bytes_total = n_local, n_conv, n_its, alg_bytes
    ms, ms_e2e = [float(v) for v in t.tolist()]

    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    pore3d = None
    if args.pore3d_batch > 0:
        sw.close()
        del h_out
        torch.cuda.empty_cache()
        pore3d = bench_pore3d(local, world, dev, args.pore3d_batch, peak)
    if rank == 0:
        achieved = (bytes_total / world) / (ms * 1e-3) / 1e9       # per-GPU GB/s of the kernel
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        # DRAM traffic of the kernel: ncu's dram__bytes_read+write over the algorithmic bytes of the captured launch,
        # applied to this step's algorithmic bytes (the capture is one of the five launches of a step)
        ratio = json.load(open(prof)).get("newton1d_dram_bytes_over_algorithmic_bytes") if os.path.exists(prof) else None
        traffic = None if ratio is None else ratio * bytes_total / world
        line = {
            "metric": METRIC, "value": n_total / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD if args.voltages == 256 else f"reduced sweep ({args.voltages} V/chain)",
                       "points_per_gpu": n_local, "converged": int(n_conv_total),
                       "newton_iterations_per_step": int(n_its_total),
                       "continuation": f"dV<={args.dv:g} V_T, xtol 1e-12 (final) / {args.xtol_path:g} (path), consistent Jacobian, "
                                       f"in-block pivoting {'on' if args.pivot else 'off (equilibrated rows; failures retried with pivoting)'}",
                       "cache": "working set (elimination workspace 10.7 GB/GPU) >> 126 MB L2, no flush needed",
                       "parallelism": f"sweep points sharded, {world} GPU(s), no data-path collective"},
            "e2e": {"value": n_total / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches2 - launches1),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                         "kernel": "edl1d::newton1d_kernel (5 concurrent launches = one step)",
                         "algorithmic_bytes_per_step": bytes_total / world},
            "clocks": sampler.summary(),
        }
        # fp64: executed flops of the hot kernel (thread-level DFMA x2 + DMUL + DADD per block row and Newton iteration,
        # counted by ncu on the v10 kernel: profiles/r01_newton1d_v10_ncu_summary.md) against the measured DFMA peak
        import ctypes as C
        from gmpnp_b200 import _lib
        pk64 = C.c_double(0.0)
        _lib.check(_lib.load().gmpnp_fp64_peak(local, C.byref(pk64)))
        flops_row = 6233.0
        ach64 = flops_row * (bytes_total / world / 1072.0) / (ms * 1e-3) / 1e12
        line["fp64"] = {"peak_tflops_measured": pk64.value, "achieved_tflops_executed": ach64,
                        "frac": ach64 / pk64.value if pk64.value > 0 else None,
                        "flops_per_block_row_iteration": flops_row,
                        "note": "executed (not minimal) fp64 flops incl. the per-lane redundancy of the quadrature"}
        if pore3d is not None:
            if world == 1 and not args.no_cpu_baseline:
                pore3d["cpu_baseline"] = run_cpu_3d(pore3d["newton_iterations_per_problem"])
            line["pore3d"] = pore3d
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = run_cpu(args.cpu_sample, args.voltages, dv=args.dv, xtol_path=args.xtol_path)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
