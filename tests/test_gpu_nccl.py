"""The multi-process paths on NCCL (2 ranks, one GPU each): the sharded sweep's one collective
(`sweep.gather_results`) and the mesh-partitioned 3D solve with `TorchComm` (halo exchange by batched isend/irecv,
all-reduced dot products).  Skipped on a box with fewer than two GPUs; the host logic of both is covered on gloo by
tests/test_sweep_host.py and tests/test_partition_host.py, the numerics of the partitioned mode on one GPU by
tests/test_gpu_dist3d.py."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from gmpnp_b200 import meshio, params, partition, solver3d, sweep
        from gmpnp_b200._lib import NewtonOpts
        from gmpnp_b200.dist3d import PartitionedPore, TorchComm
        # ---- 1. sharded 1D sweep + the final gather over NCCL ------------------------------------------------
        pts = sweep.config2_points(4, meshes=(1e-6, 5e-6))
        mine = sweep.shard(pts, rank, world)
        sw = sweep.Sweep1D(mine, device=rank, dv_max=0.75, xtol_path=1.0)
        sw.upload()
        outs = sw.solve_resident()
        sw.finish(outs)
        rows, idx = sw.results_device(outs)
        table = sweep.gather_results(rows, idx, len(pts), world)
        torch.cuda.synchronize()
        ok_sweep = bool((table[:, 0] == 0).all()) and table.shape == (len(pts), sweep.N_SUMMARY)
        # every rank holds the same table; the rows of the OTHER rank's points are filled too
        other = torch.as_tensor([p.index for p in sweep.shard(pts, 1 - rank, world)], device=dev)
        ok_sweep = ok_sweep and bool((table[other, 1] > 0).all())
        mine_rows = table[idx]
        ok_sweep = ok_sweep and bool(torch.equal(mine_rows, rows))
        sw.close()
        # ---- 2. mesh-partitioned Newton solve of config 3 over NCCL vs the single-mesh path ---------------------
        mesh = meshio.load_mesh("L_50_R_5")
        prm = params.params_3d(L=50e-9, R=5e-9)
        part = partition.partition_z(mesh, world, ranks=[rank])
        pp = PartitionedPore(mesh, 50e-9, 5e-9, prm, part, TorchComm(part[0]), device=rank)
        nv = mesh.x.shape[0]
        ug = np.ones((nv, 9)); ug[:, 8] = 0.0
        us, uns = pp.from_global(ug * 0), pp.from_global(ug)
        out = pp.newton(us, uns, lin_rtol=1e-10, lin_restart=100)
        own = us[0][0, : part[0].n_own].cpu().numpy()
        ref = solver3d.PoreProblem(mesh, 50e-9, 5e-9, [prm], device=rank)
        ref.solver.set_dirichlet(ref.dirichlet_values([float(prm.extras["eq_scaled"][0])]))
        u = torch.zeros(1, nv, 9, dtype=torch.float64, device=dev)
        un = torch.as_tensor(ug[None], device=dev).contiguous()
        o_ref = ref.solver.newton(u, un, NewtonOpts.reference_3d())
        ur = u[0].cpu().numpy()[part[0].glob[: part[0].n_own]]
        err = max(np.linalg.norm(own[:, c] - ur[:, c]) / max(np.linalg.norm(ur[:, c]), 1e-300) for c in range(9))
        q.put((rank, ok_sweep, out["converged"], out["iters"], int(o_ref["iters"][0]), float(err), pp.comm.halo_bytes))
        pp.close()
        ref.solver.close()
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_sharded_sweep_gather_and_partitioned_solve_on_nccl(lib):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_sweep, conv, its, its_ref, err, halo in res:
        assert ok_sweep, rank
        assert conv and its == its_ref, (rank, its, its_ref)
        assert err <= 1e-8, (rank, err)
        assert halo > 0
