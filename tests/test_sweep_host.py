"""Host logic of the sharded sweep: point enumeration, shards, ragged continuation paths, and the
final gather on a world_size-2 gloo group (CPU)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gmpnp_b200 import sweep


def test_config2_enumeration():
    pts = sweep.config2_points()
    assert len(pts) == 7680
    assert {p.cation for p in pts} == {"K", "Cs"} and {p.conc for p in pts} == {0.1, 0.5, 1.0}
    Vs = sorted({p.V for p in pts})
    assert len(Vs) == 256 and Vs[0] == -12.5 and abs(Vs[-1] + 12.5 / 256) < 1e-15
    assert [p.index for p in pts] == list(range(7680))


def test_shards_partition_the_sweep():
    pts = sweep.config2_points(8)
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            sh = sweep.shard(pts, r, world)
            seen += [p.index for p in sh]
            # every shard sees every mesh and (nearly) the same number of points
            assert {p.L_n for p in sh} == set(sweep.CONFIG2_LN)
            assert abs(len(sh) - len(pts) / world) <= 1
        assert sorted(seen) == list(range(len(pts)))


def test_voltage_paths_are_ragged_and_end_on_target():
    Vs = np.array([-0.05, -0.5, -0.51, -12.5, -3.0])
    path = sweep.voltage_paths(Vs, 0.5)
    assert path.shape == (5, 25)
    for b, V in enumerate(Vs):
        row = path[b][~np.isnan(path[b])]
        assert row[-1] == V
        assert np.all(np.abs(np.diff(np.concatenate([[0.0], row]))) <= 0.5 + 1e-12)
        assert np.all(np.isnan(path[b][len(row):]))
    assert np.isnan(path[0, 1:]).all() and not np.isnan(path[3]).any()


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pts = sweep.config2_points(4)[:n_total]
    mine = sweep.shard(pts, rank, world)
    idx = torch.tensor([p.index for p in mine], dtype=torch.int64)
    local = torch.stack([idx.double() * 2.0, torch.tensor([p.V for p in mine], dtype=torch.float64)], dim=1)
    table = sweep.gather_results(local, idx, n_total, world)
    if rank == 0:
        q.put(table.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [120, 77])
def test_gather_results_gloo_world2(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_total
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    table = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pts = sweep.config2_points(4)[:n_total]
    assert np.array_equal(table[:, 0], 2.0 * np.arange(n_total))
    assert np.array_equal(table[:, 1], np.array([p.V for p in pts]))


def test_gather_results_single_rank():
    local = torch.arange(12, dtype=torch.float64).reshape(6, 2)
    idx = torch.tensor([5, 0, 3, 1, 2, 4])
    out = sweep.gather_results(local, idx, 6, 1)
    assert torch.equal(out[idx], local)


def test_config4_enumeration_and_sharding():
    """Config 4: every present pore mesh x the 256-voltage grid; shards are disjoint, cover the sweep and
    keep a rank's points of one mesh together as one batch."""
    from gmpnp_b200 import meshio, sweep3d
    pts = sweep3d.config4_points()
    assert len(pts) == 11 * 256 and [p.index for p in pts] == list(range(len(pts)))
    assert pts[0].V == -12.5 / 256 and pts[255].V == -12.5 and pts[256].mesh != pts[255].mesh
    for name, L, R in sweep3d.CONFIG4_MESHES:
        m = meshio.load_mesh(name)                        # all eleven files are packaged
        assert m.dim == 3 and m.cells.shape[1] == 4
    world = 8
    seen = []
    for r in range(world):
        mine = sweep3d.shard(pts, r, world)
        seen += [p.index for p in mine]
        sw = sweep3d.Sweep3D(mine)
        assert len(sw.by_mesh) == 11 and sum(len(v) for v in sw.by_mesh.values()) == len(mine)
        assert all(len(v) == 32 for v in sw.by_mesh.values())         # 256 voltages / 8 ranks per mesh
    assert sorted(seen) == list(range(len(pts)))
