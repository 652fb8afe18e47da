"""Host logic of the mesh-partitioned 3D mode (BASELINE config 5): z-slab partition invariants, local Dirichlet
lists, and the halo exchange + all-reduce on a world_size-2 gloo group (CPU tensors, no kernels)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import cube_tet_mesh
from gmpnp_b200 import marking, meshio, partition


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partition_invariants(world):
    m = meshio.load_mesh("L_50_R_5")
    parts = partition.partition_z(m, world)
    nv, nt = m.x.shape[0], m.cells.shape[0]
    owned = np.concatenate([p.glob[: p.n_own] for p in parts])
    assert sorted(owned.tolist()) == list(range(nv))                       # every vertex has exactly one owner
    assert max(p.n_own for p in parts) - min(p.n_own for p in parts) <= 1  # balanced block rows
    owner = partition.vertex_owners(m.x, world)
    for p in parts:
        assert np.array_equal(p.x, m.x[p.glob])
        # local tets in global numbering = exactly the tets that touch an owned vertex
        touch = np.nonzero((owner[m.cells] == p.rank).any(axis=1))[0]
        assert np.array_equal(p.cell_glob, touch)
        assert np.array_equal(p.glob[p.cells], m.cells[touch])
        # every tet incident to an owned vertex is local => owned rows are complete
        assert (owner[p.glob[p.n_own:]] != p.rank).all()
        # send/recv lists pair up across ranks in the same (global id) order
        for nbr, ridx in p.recv.items():
            q = parts[nbr]
            assert np.array_equal(p.glob[ridx], q.glob[q.send[p.rank]])
            assert (ridx >= p.n_own).all() and (q.send[p.rank] < q.n_own).all()
        assert sum(len(v) for v in p.recv.values()) == p.n_ghost
        # interior rows (the first n_int) have no ghost column; every other owned row has one
        with_ghost = np.unique(p.cells[(p.cells >= p.n_own).any(axis=1)])
        bnd = with_ghost[with_ghost < p.n_own]
        assert 0 <= p.n_int <= p.n_own
        assert (bnd >= p.n_int).all() and len(bnd) == p.n_own - p.n_int
    # the slabs are ordered along z
    zc = [m.x[p.glob[: p.n_own], 2].mean() for p in parts]
    assert zc == sorted(zc)
    assert sum(len(p.cells) for p in parts) >= nt


def test_local_dirichlet_lists_cover_the_global_list():
    m = meshio.load_mesh("L_50_R_5")
    dofs, kind, _ = marking.dirichlet_sets(m, 50e-9, 5e-9)
    parts = partition.partition_z(m, 4)
    vals = marking.dirichlet_values(kind, -1.0, 2.87, 100.0, 100.0)
    seen = {}
    for p in parts:
        ld, sel = partition.local_dirichlet(p, dofs)
        assert np.all(np.diff(ld) > 0)
        g = p.glob[ld // 9] * 9 + ld % 9
        assert np.array_equal(g, dofs[sel])
        for d, s in zip(ld, sel):
            if d // 9 < p.n_own:
                seen[int(dofs[s])] = vals[s]
    assert sorted(seen) == sorted(dofs.tolist())                           # owned Dirichlet rows = the global set


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gmpnp_b200.dist3d import TorchComm
    m = cube_tet_mesh(4)
    part = partition.partition_z(m, world, ranks=[rank])[0]
    comm = TorchComm(part)
    rng = np.random.default_rng(3)
    xg = rng.normal(size=(m.x.shape[0], 9))
    xl = torch.as_tensor(partition.scatter_to_part(part, xg)).clone()
    xl[part.n_own:] = 0.0                                                  # ghosts unknown before the exchange
    comm.halo([xl])
    ok_halo = bool(np.array_equal(xl.numpy(), xg[part.glob]))
    local = torch.tensor([float((xg[part.glob[: part.n_own]] ** 2).sum()), float(part.n_own)], dtype=torch.float64)
    tot = comm.allreduce_sum([local])
    # ragged all-gather behind the distributed median (PartitionedPore.median): every rank ends up with all owned values
    vals = comm.gather_values([torch.as_tensor(xg[part.glob[: part.n_own], 3])])
    srt = torch.sort(vals)[0].numpy()
    ok_gather = bool(vals.numel() == m.x.shape[0] and np.array_equal(srt, np.sort(xg[:, 3])))
    q.put((rank, ok_halo and ok_gather, tot.tolist(), float((xg ** 2).sum()), m.x.shape[0], comm.halo_bytes,
           part.halo_doubles() * 8))
    dist.barrier()
    dist.destroy_process_group()


def test_halo_exchange_and_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_halo, tot, ref, nv, hb, hb_expected in res:
        assert ok_halo, rank
        assert abs(tot[0] - ref) <= 1e-12 * ref and tot[1] == nv
        assert hb == hb_expected > 0


def test_local_facets_give_every_owned_vertex_all_its_boundary_facets():
    """Intended boundary integrals in the partitioned mode: a part integrates the exit facets that touch one of its
    OWNED vertices; all their vertices are local, and per owned vertex the facet set equals the global one."""
    from gmpnp_b200 import marking, meshio, partition
    mesh = meshio.load_mesh("L_10_R_5")
    _, ef, ea = marking.facet_terms(mesh, 10e-9, 5e-9)
    nv = mesh.x.shape[0]
    glob_cnt = np.bincount(ef.ravel(), minlength=nv)
    for world in (2, 5):
        parts = partition.partition_z(mesh, world)
        covered = np.zeros(len(ef), dtype=int)
        for p in parts:
            lf, sel = partition.local_facets(p, ef, nv)
            assert lf.shape[1] == 3 and (len(lf) == 0 or (lf.min() >= 0 and lf.max() < p.n_local))
            assert np.array_equal(p.glob[lf], ef[sel])                      # same facets, local numbering
            covered[sel] += 1
            cnt = np.bincount(lf.ravel(), minlength=p.n_local)[: p.n_own]
            assert np.array_equal(cnt, glob_cnt[p.glob[: p.n_own]])         # owned rows see all their facets
        assert covered.min() >= 1                                            # the exit disc sits in the last slab(s)
