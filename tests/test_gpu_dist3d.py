"""Mesh-partitioned 3D mode on ONE GPU: all parts live in one process (LocalComm), so the distributed SpMV,
dot products, GMRES and Newton step are compared with the single-mesh path on the same data."""
import numpy as np
import pytest

from conftest import cube_tet_mesh

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _setup(world, mesh=None, kappa_scale=1.0, coarse=True, intended_bcs=False):
    from gmpnp_b200 import meshio, params, partition, solver3d
    from gmpnp_b200.dist3d import LocalComm, PartitionedPore
    mesh = mesh or meshio.load_mesh("L_50_R_5")
    prm = params.params_3d(L=50e-9, R=5e-9)
    if kappa_scale != 1.0:
        # a short pseudo-time step: with block-Jacobi alone (no coarse space in partitioned mode) the pure-Neumann
        # species blocks of the reference step (dt_scaled = 73.84) converge too slowly for a unit test
        prm = prm.with_(kappa=prm.kappa * kappa_scale)
    parts = partition.partition_z(mesh, world)
    pp = PartitionedPore(mesh, 50e-9, 5e-9, prm, parts, LocalComm(parts), coarse=coarse, intended_bcs=intended_bcs)
    ref = solver3d.PoreProblem(mesh, 50e-9, 5e-9, [prm], intended_bcs=intended_bcs)
    ref.solver.set_dirichlet(ref.dirichlet_values([float(prm.extras["eq_scaled"][0])]))
    return mesh, prm, parts, pp, ref


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_assembly_and_spmv_match_the_single_mesh_path(lib, world):
    from gmpnp_b200 import partition
    mesh, prm, parts, pp, ref = _setup(world)
    rng = np.random.default_rng(5)
    nv = mesh.x.shape[0]
    ug = np.ones((nv, 9)); ug[:, 8] = 0.0
    ug += 0.01 * rng.random((nv, 9))
    ung = np.ones((nv, 9)); ung[:, 8] = 0.0
    dev = pp.device
    Fg, Jg = ref.solver.assemble(torch.as_tensor(ug[None], device=dev), torch.as_tensor(ung[None], device=dev))
    us, uns = pp.from_global(ug), pp.from_global(ung)
    for u, p in zip(us, parts):
        u[:, p.n_own:] = 0.0                                     # ghosts must come from the halo exchange
    Fs, nrm = pp.assemble(us, uns)
    Fd = partition.gather_owned(parts, [F[0].cpu().numpy() for F in Fs], nv)
    Fr = Fg[0].cpu().numpy()
    assert np.abs(Fd - Fr).max() <= 1e-12 * np.abs(Fr).max()
    assert abs(nrm - np.linalg.norm(Fr)) <= 1e-12 * np.linalg.norm(Fr)
    xg = rng.normal(size=(nv, 9))
    yr = ref.solver.spmv(Jg, torch.as_tensor(xg[None], device=dev))[0].cpu().numpy()
    xs = pp.from_global(xg)
    for x, p in zip(xs, parts):
        x[:, p.n_own:] = 0.0
    for overlap in (False, True):                                # interior rows on a side stream during the exchange
        for x, p in zip(xs, parts):
            x[:, p.n_own:] = 0.0
        ys = pp.spmv(xs, overlap=overlap)
        torch.cuda.synchronize()
        yd = partition.gather_owned(parts, [y[0].cpu().numpy() for y in ys], nv)
        assert np.abs(yd - yr).max() <= 1e-12 * np.abs(yr).max(), overlap
    assert pp.comm.halo_bytes > 0
    # reduced dot products over owned rows = global dot products
    ws = [x.view(-1) for x in xs]
    d = pp.dot_owned([x.view(1, -1) for x in xs], 1, ws, extra_self=True).cpu().numpy()
    assert abs(d[0] - (xg ** 2).sum()) <= 1e-12 * (xg ** 2).sum() and abs(d[1] - d[0]) <= 1e-12 * d[0]


def test_partitioned_gmres_solves_the_global_system(lib):
    """Distributed GMRES + block-Jacobi on a small box mesh (well conditioned with the time term): the true
    residual of the gathered solution against the single-mesh operator is at the requested tolerance."""
    from gmpnp_b200 import partition
    mesh = cube_tet_mesh(4, scale=(0.1, 0.1, 1.0))
    mesh_, prm, parts, pp, ref = _setup(3, mesh, kappa_scale=1.0e4)
    nv = mesh.x.shape[0]
    ug = np.ones((nv, 9)); ug[:, 8] = 0.0
    dev = pp.device
    ut = torch.as_tensor(ug[None], device=dev)
    Fg, Jg = ref.solver.assemble(ut * 0, ut)                     # first Newton step of the reference march: u = 0
    us = pp.from_global(ug * 0); uns = pp.from_global(ug)
    Fs, _ = pp.assemble(us, uns)
    its_log = []
    xs, its, rel = pp.gmres(Fs, m=60, maxit=600, rtol=1e-10, callback=lambda k, r: its_log.append(r))
    assert rel <= 1e-10 and 0 < its <= 600
    xd = partition.gather_owned(parts, [x[0].cpu().numpy() for x in xs], nv)
    res = Fg[0].cpu().numpy() - ref.solver.spmv(Jg, torch.as_tensor(xd[None], device=dev))[0].cpu().numpy()
    assert np.linalg.norm(res) <= 1e-8 * np.linalg.norm(Fg[0].cpu().numpy())
    assert pp.stats["allreduce"] >= 3 * its                      # CGS2 (fused norm) + the coarse restriction


def test_partitioned_newton_step_matches_single_mesh_newton(lib):
    """One reference solve() (3D:789-799, relaxation 0.9, residual criterion) on the partitioned box mesh: same
    Newton count and the same iterate as the single-mesh CUDA path."""
    from gmpnp_b200 import partition
    from gmpnp_b200._lib import NewtonOpts
    mesh = cube_tet_mesh(4, scale=(0.1, 0.1, 1.0))
    mesh_, prm, parts, pp, ref = _setup(2, mesh, kappa_scale=1.0e4)
    nv = mesh.x.shape[0]
    ug = np.ones((nv, 9)); ug[:, 8] = 0.0
    dev = pp.device
    u = torch.zeros(1, nv, 9, dtype=torch.float64, device=dev)
    un = torch.as_tensor(ug[None], device=dev).contiguous()
    o = NewtonOpts.reference_3d()
    out_ref = ref.solver.newton(u, un, o)
    us = pp.from_global(ug * 0); uns = pp.from_global(ug)
    out = pp.newton(us, uns, lin_rtol=1e-10, lin_restart=100)
    assert out["converged"] and int(out_ref["status"][0]) == 0
    assert out["iters"] == int(out_ref["iters"][0])
    assert abs(out["r0"] - float(out_ref["r0"][0])) <= 1e-10 * out["r0"]
    ud = partition.gather_owned(parts, [x[0].cpu().numpy() for x in us], nv)
    ur = u[0].cpu().numpy()
    for c in range(9):
        assert np.linalg.norm(ud[:, c] - ur[:, c]) <= 1e-8 * max(np.linalg.norm(ur[:, c]), 1e-300), c


def test_partitioned_newton_on_config3_with_distributed_coarse_space(lib):
    """Config 3 itself (L_50_R_5, reference time step): with the z-slab coarse space summed over the parts the
    partitioned GMRES converges like the single-mesh one (block-Jacobi alone needs thousands of iterations here),
    and the damped Newton solve reproduces the single-mesh iterate."""
    from gmpnp_b200 import partition
    from gmpnp_b200._lib import NewtonOpts
    mesh, prm, parts, pp, ref = _setup(3)
    nv = mesh.x.shape[0]
    ug = np.ones((nv, 9)); ug[:, 8] = 0.0
    dev = pp.device
    u = torch.zeros(1, nv, 9, dtype=torch.float64, device=dev)
    un = torch.as_tensor(ug[None], device=dev).contiguous()
    out_ref = ref.solver.newton(u, un, NewtonOpts.reference_3d())
    us = pp.from_global(ug * 0); uns = pp.from_global(ug)
    out = pp.newton(us, uns, lin_rtol=1e-10, lin_restart=100)
    assert out["converged"] and out["iters"] == int(out_ref["iters"][0])
    assert out["lin_iters"] <= 1.5 * int(out_ref["lin_iters"][0]) + 50, (out["lin_iters"], int(out_ref["lin_iters"][0]))
    ud = partition.gather_owned(parts, [x[0].cpu().numpy() for x in us], nv)
    ur = u[0].cpu().numpy()
    for c in range(9):
        assert np.linalg.norm(ud[:, c] - ur[:, c]) <= 1e-8 * max(np.linalg.norm(ur[:, c]), 1e-300), c


def test_partitioned_mode_with_intended_boundary_integrals(lib):
    """`--intended_bcs` (3D:474-499) in the mesh-partitioned mode: exit facets are given to every part that owns one
    of their vertices, wall weights are restricted per vertex; residual, Jacobian action and the damped Newton solve
    equal the single-mesh path with the same terms -- and differ from the as-executed form."""
    from gmpnp_b200 import partition
    from gmpnp_b200._lib import NewtonOpts
    mesh, prm, parts, pp, ref = _setup(3, intended_bcs=True)
    rng = np.random.default_rng(8)
    nv = mesh.x.shape[0]
    ug = np.ones((nv, 9)); ug[:, 8] = 0.0
    ug += 0.01 * rng.random((nv, 9))
    ung = np.ones((nv, 9)); ung[:, 8] = 0.0
    dev = pp.device
    Fg, Jg = ref.solver.assemble(torch.as_tensor(ug[None], device=dev), torch.as_tensor(ung[None], device=dev))
    us, uns = pp.from_global(ug), pp.from_global(ung)
    Fs, nrm = pp.assemble(us, uns)
    Fd = partition.gather_owned(parts, [F[0].cpu().numpy() for F in Fs], nv)
    Fr = Fg[0].cpu().numpy()
    assert np.abs(Fd - Fr).max() <= 1e-12 * np.abs(Fr).max()
    xg = rng.normal(size=(nv, 9))
    yr = ref.solver.spmv(Jg, torch.as_tensor(xg[None], device=dev))[0].cpu().numpy()
    ys = pp.spmv(pp.from_global(xg))
    yd = partition.gather_owned(parts, [y[0].cpu().numpy() for y in ys], nv)
    assert np.abs(yd - yr).max() <= 1e-12 * np.abs(yr).max()
    ref.solver.set_facet_terms()                                  # as executed: the terms really were in
    F0, _ = ref.solver.assemble(torch.as_tensor(ug[None], device=dev), torch.as_tensor(ung[None], device=dev), want_J=False)
    assert np.abs(F0[0].cpu().numpy() - Fr).max() > 1e-3
    # one reference solve() from u = 0
    mesh, prm, parts, pp, ref = _setup(2, intended_bcs=True)
    ug = np.ones((nv, 9)); ug[:, 8] = 0.0
    u = torch.zeros(1, nv, 9, dtype=torch.float64, device=dev)
    un = torch.as_tensor(ug[None], device=dev).contiguous()
    out_ref = ref.solver.newton(u, un, NewtonOpts.reference_3d())
    us = pp.from_global(ug * 0); uns = pp.from_global(ug)
    out = pp.newton(us, uns, lin_rtol=1e-10, lin_restart=100)
    assert out["converged"] and out["iters"] == int(out_ref["iters"][0])
    ud = partition.gather_owned(parts, [x[0].cpu().numpy() for x in us], nv)
    ur = u[0].cpu().numpy()
    for c in range(9):
        assert np.linalg.norm(ud[:, c] - ur[:, c]) <= 1e-8 * max(np.linalg.norm(ur[:, c]), 1e-300), c


def test_partitioned_march_with_distributed_median_matches_single_mesh_march(lib):
    """The reference's loop 3D:782-858 on a partitioned mesh (PartitionedPore.march: Dirichlet rebuild, damped Newton,
    Sechenov update from medians gathered over the parts, u_n <- u) against the library march of the single-mesh path:
    same Newton counts, same CO2 entry values, same states, on L_10_R_5 cut into 3 z-slabs."""
    from gmpnp_b200 import meshio, params, partition, solver3d
    from gmpnp_b200.dist3d import LocalComm, PartitionedPore
    mesh = meshio.load_mesh("L_10_R_5")
    prm = params.params_3d(L=10e-9, R=5e-9, voltage_multiplier=-0.75)
    parts = partition.partition_z(mesh, 3)
    pp = PartitionedPore(mesh, 10e-9, 5e-9, prm, parts, LocalComm(parts))
    ref = solver3d.PoreProblem(mesh, 10e-9, 5e-9, [prm])
    steps = 3
    out_ref = ref.march(steps, history=True)
    out = pp.march(steps, lin_rtol=1e-10, lin_restart=100)
    assert out["iters"] == out_ref["iters"][:, 0].tolist()
    co2_ref = out_ref["co2_entry"][:, 0]
    assert np.abs(np.array(out["co2_entry"]) - co2_ref).max() <= 1e-9 * np.abs(co2_ref).max()
    nv = mesh.x.shape[0]
    ud = partition.gather_owned(parts, [x[0].cpu().numpy() for x in out["us"]], nv)
    ur = out_ref["history"][steps][0]
    for c in range(9):
        assert np.linalg.norm(ud[:, c] - ur[:, c]) <= 1e-8 * max(np.linalg.norm(ur[:, c]), 1e-300), c
    # the medians themselves: the gathered order statistic equals np.median of the gathered field
    for c in (1, 7):
        assert pp.median(out["us"], c) == float(np.median(ud[:, c]))
    pp.close(); ref.solver.close()
