"""Oracle-side checks of the 3D helpers that do not need a GPU."""
import numpy as np

from conftest import cube_tet_mesh
from oracle import solver as osolver


def test_p1_gradient_projection_is_exact_for_linear_fields_and_conserves_the_mean():
    m = cube_tet_mesh(3)
    lin = (2.0 * m.x[:, 0] - 3.0 * m.x[:, 1] + 0.5 * m.x[:, 2] + 1.0)[:, None]
    g = osolver.p1_gradient_projection_3d(m.x, m.cells, lin)
    assert np.abs(g[:, 0, :] - np.array([2.0, -3.0, 0.5])).max() < 1e-12
    # L2 projection preserves integrals: int g_d = int (grad f)_d = sum_t vol_t grad f_t
    rng = np.random.default_rng(2)
    f = rng.normal(size=(m.x.shape[0], 2))
    g = osolver.p1_gradient_projection_3d(m.x, m.cells, f)
    from oracle import forms
    gl, vol = forms.geometry(m.x, m.cells)
    gradf = np.einsum("cad,cak->ckd", gl, f[m.cells])
    lhs = np.einsum("c,cakd->kd", vol / 4.0, g[m.cells])
    rhs = np.einsum("c,ckd->kd", vol, gradf)
    assert np.abs(lhs - rhs).max() < 1e-12 * max(1.0, np.abs(rhs).max())


def _rd3_setup(nsub=3):
    """A cube mesh with made-up wall weights / exit facets (the algebra does not care where the facets are)."""
    from gmpnp_b200 import meshio, rxn_diff3d
    from oracle import rxn_diff3d as ord3
    m = cube_tet_mesh(nsub)
    p = rxn_diff3d.params_rxn_diff_3d(L=50e-9, R=5e-9)
    facets, cnt = meshio.tet_facets(m.cells)
    ext = facets[cnt == 1]
    top = ext[(np.abs(m.x[ext][:, :, 2] - m.x[:, 2].max()) < 1e-12).all(axis=1)]
    X = m.x[top]
    area = 0.5 * np.linalg.norm(np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), axis=1)
    rng = np.random.default_rng(5)
    wall_w = rng.random(m.x.shape[0]) * 0.01
    return m, p, wall_w, top, area, ord3


def test_rxn_diff_3d_independent_oracle_matches_the_embedded_gmpnp_oracle():
    """3D/rxn_diff_CO2ER_pore.py restated twice: the 7-species oracle with exact monomial integrals
    (oracle/rxn_diff3d.py) and the 9-component GMPNP oracle with z = nu = 0 (what the CUDA path runs).  Residual rows
    and Jacobian action of the seven species agree to round-off; the passenger rows of the embedding vanish."""
    m, p, wall_w, top, area, ord3 = _rd3_setup()
    nv = m.x.shape[0]
    rng = np.random.default_rng(11)
    u7 = np.exp(0.3 * rng.normal(size=(nv, 7)))
    un7 = np.ones((nv, 7))
    d7 = ord3.RxnDiff3D(m.x, m.cells, p, wall_w, top, area)
    F7 = d7.residual(u7.ravel(), un7.ravel()).reshape(nv, 7)
    ft = (wall_w, top, area, p.extras["J_wall"], p.extras["k_exit"])
    d9 = osolver.Discretisation(m.x, m.cells, 9, facet_terms=ft)
    u9 = np.concatenate([u7, np.ones((nv, 1)), np.zeros((nv, 1))], axis=1)
    un9 = np.concatenate([un7, np.ones((nv, 1)), np.zeros((nv, 1))], axis=1)
    F9 = d9.residual(u9.ravel(), un9.ravel(), p).reshape(nv, 9)
    scale = np.abs(F7).max(axis=0)
    assert (np.abs(F9[:, :7] - F7).max(axis=0) <= 1e-12 * scale).all()
    assert np.abs(F9[:, 7:]).max() <= 1e-15           # K @ 1 is zero only to round-off
    x7 = rng.normal(size=(nv, 7))
    x9 = np.concatenate([x7, np.zeros((nv, 2))], axis=1)
    Jx7 = (d7.jacobian(u7.ravel()) @ x7.ravel()).reshape(nv, 7)
    Jx9 = (d9.jacobian(u9.ravel(), p) @ x9.ravel()).reshape(nv, 9)
    assert (np.abs(Jx9[:, :7] - Jx7).max(axis=0) <= 1e-11 * np.abs(Jx7).max(axis=0)).all()
    # the species rows do not see the passengers either: columns of cation and potential are zero in rows 0..6
    xp = np.zeros((nv, 9)); xp[:, 7:] = rng.normal(size=(nv, 2))
    assert np.abs((d9.jacobian(u9.ravel(), p) @ xp.ravel()).reshape(nv, 9)[:, :7]).max() == 0.0


def test_rxn_diff_3d_oracle_jacobian_is_the_derivative_of_its_residual():
    m, p, wall_w, top, area, ord3 = _rd3_setup(2)
    nv = m.x.shape[0]
    rng = np.random.default_rng(12)
    u = np.exp(0.3 * rng.normal(size=nv * 7))
    un = np.ones(nv * 7)
    d = ord3.RxnDiff3D(m.x, m.cells, p, wall_w, top, area)
    x = rng.normal(size=nv * 7)
    # the residual is quadratic in u: the central difference is exact up to round-off
    h = 1e-3
    fd = (d.residual(u + h * x, un) - d.residual(u - h * x, un)) / (2 * h)
    Jx = d.jacobian(u) @ x
    assert np.abs(fd - Jx).max() <= 1e-9 * np.abs(Jx).max()


def test_rxn_diff_3d_golden_vectors_present_and_consistent():
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "rxn_diff_3d_L10R5.npz"))
    assert g["steps"].shape == (2, 1767, 7) and g["its"].tolist() == [7, 5]
    assert np.isfinite(g["steps"]).all() and g["steps"][:, :, 4].min() > 0.0


def test_two_independent_restatements_agree_on_the_polynomial_terms():
    """oracle/pnp3d_exact.py (exact monomial integrals, written from the reference's forms, no quadrature tables) against
    oracle/forms.py + oracle/quadrature.py with nu = 0: residual and Jacobian of the full 9-component problem agree to
    round-off on a generic tet mesh -- every term except the rational steric one is pinned by two restatements."""
    import numpy as np
    from conftest import admissible_state, cube_tet_mesh
    from gmpnp_b200 import params
    from oracle import pnp3d_exact, solver as osolver
    mesh = cube_tet_mesh(3, scale=(0.2, 0.3, 1.0))
    n = mesh.num_vertices
    for kw in (dict(L=50e-9, R=5e-9), dict(L=100e-9, R=5e-9, concentration_elec=0.5, time_step=1e-5)):
        prm = params.params_3d(**kw)
        p0 = prm.with_(nu=np.zeros(8))
        rng = np.random.default_rng(3)
        u = admissible_state(rng, n, 8, prm.nu, V=-5.0).ravel()
        un = admissible_state(rng, n, 8, prm.nu, V=-5.0).ravel()
        disc = osolver.Discretisation(mesh.x, mesh.cells, 9)
        ex = pnp3d_exact.Pnp3DExact(mesh.x, mesh.cells, p0)
        F1, F2 = disc.residual(u, un, p0), ex.residual(u, un)
        assert np.abs(F1 - F2).max() <= 1e-12 * np.abs(F2).max()
        A1, A2 = disc.jacobian(u, p0).toarray(), ex.jacobian(u).toarray()
        rows = np.abs(A2).max(axis=1, keepdims=True)
        assert (np.abs(A1 - A2) <= 1e-12 * rows).all()
        # the exact Jacobian is the derivative of the exact residual (central differences, relative step 1e-6)
        v = rng.normal(size=u.shape)
        h = 1e-6
        fd = (ex.residual(u + h * v, un) - ex.residual(u - h * v, un)) / (2 * h)
        assert np.abs(fd - A2 @ v).max() <= 1e-6 * np.abs(A2 @ v).max()
