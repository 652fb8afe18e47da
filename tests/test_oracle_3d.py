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
