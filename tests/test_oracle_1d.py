"""The oracle itself: Jacobian consistency, exact identities, golden vectors, Stern brackets."""
import os

import numpy as np
import pytest

from gmpnp_b200 import meshio, params
from oracle import forms, quadrature, solver
from conftest import GOLDEN, admissible_state

# the only numbers in the reference that pin results: 1D/Stern_CO2ER.py:66-68
STERN = {-2.5: (-0.08032108300135771, 74.56149297894756), -5.0: (-0.2524415478848975, 57.64572780716129),
         -7.5: (-0.4612956299192668, 50.16243860179017), -10.0: (-0.6149631587776277, 49.311548142969336),
         -12.5: (-0.7310301485096051, 49.2556833480052)}


def small_problem(ncomp=7, seed=0):
    m = meshio.graded_interval(5, 0.01, 3)
    prm = params.params_1d()
    rng = np.random.default_rng(seed)
    u = admissible_state(rng, m.num_vertices, 6, prm.nu)
    un = admissible_state(rng, m.num_vertices, 6, prm.nu)
    return m, prm, u, un


def test_quadrature_rules():
    for rule, deg in ((quadrature.interval_rule(2), 3), (quadrature.interval_rule(3), 5)):
        lam, w = rule
        for k in range(deg + 1):
            assert abs((w * lam[:, 1] ** k).sum() - 1.0 / (k + 1)) < 1e-15
    lam, w = quadrature.tet_rule_degree3()
    assert len(w) == 5 and abs(w.sum() - 1) < 1e-15 and w.min() < 0
    lam, w = quadrature.tet_rule_degree4()
    assert len(w) == 14 and abs(w.sum() - 1) < 1e-14


def test_jacobian_matches_complex_step_of_residual_same_rule():
    m, prm, u, un = small_problem()
    d = solver.Discretisation(m.x, m.cells, 7)
    d.ruleF = d.ruleJ
    J = d.jacobian(u.ravel(), prm).toarray()
    Jc = np.zeros_like(J)
    for k in range(d.ndof):
        up = u.ravel().astype(complex)
        up[k] += 1e-30j
        Jc[:, k] = d.residual(up, un.ravel().astype(complex), prm).imag / 1e-30
    assert np.abs(J - Jc).max() <= 1e-12 * np.abs(J).max()


def test_exact_identities():
    m, prm, u, un = small_problem()
    d = solver.Discretisation(m.x, m.cells, 7)
    ones = np.tile([1.0] * 6 + [0.0], m.num_vertices)
    # water and carbonate terms of R vanish at u = 1 up to the non-equilibrium CO2 hydration imbalance
    mr = forms.minus_R(np.ones((1, 7)), prm)[0]
    assert abs(mr[0]) / (prm.s[0] * prm.rate["kw1"]) < 1e-10          # water: kw2 cH cOH = kw1
    # residual is affine in the point fluxes
    F0 = d.residual(u.ravel(), un.ravel(), prm, 0 * prm.jflux)
    F1 = d.residual(u.ravel(), un.ravel(), prm, prm.jflux)
    F2 = d.residual(u.ravel(), un.ravel(), prm, 2 * prm.jflux)
    assert np.allclose(F2 - F1, F1 - F0, rtol=0, atol=1e-9 * np.abs(F1).max())
    # PNP = MPNP with nu = 0
    p0 = params.params_1d(model="PNP")
    assert np.all(p0.nu == 0)
    Fp = d.residual(u.ravel(), un.ravel(), p0)
    Fm = d.residual(u.ravel(), un.ravel(), prm.with_(nu=np.zeros(6)))
    assert np.array_equal(Fp, Fm)
    # rows of the pure-diffusion element matrix sum to zero
    pd = prm.with_(nu=np.zeros(6), z=np.zeros(6), kappa=0.0, s=np.zeros(5), q=0.0)
    Je = forms.element_jacobian(d.gather(ones), d.g, d.vol, pd, d.ruleJ)
    assert np.abs(Je[:, :, :6, :, :6].sum(axis=3)).max() < 1e-9 * np.abs(Je).max()


def test_golden_assemble_is_reproduced():
    g = np.load(os.path.join(GOLDEN, "assemble_1d.npz"))
    x, u, un = g["x"], g["u"], g["un"]
    prm = params.params_1d()
    n = len(x)
    cells = np.stack([np.arange(n - 1), np.arange(1, n)], 1)
    d = solver.Discretisation(x, cells, 7)
    bd, bv = solver.bc_1d(n, 7, prm.V)
    F = solver.apply_bc_residual(d.residual(u.ravel(), un.ravel(), prm, prm.jflux), u.ravel(), bd, bv)
    assert np.allclose(F.reshape(n, 7), g["F"], rtol=1e-13, atol=0)


def test_golden_march_1um_and_newton_counts():
    g = np.load(os.path.join(GOLDEN, "march_1um.npz"))
    m = meshio.load_mesh("1D_variable_1um_mesh_1090")
    prm = params.params_1d(L_n=1.0e-6)
    hist, its, _ = solver.march_1d(m.x[:, 0], prm, 2)
    assert its == g["its"][:2].tolist()
    assert np.allclose(hist, g["hist"][:3], rtol=1e-10, atol=1e-12)
    assert np.all(hist[0][:, :6] == 1) and np.all(hist[0][:, 6] == 0)     # row 0 = initial state (1D:623-629)


def test_golden_config1_first_steps():
    g = np.load(os.path.join(GOLDEN, "march_50um.npz"))
    assert g["its"].tolist() == [7, 3, 3]      # SURVEY 3.1 probe: 7, 3x9, 2x90


def test_stern_soft_known_answers():
    """Golden (field_OHP, eps_rel_OHP) of 1D/Stern_CO2ER.py:66-68 lie between the oracle's dry-run and
    steady values; the steady values must agree within 1.5 % / 0.3 %."""
    g = np.load(os.path.join(GOLDEN, "steady_50um.npz"))
    for V, (E, eps) in STERN.items():
        f, e = g[f"ohp_{V}"]
        assert abs(f - E) / abs(E) < 1.5e-2, (V, f, E)
        assert abs(e - eps) / eps < 3e-3, (V, e, eps)


@pytest.mark.slow
def test_steady_oracle_reproduces_golden_at_minus_one():
    g = np.load(os.path.join(GOLDEN, "steady_50um.npz"))
    m = meshio.load_mesh("1D_variable_50um_mesh_5990")
    prm = params.params_1d()
    u, its = solver.steady_1d(m.x[:, 0], prm, [-0.5, -1.0])
    ref = g["u_-1.0"]
    for c in range(7):
        assert np.linalg.norm(u[:, c] - ref[:, c]) <= 1e-9 * np.linalg.norm(ref[:, c])


def test_h_ohp_ladder_branches():
    f = solver.h_ohp_ladder
    assert f(0.001, -0.1, 1.1) == 0.001 / 1.1
    assert f(0.001, 1.0, 1.1) == 0.001 / 1.05
    assert f(0.001, 1.06, 1.1) == 0.001 / 1.01
    assert f(0.001, 1.2, 1.1) == 0.001 * 1.04
    assert f(0.001, 1.6, 1.1) == 0.001 * 1.15
    assert f(0.001, 1.09, 1.1) == 0.001


@pytest.mark.slow
@pytest.mark.parametrize("V", [-2.5, -10.0])
def test_oracle_reproduces_the_reference_stern_table(V):
    """PINNED AGAINST THE REFERENCE'S OWN NUMBERS.  The (field_OHP, eps_rel_OHP) pairs pasted into
    1D/Stern_CO2ER.py:66-68 are the output of the default non-dry run of 1D/MPNP_CO2ER_EDL.py, which integrates
    20 000 steps of 1e-5 s (the `del_t` rebinding at 1D:643-646 never reaches the form), i.e. the state at t = 0.2 s.
    The oracle's backward-Euler march from u = 0 to t = 0.2 s reproduces them: to 3e-13 by the literal 20 000-step
    replay at V = -2.5, to 1e-7 with steps of 1e-4 s and to 3e-10 extrapolated to the reference's own step at all five voltages (tests/golden/stern_pin_results.json, generated by
    tests/studies/stern_pin_study.py over ~2 h of CPU); here a 67-step march, good to 5e-5 (first-order in dt)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("stern_pin_study",
                                                  os.path.join(os.path.dirname(__file__), "studies", "stern_pin_study.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    f, e, steps, its = mod.march_to(V)
    E, eps = STERN[V]
    assert abs(f / E - 1) < 5e-5, (f, E)
    assert abs(e / eps - 1) < 1e-5, (e, eps)
    assert steps < 100
