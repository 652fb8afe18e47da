"""Host logic: parameter layer KATs and mesh readers (App. E)."""
import os

import numpy as np
import pytest

from gmpnp_b200 import meshio, params
from conftest import REFERENCE


def test_params_1d_known_answers():
    p = params.params_1d()
    e = p.extras
    assert e["L_debye"] == 9.71397448177011e-10
    assert e["L_D"] == 1.9427948963540217e-05
    assert p.dt_scaled == 0.19003550024393479
    assert p.q == 1060892811.0140632
    assert p.jflux[4] == 0.0318624344537165
    assert p.jflux[1] == -13786.321496023305
    assert p.thermal_voltage == 0.025683333333333336
    assert p.kappa == 270855.9046587216
    assert abs(p.nu.sum() - 0.04854025368036322) < 1e-16
    assert 1 / p.nu[-1] == 57.23810941552707
    assert e["scale_R"][0] == 1914.0254664894212 and e["scale_R"][1] == 6650.8661477190435
    assert e["scale_vol"][2] == 0.03080801880571311


@pytest.mark.parametrize("conc,cat,Ld,JOH,sumnu", [
    (0.1, "Cs", 9.71397448177011e-10, -13786.321496023305, 0.04822547073284322),
    (0.5, "K", 4.344221454587252e-10, -2401.443665367024, 0.24139934294549267),
    (1.0, "K", 3.071828449514733e-10, -1011.8669508116575, 0.48154984571760906),
])
def test_params_1d_table(conc, cat, Ld, JOH, sumnu):
    p = params.params_1d(concentration_elec=conc, cation=cat)
    assert p.extras["L_debye"] == Ld
    assert p.jflux[1] == JOH
    assert abs(p.nu.sum() - sumnu) < 1e-15


def test_params_cs_extension():
    # documented extension: C0_Cs := C0_K at 0.5 / 1.0 M (SURVEY finding 5)
    p = params.params_1d(concentration_elec=0.5, cation="Cs")
    assert p.c0[-1] == 500.0
    with pytest.raises(KeyError):
        params.params_3d(cation="Cs")          # parameters_pore.yaml has no h_ion_Cs (3D:210)


def test_params_3d_known_answers():
    p = params.params_3d(L=50e-9, R=5e-9)
    assert p.time_constant == 1.3542795232936075e-05
    assert p.dt_scaled == 73.84000000000002
    assert p.q == 1060.892811014063
    assert abs(p.nu.sum() - 0.47950814749350074) < 1e-15
    assert np.allclose(p.extras["eq_conc"], (32.2031, 0.042621750000000035, 0.0038883000000000034), rtol=1e-15)
    assert np.allclose(p.extras["eq_scaled"], (2.872790735054118, 100.0, 100.0), rtol=1e-14)
    assert p.extras["Re"] == 28.005617977528093
    assert abs(params.sechenov_co2_scaled(p, 1, 1, 1, 1) * p.c0[4] - 22.69140767731971) < 1e-12
    J = p.extras["J_wall"]
    assert np.allclose([J[4], J[5], J[6], J[1]],
                       [0.001149679417614235, -28.44962180585293, -7.404196444343775, -4.806788491809415], rtol=1e-14)


def test_mesh_names():
    assert params.mesh_name_1d(50e-6) == "1D_variable_50um_mesh_5990"
    assert params.mesh_name_1d(200e-6) == "1D_variable_200um_mesh_4998"
    assert params.mesh_name_3d(50e-9, 2.5e-9) == "L_50_R_2"       # int() truncation, finding 6
    assert params.mesh_name_3d(100e-9, 5e-9) == "L_100_R_5"


@pytest.mark.parametrize("name,nv,nc", [
    ("1D_variable_1um_mesh_1090", 1091, 1090), ("1D_variable_50um_mesh_5990", 5991, 5990),
    ("1D_variable_200um_mesh_4998", 4999, 4998), ("L_50_R_5", 3679, 17297), ("L_10_R_5", 1767, 7696),
])
def test_packaged_meshes(name, nv, nc):
    m = meshio.load_mesh(name)
    assert m.num_vertices == nv and m.num_cells == nc
    if m.dim == 1:
        x = m.x[:, 0]
        assert x[0] == 0.0 and x[-1] == 1.0 and np.all(np.diff(x) > 0)
        assert np.array_equal(m.cells, np.stack([np.arange(nc), np.arange(1, nc + 1)], 1))
    else:
        assert abs(m.x[:, 2].min()) == 0.0 and m.x[:, 2].max() == 1.0
        e = meshio.tet_edges(m.cells)
        assert len(e) == {"L_50_R_5": 22431, "L_10_R_5": 10342}[name]
        f, cnt = meshio.tet_facets(m.cells)
        assert (cnt == 1).sum() == {"L_50_R_5": 2912, "L_10_R_5": 1760}[name]


def test_graded_interval_reproduces_reference_structure():
    ref = meshio.load_mesh("1D_variable_50um_mesh_5990").x[:, 0]
    gen = meshio.graded_interval(1000, 0.002, 4990).x[:, 0]
    assert np.abs(ref - gen).max() < 1e-14


def test_xml_roundtrip(tmp_path):
    m = meshio.load_mesh("L_10_R_5")
    sub = meshio.Mesh(x=m.x, cells=m.cells[:50], name="t")
    for fn in ("t.xml", "t.xml.gz"):
        p = str(tmp_path / fn)
        meshio.write_dolfin_xml(sub, p)
        r = meshio.read_dolfin_xml(p)
        assert np.array_equal(r.x, sub.x) and np.array_equal(r.cells, sub.cells)


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present")
def test_packaged_copies_match_reference_files():
    util = os.path.join(REFERENCE, "utilities")
    for name in ("1D_variable_10um_mesh_1990", "L_50_R_4"):
        a = meshio.load_mesh(name)
        b = meshio.load_mesh(name, utilities_dir=util)
        assert np.array_equal(a.x, b.x) and np.array_equal(a.cells, b.cells)
    assert params.params_1d(utilities_dir=util).pack().tolist() == params.params_1d().pack().tolist()
    assert params.params_3d(utilities_dir=util).pack().tolist() == params.params_3d().pack().tolist()


def test_red_refine_counts():
    m = meshio.load_mesh("L_10_R_5")
    r = meshio.red_refine(m, project_radius=0.5)
    assert r.num_cells == 8 * m.num_cells
    assert r.num_vertices == m.num_vertices + 10342
    # volumes positive-measure and total volume close to the cylinder's
    X = r.x[r.cells]
    vol = np.abs(np.linalg.det(X[:, 1:] - X[:, :1])) / 6
    assert vol.min() > 0
    assert abs(vol.sum() - np.pi * 0.25) / (np.pi * 0.25) < 2e-2
