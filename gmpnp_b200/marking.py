"""Boundary marking and Dirichlet sets of the 3D pore problem (host side, once per mesh).

Replica of 3D/MPNP_CO2ER_pore.py:335-379 (SubDomain classes + MeshFunction marking) and
:460-467 (the six DirichletBC objects), with dolfin's semantics [upstream]:

* ``SubDomain.mark(mf, id)`` with check_midpoint=True marks a facet iff all three vertices AND
  the facet midpoint satisfy ``inside``; ``on_boundary`` is ignored by all three classes, so
  interior facets can be (and for narrow pores are) marked -- SURVEY finding 4 / App. F.
* marking order entry(1) -> exit(3) -> wall(2): the last marker wins on a facet.
* ``DirichletBC(V.sub(k), g, markers, id)`` constrains the DOFs of every vertex of every facet
  carrying ``id``; the BC list is applied in order, so a later BC overwrites an earlier one on
  shared DOFs (wall potential V wins over 0 on the rims).
"""
from __future__ import annotations

import numpy as np

from . import meshio


def wall_tolerance(L: float, R: float) -> float:
    """3D:352-355."""
    if (R == 5.0e-9 or R == 50.0e-9) and L == 10.0e-9:
        return 5.0e-3
    return 1.0e-3


def mark_facets(mesh: meshio.Mesh, aspect_pore: float, wall_tol: float):
    """Returns (facets[nf,3], shared_count[nf], marker[nf]) with marker in {1,2,3,9999}."""
    facets, cnt = meshio.tet_facets(mesh.cells)
    X = mesh.x[facets]                          # [nf, 3 verts, 3]
    mid = X.mean(axis=1)
    pts = np.concatenate([X, mid[:, None, :]], axis=1)     # 3 vertices + midpoint

    def all_inside(pred):
        return pred(pts).all(axis=1)

    tol = 1.0e-12
    entry = all_inside(lambda p: np.abs(p[..., 2] - 0.0) <= tol)
    exit_ = all_inside(lambda p: np.abs(p[..., 2] - 1.0) <= tol)
    wall = all_inside(lambda p: np.abs(p[..., 0] ** 2 + p[..., 1] ** 2 - aspect_pore ** 2) <= wall_tol)
    marker = np.full(len(facets), 9999, dtype=np.int64)
    marker[entry] = 1
    marker[exit_] = 3
    marker[wall] = 2
    return facets, cnt, marker


def dirichlet_sets(mesh: meshio.Mesh, L: float, R: float, ncomp: int = 9):
    """Dirichlet DOFs of 3D:460-467 in application order, de-duplicated so the last BC wins.

    Returns (dofs[int32], kind[int8]) with kind: 0 -> value 0 (potential at entry/exit),
    1 -> wall potential V, 2/3/4 -> CO2 / CO / H2 entry concentration; plus a dict of counts
    (unit-test targets of SURVEY App. F)."""
    aspect = R / L
    facets, cnt, marker = mark_facets(mesh, aspect, wall_tolerance(L, R))

    def verts(mid_):
        return np.unique(facets[marker == mid_])

    v1, v3, v2 = verts(1), verts(3), verts(2)
    ip = ncomp - 1
    order = [(v1, ip, 0), (v3, ip, 0), (v2, ip, 1), (v1, 4, 2), (v1, 5, 3), (v1, 6, 4)]
    val = {}
    for vs, comp, kind in order:
        for v in vs:
            val[int(v) * ncomp + comp] = kind           # later BC overwrites
    dofs = np.array(sorted(val.keys()), dtype=np.int32)
    kind = np.array([val[int(d)] for d in dofs], dtype=np.int8)
    # diagnostics
    ext = cnt == 1
    ext_verts = np.unique(facets[ext])
    interior_pinned = np.setdiff1d(v2, ext_verts)
    X = mesh.x[facets[marker == 2]]
    area2 = 0.5 * np.linalg.norm(np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), axis=1)
    ext2 = ext[marker == 2]
    info = dict(mk1=int((marker == 1).sum()), mk3=int((marker == 3).sum()), mk2=int((marker == 2).sum()),
                phi_V_verts=int(len(v2)), phi_0_verts=int(len(np.setdiff1d(np.union1d(v1, v3), v2))),
                interior_pinned=int(len(interior_pinned)), entry_gas_verts=int(len(v1)),
                wall_area_exterior=float(area2[ext2].sum()), wall_area_expected=float(2 * np.pi * aspect))
    return dofs, kind, info


def dirichlet_values(kind: np.ndarray, V: float, co2: float, co: float, h2: float) -> np.ndarray:
    """Values for the DOF list of :func:`dirichlet_sets`."""
    table = np.array([0.0, V, co2, co, h2])
    return table[kind.astype(np.int64)]


def facet_terms(mesh: meshio.Mesh, L: float, R: float):
    """Geometry of the INTENDED boundary integrals (3D:474-499; `ds(2)` wall, `ds(3)` pore exit; dolfin's ``ds``
    runs over EXTERIOR facets only, so interior facets that carry marker 2 do not contribute).

    Returns (wall_w[nv] = sum over the exterior wall facets of a vertex of area/3  -- the P1 integral of a
    constant flux --, exit_facets[nf, 3], exit_area[nf])."""
    facets, cnt, marker = mark_facets(mesh, R / L, wall_tolerance(L, R))
    ext = cnt == 1
    X = mesh.x[facets]
    area = 0.5 * np.linalg.norm(np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), axis=1)
    wall = ext & (marker == 2)
    wall_w = np.zeros(mesh.x.shape[0])
    np.add.at(wall_w, facets[wall].ravel(), np.repeat(area[wall] / 3.0, 3))
    ex = ext & (marker == 3)
    return wall_w, np.ascontiguousarray(facets[ex], dtype=np.int32), np.ascontiguousarray(area[ex])
