"""Mesh input for the GMPNP hot path (host side, runs once per experiment).

Replaces dolfin's ``Mesh(<file>.xml[.gz])`` reader used at
1D/MPNP_CO2ER_EDL.py:231-234 and 3D/MPNP_CO2ER_pore.py:329-332.  The dolfin XML
layout is ``<mesh celltype dim><vertices><vertex index x [y z]/>...<cells>
<interval|tetrahedron index v0 v1 [v2 v3]/>``.  Besides the XML reader there is
a compact ``.npz`` form (``x`` float64 [nv, dim], ``cells`` int32 [nc, dim+1])
used for the packaged copies of the reference meshes, which have to travel to
machines where the reference checkout does not exist.
"""
from __future__ import annotations

import gzip
import os
import re
from dataclasses import dataclass

import numpy as np

_PKG_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

_VERT_RE = re.compile(
    rb'<vertex\s+index="(\d+)"\s+x="([^"]+)"(?:\s+y="([^"]+)")?(?:\s+z="([^"]+)")?')
_CELL_RE = re.compile(
    rb'<(?:interval|triangle|tetrahedron)\s+index="(\d+)"\s+v0="(\d+)"\s+v1="(\d+)"'
    rb'(?:\s+v2="(\d+)")?(?:\s+v3="(\d+)")?')


@dataclass
class Mesh:
    """Vertex coordinates and cell connectivity, in file order."""
    x: np.ndarray      # [nv, dim] float64
    cells: np.ndarray  # [nc, dim+1] int32
    name: str = ""

    @property
    def dim(self) -> int:
        return self.x.shape[1]

    @property
    def num_vertices(self) -> int:
        return self.x.shape[0]

    @property
    def num_cells(self) -> int:
        return self.cells.shape[0]


def read_dolfin_xml(path: str) -> Mesh:
    """Parse a dolfin XML mesh (optionally gzipped)."""
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as f:
        raw = f.read()
    m = re.search(rb'<mesh\s+celltype="(\w+)"\s+dim="(\d+)"', raw)
    if m is None:
        raise ValueError(f"{path}: not a dolfin XML mesh")
    dim = int(m.group(2))
    verts = _VERT_RE.findall(raw)
    cells = _CELL_RE.findall(raw)
    if not verts or not cells:
        raise ValueError(f"{path}: no vertices/cells found")
    x = np.zeros((len(verts), dim), dtype=np.float64)
    for v in verts:
        i = int(v[0])
        for d in range(dim):
            x[i, d] = float(v[1 + d])
    c = np.zeros((len(cells), dim + 1), dtype=np.int32)
    for t in cells:
        i = int(t[0])
        for d in range(dim + 1):
            c[i, d] = int(t[1 + d])
    name = os.path.basename(path)
    for suf in (".gz", ".xml"):
        if name.endswith(suf):
            name = name[: -len(suf)]
    return Mesh(x=x, cells=c, name=name)


def write_dolfin_xml(mesh: Mesh, path: str) -> None:
    """Write a mesh in the dolfin XML layout (used by tests and for refined meshes)."""
    ctype = {1: "interval", 2: "triangle", 3: "tetrahedron"}[mesh.dim]
    ax = "xyz"
    lines = ['<?xml version="1.0"?>', '<dolfin xmlns:dolfin="http://fenicsproject.org">',
             f'  <mesh celltype="{ctype}" dim="{mesh.dim}">',
             f'    <vertices size="{mesh.num_vertices}">']
    for i, p in enumerate(mesh.x):
        co = " ".join(f'{ax[d]}="{float(p[d])!r}"' for d in range(mesh.dim))
        lines.append(f'      <vertex index="{i}" {co} />')
    lines.append("    </vertices>")
    lines.append(f'    <cells size="{mesh.num_cells}">')
    for i, c in enumerate(mesh.cells):
        vs = " ".join(f'v{d}="{int(c[d])}"' for d in range(mesh.dim + 1))
        lines.append(f'      <{ctype} index="{i}" {vs} />')
    lines += ["    </cells>", "  </mesh>", "</dolfin>"]
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "wb") as f:
        f.write(("\n".join(lines) + "\n").encode())


def save_npz(mesh: Mesh, path: str) -> None:
    np.savez_compressed(path, x=mesh.x, cells=mesh.cells)


def load_npz(path: str) -> Mesh:
    d = np.load(path)
    return Mesh(x=np.ascontiguousarray(d["x"], dtype=np.float64),
                cells=np.ascontiguousarray(d["cells"], dtype=np.int32),
                name=os.path.basename(path)[:-4])


def load_mesh(name: str, utilities_dir: str | None = None) -> Mesh:
    """Load mesh ``name`` (file stem, e.g. ``L_50_R_5`` or
    ``1D_variable_50um_mesh_5990``).  Looks for the dolfin XML in
    ``utilities_dir`` first (drop-in with the reference's utilities folder),
    then for the packaged npz copy."""
    if utilities_dir:
        for ext in (".xml", ".xml.gz"):
            p = os.path.join(utilities_dir, name + ext)
            if os.path.exists(p):
                return read_dolfin_xml(p)
    p = os.path.join(_PKG_DATA, "meshes", name + ".npz")
    if os.path.exists(p):
        return load_npz(p)
    raise FileNotFoundError(
        f"mesh '{name}' not found in {utilities_dir or '(no utilities_dir)'} nor in {_PKG_DATA}/meshes")


def graded_interval(fine_cells: int, fine_len: float, coarse_cells: int) -> Mesh:
    """A graded [0,1] interval mesh with the structure of the reference's
    ``1D_variable_*`` files: ``fine_cells`` uniform cells on [0, fine_len], then
    ``coarse_cells`` uniform cells on [fine_len, 1]."""
    xf = np.linspace(0.0, fine_len, fine_cells + 1)
    xc = np.linspace(fine_len, 1.0, coarse_cells + 1)[1:]
    x = np.concatenate([xf, xc])[:, None]
    n = x.shape[0]
    cells = np.stack([np.arange(n - 1), np.arange(1, n)], axis=1).astype(np.int32)
    return Mesh(x=x, cells=cells, name=f"graded_{n - 1}")


# ---------------------------------------------------------------------------
# tetrahedral topology helpers (3D path; done once on the host)
# ---------------------------------------------------------------------------

def tet_edges(cells: np.ndarray) -> np.ndarray:
    """Unique undirected vertex pairs (i<j) of a tet mesh, sorted."""
    pairs = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
    e = np.concatenate([cells[:, p] for p in pairs], axis=0).astype(np.int64)
    e.sort(axis=1)
    return np.unique(e, axis=0)


def tet_facets(cells: np.ndarray):
    """All facets (sorted vertex triples), unique, with the number of cells that
    share each one (1 = exterior facet)."""
    tri = [(1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2)]
    f = np.concatenate([cells[:, t] for t in tri], axis=0).astype(np.int64)
    f.sort(axis=1)
    uf, cnt = np.unique(f, axis=0, return_counts=True)
    return uf, cnt


def red_refine(mesh: Mesh, project_radius: float | None = None) -> Mesh:
    """One level of uniform (red) refinement of a tet mesh: every tet is split
    into 8 (4 corner tets + 4 from the inner octahedron, shortest diagonal).
    New vertices on exterior wall facets can be projected to ``project_radius``
    (cylinder r) so that refined pores stay round (SURVEY §8d cfg 5)."""
    assert mesh.dim == 3
    x, c = mesh.x, mesh.cells.astype(np.int64)
    nv = x.shape[0]
    edges = tet_edges(c)
    key = edges[:, 0] * nv + edges[:, 1]

    def mid(a, b):
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        return nv + np.searchsorted(key, lo * nv + hi)

    xm = 0.5 * (x[edges[:, 0]] + x[edges[:, 1]])
    if project_radius is not None:
        r0 = np.hypot(x[edges[:, 0], 0], x[edges[:, 0], 1])
        r1 = np.hypot(x[edges[:, 1], 0], x[edges[:, 1], 1])
        on_wall = (np.abs(r0 - project_radius) < 1e-9) & (np.abs(r1 - project_radius) < 1e-9)
        # only exterior edges: both end points on the wall and the edge lies on an exterior facet
        uf, cnt = tet_facets(c)
        ext = uf[cnt == 1]
        ek = np.concatenate([ext[:, [0, 1]], ext[:, [0, 2]], ext[:, [1, 2]]], axis=0)
        ekey = np.unique(ek[:, 0] * nv + ek[:, 1])
        is_ext = np.isin(key, ekey)
        sel = on_wall & is_ext
        rm = np.hypot(xm[sel, 0], xm[sel, 1])
        xm[sel, 0] *= project_radius / rm
        xm[sel, 1] *= project_radius / rm
    xn = np.concatenate([x, xm], axis=0)
    v0, v1, v2, v3 = c[:, 0], c[:, 1], c[:, 2], c[:, 3]
    m01, m02, m03 = mid(v0, v1), mid(v0, v2), mid(v0, v3)
    m12, m13, m23 = mid(v1, v2), mid(v1, v3), mid(v2, v3)
    corner = [np.stack(t, 1) for t in ((v0, m01, m02, m03), (m01, v1, m12, m13),
                                       (m02, m12, v2, m23), (m03, m13, m23, v3))]
    # inner octahedron split along the diagonal m02-m13 (fixed choice; quality is adequate
    # for the benchmark meshes and keeps the refinement deterministic)
    octa = [np.stack(t, 1) for t in ((m01, m02, m03, m13), (m01, m02, m12, m13),
                                     (m02, m03, m13, m23), (m02, m12, m13, m23))]
    cn = np.concatenate(corner + octa, axis=0).astype(np.int32)
    return Mesh(x=xn, cells=cn, name=mesh.name + "_r")
