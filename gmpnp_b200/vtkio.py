"""ParaView output of nodal P1 fields: what ``File('solution_X.pvd') << u_X`` leaves behind in the reference
(3D/MPNP_CO2ER_pore.py:863-880, 3D/rxn_diff_CO2ER_pore.py:619-632).

dolfin's VTK writer produces a ``.pvd`` collection that points at ``<stem>000000.vtu`` -- an ASCII unstructured grid
with the (dimensionless) mesh coordinates, the tets (VTK cell type 10) and one Float64 point-data array.  The same two
files are written here with plain Python; ParaView/VisIt read them as they read dolfin's.
"""
from __future__ import annotations

import os

import numpy as np

VTK_TETRA, VTK_LINE = 10, 3


def write_vtu(path, x, cells, point_data: dict):
    """One ``.vtu`` piece: ``x[nv, dim]``, ``cells[nc, 2 or 4]``, ``point_data`` name -> values[nv]."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    cells = np.asarray(cells, dtype=np.int64)
    nv, nc, npc = x.shape[0], cells.shape[0], cells.shape[1]
    ctype = {4: VTK_TETRA, 2: VTK_LINE}[npc]
    x3 = np.zeros((nv, 3))
    x3[:, : x.shape[1]] = x
    first = next(iter(point_data))
    with open(path, "w") as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid"  version="0.1"  >\n<UnstructuredGrid>\n')
        f.write(f'<Piece  NumberOfPoints="{nv}" NumberOfCells="{nc}">\n')
        f.write('<Points>\n<DataArray  type="Float64"  NumberOfComponents="3"  format="ascii">')
        f.write("  ".join(" ".join(repr(float(c)) for c in row) for row in x3))
        f.write('</DataArray>\n</Points>\n<Cells>\n<DataArray  type="UInt32"  Name="connectivity"  format="ascii">')
        f.write("  ".join(" ".join(str(int(v)) for v in row) for row in cells))
        f.write('</DataArray>\n<DataArray  type="UInt32"  Name="offsets"  format="ascii">')
        f.write(" ".join(str(npc * (i + 1)) for i in range(nc)))
        f.write('</DataArray>\n<DataArray  type="UInt8"  Name="types"  format="ascii">')
        f.write(" ".join([str(ctype)] * nc))
        f.write(f'</DataArray>\n</Cells>\n<PointData  Scalars="{first}">\n')
        for name, vals in point_data.items():
            vals = np.asarray(vals, dtype=np.float64).reshape(-1)
            assert vals.shape[0] == nv, (name, vals.shape, nv)
            f.write(f'<DataArray  type="Float64"  Name="{name}"  format="ascii">')
            f.write("  ".join(repr(float(v)) for v in vals))
            f.write("</DataArray>\n")
        f.write("</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>")


def write_pvd(path, x, cells, values, name="f", timestep=0):
    """``File(path) << u``: ``path`` ends in .pvd; writes it and ``<stem>000000.vtu`` next to it.  Returns both paths."""
    assert path.endswith(".pvd"), path
    stem = path[:-4]
    vtu = stem + "%06d.vtu" % 0
    write_vtu(vtu, x, cells, {name: values})
    with open(path, "w") as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="Collection" version="0.1">\n  <Collection>\n')
        f.write(f'    <DataSet timestep="{timestep}" part="0" file="{os.path.basename(vtu)}" />\n')
        f.write("  </Collection>\n</VTKFile>\n")
    return path, vtu


def read_vtu_point_data(path):
    """Minimal reader for the files above (tests, post-processing): (points[nv,3], cells[nc,npc], {name: values})."""
    import xml.etree.ElementTree as ET
    root = ET.parse(path).getroot()
    piece = root.find("UnstructuredGrid").find("Piece")
    nv, nc = int(piece.get("NumberOfPoints")), int(piece.get("NumberOfCells"))
    pts = np.array(piece.find("Points").find("DataArray").text.split(), dtype=np.float64).reshape(nv, 3)
    arrs = {a.get("Name"): a for a in piece.find("Cells").findall("DataArray")}
    conn = np.array(arrs["connectivity"].text.split(), dtype=np.int64)
    cells = conn.reshape(nc, -1)
    data = {a.get("Name"): np.array(a.text.split(), dtype=np.float64) for a in piece.find("PointData").findall("DataArray")}
    return pts, cells, data
