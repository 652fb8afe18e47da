"""Mesh partitioning for the mesh-partitioned 3D mode (BASELINE config 5).

The reference has no distributed path (no MPI usage; dolfin's own partitioner is never
engaged by the scripts).  Here one refined pore mesh is cut into ``world`` slabs along the pore axis z
(the cylinder axis; NVSwitch makes the neighbour choice irrelevant, so the simplest cut with the
smallest interfaces is used):

* every vertex has exactly one OWNER rank (contiguous ranges of the z-sorted vertex list, so all
  ranks own the same number of block rows +- 1);
* a rank's LOCAL mesh = all tets that touch an owned vertex; its local vertex numbering puts the
  owned vertices first -- INTERIOR ones (no ghost neighbour) before BOUNDARY ones, ascending global
  id inside each group -- and the GHOST vertices (vertices of local tets owned by other ranks) after
  them, so "the first n_own rows" is the owned part of every local vector and "the first n_int rows"
  can be multiplied while the halo exchange is still in flight;
* the rows of owned vertices are therefore assembled completely from local tets (no exchange of
  matrix entries); ghost rows are incomplete and never used;
* per neighbour, ``send`` lists the owned vertices that are ghosts over there and ``recv`` the local
  ghost slots they fill -- both ordered by global id, so the two sides agree without negotiation.

Everything here is NumPy on the host, computed once per mesh; every rank runs the same deterministic
code on the full mesh and keeps its own part.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import meshio


@dataclass
class MeshPart:
    rank: int
    world: int
    glob: np.ndarray                 # [n_local] global vertex id of every local vertex (owned first)
    n_own: int
    n_int: int                       # the first n_int owned vertices have no ghost neighbour (interior rows)
    x: np.ndarray                    # [n_local, 3]
    cells: np.ndarray                # [t_local, 4] in local numbering
    cell_glob: np.ndarray            # [t_local] global tet ids
    send: dict = field(default_factory=dict)     # nbr rank -> local indices (owned) to send
    recv: dict = field(default_factory=dict)     # nbr rank -> local indices (ghost) to fill

    @property
    def n_local(self) -> int:
        return int(self.glob.shape[0])

    @property
    def n_ghost(self) -> int:
        return self.n_local - self.n_own

    def local_mesh(self) -> meshio.Mesh:
        return meshio.Mesh(x=self.x, cells=self.cells, name=f"part{self.rank}of{self.world}")

    def halo_doubles(self, ncomp: int = 9) -> int:
        """Doubles this rank receives per halo exchange."""
        return ncomp * sum(len(v) for v in self.recv.values())


def vertex_owners(x: np.ndarray, world: int) -> np.ndarray:
    """Owner rank of every vertex: equal-count slabs of the (z, id)-sorted vertex list."""
    nv = x.shape[0]
    order = np.lexsort((np.arange(nv), x[:, 2]))
    owner = np.empty(nv, dtype=np.int32)
    owner[order] = (np.arange(nv, dtype=np.int64) * world // nv).astype(np.int32)
    return owner


def partition_z(mesh: meshio.Mesh, world: int, ranks=None) -> list:
    """Cut ``mesh`` into ``world`` z-slabs; returns the :class:`MeshPart` of every rank in ``ranks``
    (default: all)."""
    assert mesh.dim == 3 and world >= 1
    x = np.asarray(mesh.x, dtype=np.float64)
    cells = np.asarray(mesh.cells, dtype=np.int64)
    owner = vertex_owners(x, world)
    cell_owner = owner[cells]                                     # [T, 4]
    ranks = list(range(world)) if ranks is None else list(ranks)
    # ghost sets of ALL ranks are needed to build the send lists: local vertices of rank s that s does not own
    local_verts = {}
    local_cells = {}
    for s in range(world):
        tsel = np.nonzero((cell_owner == s).any(axis=1))[0]
        local_cells[s] = tsel
        local_verts[s] = np.unique(cells[tsel])
    parts = []
    for r in ranks:
        lv = local_verts[r]
        own = np.nonzero(owner == r)[0]                          # ascending global id (incl. isolated vertices)
        ghost = lv[owner[lv] != r]
        # boundary-owned vertices: share a local tet with a ghost vertex
        tc = cells[local_cells[r]]
        has_ghost = (cell_owner[local_cells[r]] != r).any(axis=1)
        bverts = np.unique(tc[has_ghost])
        is_b = np.zeros(x.shape[0], dtype=bool)
        is_b[bverts] = True
        own = np.concatenate([own[~is_b[own]], own[is_b[own]]])
        n_int = int((~is_b[own]).sum())
        glob = np.concatenate([own, ghost]).astype(np.int64)
        g2l = -np.ones(x.shape[0], dtype=np.int64)
        g2l[glob] = np.arange(len(glob))
        lc = g2l[cells[local_cells[r]]]
        assert (lc >= 0).all()
        part = MeshPart(rank=r, world=world, glob=glob, n_own=int(len(own)), n_int=n_int, x=x[glob].copy(),
                        cells=lc.astype(np.int32), cell_glob=local_cells[r].astype(np.int64))
        # receive: my ghosts grouped by owner (ascending global id inside a group)
        for s in np.unique(owner[ghost]):
            gs = ghost[owner[ghost] == s]
            part.recv[int(s)] = g2l[gs]
        # send: my owned vertices that are ghosts on rank s
        for s in range(world):
            if s == r:
                continue
            lvs = local_verts[s]
            mine = lvs[owner[lvs] == r]
            if len(mine):
                part.send[int(s)] = g2l[mine]
        parts.append(part)
    return parts


def local_dirichlet(part: MeshPart, dofs: np.ndarray, ncomp: int = 9):
    """Restrict the global Dirichlet DOF list (``marking.dirichlet_sets``) to a part.
    Returns (local_dofs[int32] sorted, index into the global list for the values)."""
    dofs = np.asarray(dofs, dtype=np.int64)
    gv, comp = dofs // ncomp, dofs % ncomp
    g2l = -np.ones(int(max(part.glob.max(), gv.max() if len(gv) else 0)) + 1, dtype=np.int64)
    g2l[part.glob] = np.arange(part.n_local)
    lv = g2l[gv]
    sel = np.nonzero(lv >= 0)[0]
    ld = lv[sel] * ncomp + comp[sel]
    order = np.argsort(ld, kind="stable")
    return ld[order].astype(np.int32), sel[order]


def local_facets(part: MeshPart, facets: np.ndarray, n_global: int):
    """Boundary facets (global vertex ids, [nf, 3]) a part must integrate: those with at least one OWNED vertex --
    the tet behind such a facet touches an owned vertex, so it is local and all three vertices exist locally.  Every
    owned row then receives every facet that touches it; ghost rows stay incomplete and are never used.
    Returns (local vertex ids [k, 3] int32, indices into ``facets``)."""
    g2l = -np.ones(n_global, dtype=np.int64)
    g2l[part.glob] = np.arange(part.n_local)
    lf = g2l[np.asarray(facets, dtype=np.int64).reshape(-1, 3)]
    keep = (lf >= 0).all(axis=1) & ((lf >= 0) & (lf < part.n_own)).any(axis=1)
    return lf[keep].astype(np.int32), np.nonzero(keep)[0]


def scatter_to_part(part: MeshPart, xg: np.ndarray) -> np.ndarray:
    """Global nodal array [nv, ...] -> local array [n_local, ...] (owned and ghost slots filled)."""
    return np.ascontiguousarray(xg[part.glob])


def gather_owned(parts, xs, nv: int) -> np.ndarray:
    """Owned rows of every part's local array -> global array."""
    out = np.zeros((nv,) + tuple(xs[0].shape[1:]), dtype=xs[0].dtype)
    for p, xl in zip(parts, xs):
        out[p.glob[: p.n_own]] = xl[: p.n_own]
    return out
