"""
Host-side simplex mesh container standing in for ``dolfin.Mesh`` / ``dolfin.MeshFunction``.

The reference hands a ``(mesh, mesh_functions, mesh_subdomains)`` tuple to its
residual constructors (``/root/reference/src/femvf/load.py:45-54``,
``tests/fixture_mesh.py:104-116``).  DOLFIN is not available, so this module
provides the few members of those objects that the hot path and ``StateFile``
touch (``coordinates()``, ``cells()``, ``topology().dim()``)
plus the topology (facets, edges, parent cells, outward normals) that DOLFIN
would otherwise supply.

DOF numbering convention: scalar P1 DOF ``i`` is mesh
vertex ``i``; vector DOFs are node-major interleaved, ``(d*i, ..., d*i+d-1)``.
"""

from __future__ import annotations

import numpy as np


class _Topology:
    def __init__(self, dim: int):
        self._dim = dim

    def dim(self) -> int:
        return self._dim


class MeshFunction:
    """Integer marker per mesh entity of one topological dimension."""

    def __init__(self, dim: int, values):
        self._dim = int(dim)
        self._values = np.asarray(values, dtype=np.int64).copy()

    def dim(self) -> int:
        return self._dim

    def array(self) -> np.ndarray:
        return self._values

    def where_equal(self, value: int) -> np.ndarray:
        return np.nonzero(self._values == value)[0]

    def __getitem__(self, idx):
        return self._values[idx]

    def __setitem__(self, idx, value):
        self._values[idx] = value

    def __len__(self):
        return self._values.size


def _unique_rows(a: np.ndarray):
    """Unique rows of an int array with inverse map (rows are pre-sorted per row)."""
    a = np.ascontiguousarray(a)
    order = np.lexsort(a.T[::-1])
    s = a[order]
    new = np.ones(len(s), dtype=bool)
    if len(s) > 1:
        new[1:] = np.any(s[1:] != s[:-1], axis=1)
    ids = np.cumsum(new) - 1
    inverse = np.empty(len(a), dtype=np.int64)
    inverse[order] = ids
    return s[new], inverse


class Mesh:
    """
    P1 simplex mesh (triangles or tetrahedra).

    Entities are numbered as: vertices (rows of ``coordinates()``), facets
    (codim 1, rows of ``facets``; sorted vertex ids, lexicographic order),
    edges (dim 1; equal to facets in 2D), cells (rows of ``cells()``).
    """

    def __init__(self, coords, cells):
        self._coords = np.ascontiguousarray(coords, dtype=np.float64)
        self._cells = np.ascontiguousarray(cells, dtype=np.int64)
        d = self._coords.shape[1]
        if self._cells.shape[1] != d + 1:
            raise ValueError("cells must have dim+1 vertices (P1 simplices)")
        self._topology = _Topology(d)
        self._facets = None
        self._edges = None
        self._orient_cells()

    def _orient_cells(self):
        """Make every cell positively oriented (positive signed volume)."""
        vol = self.signed_volumes()
        neg = vol < 0
        if np.any(neg):
            c = self._cells
            c[neg, 0], c[neg, 1] = c[neg, 1].copy(), c[neg, 0].copy()
        if np.any(self.signed_volumes() <= 0):
            raise ValueError("degenerate cell in mesh")

    # --- dolfin-like accessors -------------------------------------------------
    def coordinates(self) -> np.ndarray:
        return self._coords

    def cells(self) -> np.ndarray:
        return self._cells

    def topology(self) -> _Topology:
        return self._topology

    def num_vertices(self) -> int:
        return self._coords.shape[0]

    def num_cells(self) -> int:
        return self._cells.shape[0]

    # --- geometry ---------------------------------------------------------------
    def signed_volumes(self) -> np.ndarray:
        x = self._coords[self._cells]
        e = x[:, 1:, :] - x[:, :1, :]
        d = self._topology.dim()
        det = np.linalg.det(e)
        return det / (2.0 if d == 2 else 6.0)

    # --- topology ---------------------------------------------------------------
    def _build_facets(self):
        d = self._topology.dim()
        cells = self._cells
        ne = cells.shape[0]
        # local facet k is the one opposite local vertex k
        loc = [[j for j in range(d + 1) if j != k] for k in range(d + 1)]
        allf = np.concatenate([cells[:, l] for l in loc], axis=0)  # ((d+1)*ne, d)
        allf_sorted = np.sort(allf, axis=1)
        facets, inv = _unique_rows(allf_sorted)
        nf = facets.shape[0]
        cell_of = np.tile(np.arange(ne), d + 1)
        opp_of = np.repeat(np.arange(d + 1), ne)
        facet_cells = -np.ones((nf, 2), dtype=np.int64)
        facet_opp = -np.ones((nf, 2), dtype=np.int64)
        order = np.argsort(inv, kind='stable')
        inv_s = inv[order]
        first = np.ones(len(inv_s), dtype=bool)
        first[1:] = inv_s[1:] != inv_s[:-1]
        facet_cells[inv_s[first], 0] = cell_of[order][first]
        facet_opp[inv_s[first], 0] = opp_of[order][first]
        second = ~first
        facet_cells[inv_s[second], 1] = cell_of[order][second]
        facet_opp[inv_s[second], 1] = opp_of[order][second]
        self._facets = facets
        self._facet_cells = facet_cells
        self._facet_opp = facet_opp
        self._cell_facets = inv.reshape(d + 1, ne).T.copy()  # (ne, d+1)

    @property
    def facets(self) -> np.ndarray:
        """(nf, d) sorted vertex ids of every codim-1 entity."""
        if self._facets is None:
            self._build_facets()
        return self._facets

    @property
    def facet_cells(self) -> np.ndarray:
        """(nf, 2) incident cells, second is -1 for exterior facets."""
        if self._facets is None:
            self._build_facets()
        return self._facet_cells

    @property
    def facet_opposite(self) -> np.ndarray:
        """(nf, 2) local index (in the incident cell) of the vertex opposite the facet."""
        if self._facets is None:
            self._build_facets()
        return self._facet_opp

    @property
    def cell_facets(self) -> np.ndarray:
        if self._facets is None:
            self._build_facets()
        return self._cell_facets

    @property
    def exterior_facets(self) -> np.ndarray:
        return np.nonzero(self.facet_cells[:, 1] < 0)[0]

    @property
    def edges(self) -> np.ndarray:
        """(nedge, 2) sorted vertex ids of every dim-1 entity."""
        if self._topology.dim() == 2:
            return self.facets
        if self._edges is None:
            c = self._cells
            pairs = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
            alle = np.concatenate([c[:, p] for p in pairs], axis=0)
            self._edges, _ = _unique_rows(np.sort(alle, axis=1))
        return self._edges

    def num_entities(self, dim: int) -> int:
        d = self._topology.dim()
        if dim == 0:
            return self.num_vertices()
        if dim == d:
            return self.num_cells()
        if dim == d - 1:
            return self.facets.shape[0]
        if dim == 1:
            return self.edges.shape[0]
        raise ValueError(f"invalid entity dimension {dim}")

    def facet_midpoints(self) -> np.ndarray:
        return self._coords[self.facets].mean(axis=1)

    def mark_facets(self, predicate, value: int, mf: MeshFunction, on_boundary_only=True):
        """Mark facets whose vertices all satisfy ``predicate(x)`` (DOLFIN ``SubDomain.mark``)."""
        f = self.facets
        x = self._coords[f]  # (nf, d, dim)
        inside = np.all(predicate(x.reshape(-1, x.shape[-1])).reshape(f.shape[0], -1), axis=1)
        inside &= predicate(x.mean(axis=1))
        if on_boundary_only:
            inside &= self.facet_cells[:, 1] < 0
        mf.array()[inside] = value


def unit_square_mesh(nx: int, ny: int) -> Mesh:
    """Structured triangle mesh of [0,1]^2, vertex order and 'right' diagonals as
    ``dolfin.UnitSquareMesh(nx, ny)`` (fixture of ``tests/fixture_mesh.py:33-38``)."""
    xs = np.linspace(0.0, 1.0, nx + 1)
    ys = np.linspace(0.0, 1.0, ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing='xy')
    coords = np.stack([X.ravel(), Y.ravel()], axis=1)
    cells = []
    for iy in range(ny):
        for ix in range(nx):
            v0 = iy * (nx + 1) + ix
            v1 = v0 + 1
            v2 = v0 + (nx + 1)
            v3 = v2 + 1
            cells.append((v0, v1, v3))
            cells.append((v0, v2, v3))
    return Mesh(coords, np.array(cells))


def unit_cube_mesh(nx: int, ny: int, nz: int) -> Mesh:
    """Structured tet mesh of [0,1]^3 with 6 tets per box (``dolfin.UnitCubeMesh`` layout)."""
    xs = np.linspace(0.0, 1.0, nx + 1)
    ys = np.linspace(0.0, 1.0, ny + 1)
    zs = np.linspace(0.0, 1.0, nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing='ij')
    coords = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    cells = []
    for iz in range(nz):
        for iy in range(ny):
            for ix in range(nx):
                v0 = iz * (nx + 1) * (ny + 1) + iy * (nx + 1) + ix
                v1 = v0 + 1
                v2 = v0 + (nx + 1)
                v3 = v1 + (nx + 1)
                v4 = v0 + (nx + 1) * (ny + 1)
                v5 = v1 + (nx + 1) * (ny + 1)
                v6 = v2 + (nx + 1) * (ny + 1)
                v7 = v3 + (nx + 1) * (ny + 1)
                cells += [
                    (v0, v1, v3, v7), (v0, v1, v7, v5), (v0, v5, v7, v4),
                    (v0, v3, v2, v7), (v0, v6, v4, v7), (v0, v2, v6, v7),
                ]
    return Mesh(coords, np.array(cells))


def fixture_mesh_tuple(mesh: Mesh):
    """
    The marking rules of the reference's test fixture (``tests/fixture_mesh.py:48-116``):
    facets on y=0 (and z=0 / z=1 in 3D) -> 'fixed'=1, every other facet keeps 0 =
    'pressure' (interior facets included, quirk Q8); cells with y>0.5 -> 'top'=1;
    the codim-2 entity at the top-right corner -> 'separation'=1.
    """
    d = mesh.topology().dim()
    eps = 3.0e-16
    facet_mf = MeshFunction(d - 1, np.zeros(mesh.num_entities(d - 1), dtype=np.int64))

    def fixed(x):
        out = x[:, 1] < eps
        if d == 3:
            out = out | (x[:, 2] > 1 - eps) | (x[:, 2] < eps)
        return out

    mesh.mark_facets(fixed, 1, facet_mf)

    cell_mf = MeshFunction(d, np.zeros(mesh.num_cells(), dtype=np.int64))
    xc = mesh.coordinates()[mesh.cells()]
    top = np.all(xc[:, :, 1] > 0.5 + eps, axis=1) & (xc.mean(axis=1)[:, 1] > 0.5 + eps)
    cell_mf.array()[top] = 1

    ncodim2 = mesh.num_entities(d - 2)
    codim2_mf = MeshFunction(d - 2, np.zeros(ncodim2, dtype=np.int64))
    if d == 2:
        x = mesh.coordinates()
        codim2_mf.array()[(x[:, 1] > 1 - eps) & (x[:, 0] > 1 - eps)] = 1
    else:
        e = mesh.edges
        xe = mesh.coordinates()[e]
        on = np.all((xe[:, :, 1] > 1 - eps) & (xe[:, :, 0] > 1 - eps), axis=1)
        codim2_mf.array()[on] = 1

    mesh_functions = (d - 2) * (None,) + (codim2_mf, facet_mf, cell_mf)
    mesh_subdomains = (d - 2) * ({},) + (
        {'separation': 1}, {'fixed': 1, 'pressure': 0}, {'top': 1, 'bottom': 0}
    )
    return mesh, mesh_functions, mesh_subdomains
