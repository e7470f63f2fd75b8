"""
Mesh utilities: mirror of the parts of ``/root/reference/src/femvf/meshutils.py`` on the
setup path of the hot path (entity filters ``:171-260``, 1D edge-loop extraction
``:267-334``).  gmsh/meshio loading (``:63-167``) is not available here.
"""

from __future__ import annotations

from typing import Callable, Iterable

import numpy as np

from .mesh import Mesh, MeshFunction
from .residuals.base import mesh_element_type_dim  # noqa: F401  (re-export, meshutils.py:24)


def load_fenics_gmsh(mesh_path: str):
    raise NotImplementedError(
        "gmsh/meshio are not available in this environment; pass a "
        "(mesh, mesh_functions, mesh_subdomains) tuple (see femvf_b200.meshgen) instead")


def filter_mesh_entities(mesh_entities: Iterable[int], filter: Callable[[int], bool]):
    return [ent for ent in mesh_entities if filter(ent)]


def edges_incident_to_facets(mesh: Mesh, facet_function: MeshFunction, facet_values: set):
    """Edge ids incident to any facet whose marker is in ``facet_values``
    (``filter_mesh_entities_by_subdomain``, ``meshutils.py:171-211``)."""
    d = mesh.topology().dim()
    tags = facet_function.array()
    sel = np.isin(tags, np.fromiter(facet_values, dtype=np.int64))
    if d == 2:
        return np.nonzero(sel)[0]
    f = mesh.facets[sel]
    pairs = np.concatenate([f[:, [0, 1]], f[:, [0, 2]], f[:, [1, 2]]], axis=0)
    pairs = np.unique(np.sort(pairs, axis=1), axis=0)
    e = mesh.edges
    base = mesh.num_vertices() + 1
    ekey = e[:, 0] * base + e[:, 1]
    pkey = pairs[:, 0] * base + pairs[:, 1]
    return np.nonzero(np.isin(ekey, pkey))[0]


def edges_on_plane(mesh: Mesh, edges: np.ndarray, origin, normal):
    """Subset of ``edges`` with midpoints on a plane (``meshutils.py:213-237``)."""
    x = mesh.coordinates()
    mid = x[mesh.edges[edges]].mean(axis=1)
    if mid.shape[1] == 2:
        mid = np.concatenate([mid, np.zeros((len(mid), 1))], axis=1)
    dist = (mid - np.asarray(origin, dtype=float)) @ np.asarray(normal, dtype=float)
    return edges[np.isclose(dist, 0)]


def vertices_from_edges(mesh: Mesh, edges: np.ndarray) -> np.ndarray:
    return np.unique(mesh.edges[edges].reshape(-1))


def sort_vertices_by_nearest_neighbours(vertex_coordinates: np.ndarray, origin=None) -> np.ndarray:
    """Greedy nearest-neighbour ordering from the point closest to ``origin``
    (``meshutils.py:295-334``)."""
    x = np.asarray(vertex_coordinates, dtype=float)
    origin = np.zeros(x.shape[-1]) if origin is None else origin
    idx_sort = [int(np.argmin(np.linalg.norm(x - origin, axis=-1)))]
    visited = np.zeros(x.shape[0], dtype=bool)
    visited[idx_sort[0]] = True
    while len(idx_sort) < x.shape[0]:
        dist = np.sum((x - x[idx_sort[-1]]) ** 2, axis=-1) ** 0.5
        dist[visited] = np.nan
        nxt = int(np.nanargmin(dist))
        idx_sort.append(nxt)
        visited[nxt] = True
    return np.array(idx_sort)


def sort_edge_vertices(mesh: Mesh, edges: np.ndarray):
    """Sorted coordinates and vertex ids of a set of connected edges (``meshutils.py:267-284``)."""
    vertices = vertices_from_edges(mesh, edges)
    surface_coordinates = mesh.coordinates()[vertices]
    idx_sort = sort_vertices_by_nearest_neighbours(surface_coordinates)
    return surface_coordinates[idx_sort], vertices[idx_sort]
