"""
Model construction: mirror of ``/root/reference/src/femvf/load.py``.

``load_fsi_model`` keeps the reference signature (``load.py:100-110``).  The solid mesh is a
``(mesh, mesh_functions, mesh_subdomains)`` tuple (``load.py:45-54``) with this package's
``Mesh``/``MeshFunction`` types; ``.msh`` paths are rejected because gmsh/meshio are absent.
"""

from __future__ import annotations

from os import path
from typing import Any, Optional

import numpy as np

from . import meshutils
from .mesh import Mesh
from .residuals import solid as slr, fluid as flr
from .models import transient


def load_fenics_model(mesh, Residual, model_type: str = 'transient', **kwargs):
    """Load a solid model (``load.py:23-65``)."""
    if isinstance(mesh, str):
        ext = path.splitext(mesh)[1]
        if ext.lower() == '.msh':
            mesh, mesh_funcs, mesh_subdomains = meshutils.load_fenics_gmsh(mesh)
        else:
            raise ValueError(f"Invalid mesh extension {ext}")
    elif isinstance(mesh, (tuple, list)):
        mesh, mesh_funcs, mesh_subdomains = mesh
    else:
        raise TypeError(f"Invalid `mesh` type {type(mesh)}")

    residual = Residual(mesh, mesh_funcs, mesh_subdomains, **kwargs)
    if model_type == 'transient':
        # NOTE: as in the reference this is always the plain model, never the
        # NodalContactModel (load.py:57-58)
        return transient.FenicsModel(residual)
    elif model_type in ('dynamical', 'linearized_dynamical'):
        raise NotImplementedError(
            "dynamical models are outside the accelerated path (SURVEY.md section 2)")
    else:
        raise ValueError(f"Invalid model type {model_type}")


def load_jax_model(mesh: np.ndarray, Residual, model_type: str = 'transient', **kwargs):
    """Load a 1D fluid model (``load.py:67-97``)."""
    residual = Residual(mesh, **kwargs)
    if model_type == 'transient':
        return transient.JaxModel(residual)
    elif model_type in ('dynamical', 'linearized_dynamical'):
        raise NotImplementedError(
            "dynamical models are outside the accelerated path (SURVEY.md section 2)")
    else:
        raise ValueError(f"Invalid model type {model_type}")


def load_fsi_model(
    solid_mesh,
    SolidResidual,
    FluidResidual,
    solid_kwargs: dict[str, Any],
    fluid_kwargs: dict[str, Any],
    model_type: str = 'transient',
    coupling: str = 'explicit',
    fluid_interface_subdomains: Optional[tuple] = ('pressure',),
    zs: Optional[np.ndarray] = None,
):
    """Load a coupled (fsi) model (``load.py:100-162``)."""
    solid = load_fenics_model(solid_mesh, SolidResidual, model_type=model_type, **solid_kwargs)

    mesh = solid.residual.mesh()
    facet_func = solid.residual.mesh_function('facet')
    filter_facet_values = set(
        solid.residual.mesh_subdomain('facet')[name] for name in fluid_interface_subdomains
    )
    pressure_function_space = solid.residual.form['control/p1'].function_space()

    s, dofs_fsi_solid, dofs_fsi_fluid = derive_1D_interface_from_facet_subdomain(
        mesh, pressure_function_space, facet_func, filter_facet_values, zs
    )

    fluid = load_jax_model(s, FluidResidual, model_type=model_type, **fluid_kwargs)

    if model_type == 'transient' and coupling == 'explicit':
        FSIModel = transient.ExplicitFSIModel
    elif model_type == 'transient' and coupling == 'implicit':
        FSIModel = transient.ImplicitFSIModel
    else:
        raise ValueError(f"Invalid `model_type` and `coupling` ({model_type}, {coupling})")

    return FSIModel(solid, fluid, dofs_fsi_solid, dofs_fsi_fluid)


def derive_1D_interface_from_facet_subdomain(mesh: Mesh, function_space, facet_function,
                                             facet_values: set, zs=None):
    """1D edge mesh and interface DOF arrays from a facet subdomain (``load.py:164-214``).

    ``vertex_to_dof_map`` of the scalar P1 space is the identity in this package's DOF
    numbering."""
    interface_coords, interface_vertices = derive_edge_mesh_from_facet_subdomain(
        mesh, facet_function, facet_values, zs
    )
    # the reference indexes with ``interface_vertices.flat`` (load.py:204-214): solid_dofs is
    # 1-D, so fluid_dofs = arange(solid_dofs.size) -- for nz z-planes of ns points, plane k owns
    # the fluid DOFs k*ns .. (k+1)*ns - 1 of the (nz, ns) fluid state
    solid_dofs = np.asarray(interface_vertices, dtype=np.int64).reshape(-1)
    fluid_dofs = np.arange(solid_dofs.size, dtype=int)
    return interface_coords, solid_dofs, fluid_dofs


def derive_edge_mesh_from_facet_subdomain(mesh: Mesh, facet_function, facet_values: set, zs=None):
    """``load.py:216-281``."""
    dim = mesh.topology().dim()
    fsi_all = meshutils.edges_incident_to_facets(mesh, facet_function, facet_values)
    if dim == 2:
        coords, vertices = derive_edge_mesh_from_edges(mesh, fsi_all)
    elif dim == 3 and zs is not None:
        mesh_list = [
            derive_edge_mesh_from_edges(
                mesh, meshutils.edges_on_plane(mesh, fsi_all, np.array([0, 0, z]),
                                               np.array([0, 0, 1])))
            for z in zs
        ]
        coords = np.array([c for c, _ in mesh_list])
        vertices = np.array([v for _, v in mesh_list], dtype=int)
    elif dim == 3 and zs is None:
        raise ValueError("`zs` must be an array for a 3D mesh")
    else:
        raise ValueError(f"Invalid mesh dimension {dim}")
    return coords, vertices


def derive_edge_mesh_from_edges(mesh: Mesh, edges: np.ndarray):
    """Arclength coordinate ``s`` and ordered vertices of an edge chain (``load.py:283-293``)."""
    vertex_coords, fsi_verts = meshutils.sort_edge_vertices(mesh, edges)
    dxyz = vertex_coords[1:] - vertex_coords[:-1]
    dx, dy = dxyz[:, 0], dxyz[:, 1]
    s = np.concatenate([[0], np.cumsum(np.sqrt(dx**2 + dy**2))])
    return s, fsi_verts
