"""
Base residual classes: mirror of ``/root/reference/src/femvf/residuals/base.py``.

The reference's ``FenicsResidual`` couples a UFL form with a mesh, mesh functions and
Dirichlet conditions (``base.py:23-65``).  Here the "form" is a description of which
predefined terms are present plus the coefficient vectors; the integrals themselves are
evaluated by the CUDA element kernels (``csrc/elem.cuh``).
"""

from __future__ import annotations

from typing import Any, Callable, Optional, Union

import numpy as np

from ..mesh import Mesh, MeshFunction

ELEMENT_TYPE_TO_IDX = {'vertex': 0, 'edge': 1, 'facet': -2, 'cell': -1}


def mesh_element_type_dim(element_type: Union[str, int]) -> int:
    """Index into the per-dimension lists of a mesh tuple (``meshutils.py:24-58``)."""
    if isinstance(element_type, str):
        if element_type not in ELEMENT_TYPE_TO_IDX:
            raise ValueError(
                f"`mesh_element_type` must be one of {ELEMENT_TYPE_TO_IDX.keys()}`")
        return ELEMENT_TYPE_TO_IDX[element_type]
    if isinstance(element_type, int):
        return element_type
    raise TypeError(
        f"`mesh_element_type` must be `str` or `int`, not `{type(element_type)}`")


class _DofMap:
    def __init__(self, space: 'FunctionSpace'):
        self._space = space

    def cell_dofs(self, idx_cell: int) -> np.ndarray:
        sp = self._space
        cell = sp.mesh().cells()[idx_cell]
        if sp.family == 'DG':
            return np.array([idx_cell])
        n = sp.value_size
        return (n * cell[:, None] + np.arange(n)[None, :]).reshape(-1)


class FunctionSpace:
    """P1 ('CG', 1) scalar/vector, DG0 scalar or real ('R') space on a mesh."""

    def __init__(self, mesh: Mesh, family: str, degree: int, value_size: int = 1):
        self._mesh = mesh
        self.family = family
        self.degree = degree
        self.value_size = value_size

    def mesh(self) -> Mesh:
        return self._mesh

    def dim(self) -> int:
        if self.family == 'CG':
            return self.value_size * self._mesh.num_vertices()
        if self.family == 'DG':
            return self._mesh.num_cells()
        return self.value_size

    def num_sub_spaces(self) -> int:
        return self.value_size if self.value_size > 1 else 0

    def dofmap(self) -> _DofMap:
        return _DofMap(self)

    def tabulate_dof_coordinates(self) -> np.ndarray:
        if self.family == 'CG':
            return np.repeat(self._mesh.coordinates(), self.value_size, axis=0)
        if self.family == 'DG':
            return self._mesh.coordinates()[self._mesh.cells()].mean(axis=1)
        return np.zeros((self.value_size, self._mesh.topology().dim()))


class Coefficient:
    """A coefficient of the form: a function on a ``FunctionSpace`` or a constant."""

    def __init__(self, space: FunctionSpace, constant: bool = False, default=0.0):
        self._space = space
        self.is_constant = constant
        self._vec = np.zeros(space.dim())
        self._vec[:] = default

    def function_space(self) -> FunctionSpace:
        return self._space

    def vector(self) -> np.ndarray:
        return self._vec

    def values(self) -> np.ndarray:
        return self._vec

    def assign(self, value):
        self._vec[:] = np.ravel(value)

    @property
    def ufl_shape(self):
        return (self._space.value_size,) if self._space.value_size > 1 else ()


class Form:
    """Coefficient mapping + the set of predefined terms (``equations/form.py:266-355``)."""

    def __init__(self, coefficients: dict, terms: dict):
        self._coefficients = coefficients
        self.terms = terms

    @property
    def coefficients(self):
        return self._coefficients

    def __iter__(self):
        return iter(self._coefficients)

    def keys(self):
        return self._coefficients.keys()

    def values(self):
        return self._coefficients.values()

    def items(self):
        return self._coefficients.items()

    def __getitem__(self, key: str) -> Coefficient:
        return self._coefficients[key]

    def __contains__(self, key: str) -> bool:
        return key in self._coefficients


class BaseResidual:
    pass


DirichletBCTuple = tuple  # (BC value, mesh element type, subdomain name)


class DirichletBC:
    """Homogeneous Dirichlet condition on the closure of a marked subdomain."""

    def __init__(self, space: FunctionSpace, value, mesh_function: MeshFunction, marker: int):
        value = np.ravel(np.asarray(getattr(value, 'values', lambda: value)(), dtype=float))
        if np.any(value != 0):
            raise NotImplementedError(
                "only homogeneous Dirichlet conditions are supported on the device path "
                "(the reference applies them to Newton residuals/Jacobians, "
                "models/transient.py:379-380, 398-399)")
        mesh = space.mesh()
        d = mesh.topology().dim()
        dim = mesh_function.dim()
        ents = mesh_function.where_equal(marker)
        if dim == d - 1:
            verts = np.unique(mesh.facets[ents])
        elif dim == 0:
            verts = ents
        elif dim == d:
            verts = np.unique(mesh.cells()[ents])
        elif dim == 1:
            verts = np.unique(mesh.edges[ents])
        else:
            raise ValueError(f"unsupported Dirichlet entity dimension {dim}")
        n = space.value_size
        self.dofs = (n * verts[:, None] + np.arange(n)[None, :]).reshape(-1).astype(np.int64)

    def get_boundary_values(self) -> dict:
        return {int(dof): 0.0 for dof in self.dofs}

    def apply(self, tensor):
        """``bc.apply(b)`` on a host vector: b[dofs] = 0."""
        tensor[self.dofs] = 0.0


class FenicsResidual(BaseResidual):
    """
    Representation of the (non-linear) solid residual: form + mesh + Dirichlet conditions
    (``residuals/base.py:23-113``).  Kept under the reference's class name.
    """

    def __init__(
        self,
        form: Form,
        mesh: Mesh,
        mesh_functions: list,
        mesh_subdomains: list,
        dirichlet_bc_specs: Optional[dict] = None,
    ):
        self._mesh = mesh
        self._ref_mesh_coords = np.array(mesh.coordinates())
        self._form = form
        self._mesh_functions = mesh_functions
        self._mesh_subdomains = mesh_subdomains

        zero_value = np.zeros(mesh.topology().dim())
        if dirichlet_bc_specs is None:
            dirichlet_bc_specs = {'state/u1': [(zero_value, 'facet', 'fixed')]}
        self._dirichlet_bc_specs = dirichlet_bc_specs
        self._dirichlet_bcs = {
            coeff_key: tuple(
                DirichletBC(
                    form[coeff_key].function_space(), value,
                    self.mesh_function(element_type),
                    self.mesh_subdomain(element_type)[subdomain],
                )
                for (value, element_type, subdomain) in bc_tuples
            )
            for coeff_key, bc_tuples in dirichlet_bc_specs.items()
        }

    @property
    def form(self) -> Form:
        return self._form

    def mesh(self) -> Mesh:
        return self._mesh

    @property
    def ref_mesh_coords(self) -> np.ndarray:
        return self._ref_mesh_coords

    def mesh_function(self, mesh_element_type: Union[str, int]) -> MeshFunction:
        return self._mesh_functions[mesh_element_type_dim(mesh_element_type)]

    def mesh_subdomain(self, mesh_element_type: Union[str, int]) -> dict:
        return self._mesh_subdomains[mesh_element_type_dim(mesh_element_type)]

    @property
    def dirichlet_bcs(self):
        return self._dirichlet_bcs

    # --- setup data for the device path -----------------------------------------
    def fixed_dofs(self) -> np.ndarray:
        bcs = self._dirichlet_bcs.get('state/u1', ())
        if not bcs:
            return np.zeros(0, dtype=np.int64)
        return np.unique(np.concatenate([bc.dofs for bc in bcs]))

    def pressure_facets(self):
        """Exterior facets of the 'pressure' subdomain: ``ds(pressure)`` (solid.py:179-180).

        Returns (facet ids, parent cell, local vertex of the parent cell opposite the facet).
        """
        mesh = self._mesh
        marker = self.mesh_subdomain('facet')['pressure']
        tags = self.mesh_function('facet').array()
        ext = mesh.exterior_facets
        fids = ext[tags[ext] == marker]
        return fids, mesh.facet_cells[fids, 0], mesh.facet_opposite[fids, 0]


ResArgs = tuple
ResReturn = dict


class JaxResidual(BaseResidual):
    """Residual given as a callable ``res(state, control, prop)`` with prototype arguments
    (``residuals/base.py:115-132``).  Kept under the reference's class name; no JAX."""

    def __init__(self, res: Callable, res_args: ResArgs):
        self._res = res
        self._res_args = res_args

    @property
    def res(self):
        return self._res

    @property
    def res_args(self):
        return self._res_args
