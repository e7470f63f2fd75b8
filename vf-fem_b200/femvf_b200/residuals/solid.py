"""
Predefined solid residuals: mirror of ``/root/reference/src/femvf/residuals/solid.py``.

Each class lists the predefined forms it sums (``solid.py:144-240``); the coefficient
dictionary is built in the same insertion order as ``equations/form.py:358-442``
(``add_form`` merges coefficient dicts left to right), because that order defines the
label order of the model's ``prop`` BlockVector (``models/transient.py:187-218``).
"""

from __future__ import annotations

from typing import Optional

import numpy as np

from .base import FenicsResidual, Form, Coefficient, FunctionSpace, DirichletBCTuple
from ..mesh import Mesh

# Coefficient specs of the predefined forms, in the reference's order
# (equations/form.py:521-524, 545-550, 738-741, 764-770, 805-810, 923-930, 970-973).
_CG1V, _CG1S, _DG0, _CONST_S, _CONST_V = 'cg1v', 'cg1s', 'dg0', 'const_s', 'const_v'

FORM_SPECS = {
    'InertialForm': [('state/a1', _CG1V, 0.0), ('prop/rho', _DG0, 0.0)],
    'IsotropicElasticForm': [
        ('state/u1', _CG1V, 0.0), ('state/v1', _CG1V, 0.0),
        ('prop/emod', _DG0, 0.0), ('prop/nu', _CONST_S, 0.45)],
    'KelvinVoigtForm': [('state/v1', _CG1V, 0.0), ('prop/eta', _DG0, 0.0)],
    'SurfacePressureForm': [('state/u1', _CG1V, 0.0), ('control/p1', _CG1S, 0.0)],
    'ManualSurfaceContactTractionForm': [
        ('state/u1', _CG1V, 0.0), ('control/tcontact', _CG1V, 0.0),
        ('prop/ycontact', _CONST_S, np.inf), ('prop/ncontact', _CONST_V, 'e_y'),
        ('prop/kcontact', _CONST_S, 1.0)],
    'RayleighDampingForm': [
        ('state/v1', _CG1V, 0.0), ('prop/rho', _DG0, 0.0), ('prop/emod', _DG0, 0.0),
        ('prop/nu', _CONST_S, 0.45), ('prop/rayleigh_m', _CONST_S, 1.0),
        ('prop/rayleigh_k', _CONST_S, 1.0)],
    # form.py:1037-1062: contributes 0 * (x . w) dx; its coefficient moves the mesh (set_prop)
    'ShapeForm': [('prop/umesh', _CG1V, 0.0)],
    'IsotropicMembraneForm': [
        ('state/u1', _CG1V, 0.0), ('prop/emod_membrane', _DG0, 0.0),
        ('prop/nu_membrane', _DG0, 0.45), ('prop/th_membrane', _DG0, 0.0)],
}


def _make_coefficient(mesh: Mesh, kind: str, default) -> Coefficient:
    d = mesh.topology().dim()
    if kind == _CG1V:
        return Coefficient(FunctionSpace(mesh, 'CG', 1, d), default=default)
    if kind == _CG1S:
        return Coefficient(FunctionSpace(mesh, 'CG', 1, 1), default=default)
    if kind == _DG0:
        return Coefficient(FunctionSpace(mesh, 'DG', 0, 1), default=default)
    if kind == _CONST_S:
        return Coefficient(FunctionSpace(mesh, 'R', 0, 1), constant=True, default=default)
    if kind == _CONST_V:
        c = Coefficient(FunctionSpace(mesh, 'R', 0, d), constant=True, default=0.0)
        if isinstance(default, str) and default == 'e_y':
            c.vector()[1] = 1.0  # form.py:789-791
        return c
    raise ValueError(kind)


def build_form(mesh: Mesh, form_names: list, terms: dict) -> Form:
    coefficients = {}
    for name in form_names:
        for key, kind, default in FORM_SPECS[name]:
            if key not in coefficients:
                coefficients[key] = _make_coefficient(mesh, kind, default)
    return Form(coefficients, terms)


class PredefinedSolidResidual(FenicsResidual):
    """Class representing a pre-defined residual (``solid.py:108-142``)."""

    FORM_NAMES: list = []
    TERMS: dict = {}

    def __init__(
        self,
        mesh: Mesh,
        mesh_functions: list,
        mesh_subdomains: list,
        dirichlet_bcs: Optional[dict] = None,
    ):
        form = self.init_form(mesh, mesh_functions, mesh_subdomains)
        super().__init__(form, mesh, mesh_functions, mesh_subdomains,
                         dirichlet_bc_specs=dirichlet_bcs)

    def init_form(self, mesh, mesh_functions, mesh_subdomains) -> Form:
        if not self.FORM_NAMES:
            raise NotImplementedError()
        return build_form(mesh, self.FORM_NAMES, dict(self.TERMS))


class Rayleigh(PredefinedSolidResidual):
    """Inertia + isotropic elasticity + Rayleigh damping - follower pressure - contact
    traction (``solid.py:144-165``)."""

    FORM_NAMES = ['InertialForm', 'IsotropicElasticForm', 'RayleighDampingForm',
                  'SurfacePressureForm', 'ManualSurfaceContactTractionForm']
    TERMS = {'membrane': False, 'damping': 'rayleigh'}


class KelvinVoigt(PredefinedSolidResidual):
    """Inertia + Kelvin-Voigt damping + isotropic elasticity - follower pressure - contact
    traction (``solid.py:168-189``)."""

    FORM_NAMES = ['InertialForm', 'KelvinVoigtForm', 'IsotropicElasticForm',
                  'SurfacePressureForm', 'ManualSurfaceContactTractionForm']
    TERMS = {'membrane': False}


class KelvinVoigtWShape(PredefinedSolidResidual):
    """``KelvinVoigt`` with a shape parameter (``solid.py:192-215``): ``ShapeForm`` adds nothing to
    the residual; its coefficient ``umesh`` (a nodal displacement of the reference mesh, last in
    the property vector) is applied to the mesh coordinates by ``FenicsModel.set_prop``
    (``models/transient.py:347-360``)."""

    FORM_NAMES = ['InertialForm', 'IsotropicElasticForm', 'KelvinVoigtForm',
                  'SurfacePressureForm', 'ManualSurfaceContactTractionForm', 'ShapeForm']
    TERMS = {'membrane': False}


class KelvinVoigtWEpithelium(PredefinedSolidResidual):
    """``KelvinVoigt`` plus an isotropic membrane on the 'pressure' surface
    (``solid.py:218-240``)."""

    FORM_NAMES = ['InertialForm', 'IsotropicMembraneForm', 'IsotropicElasticForm',
                  'KelvinVoigtForm', 'SurfacePressureForm', 'ManualSurfaceContactTractionForm']
    TERMS = {'membrane': True}
