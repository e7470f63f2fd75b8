"""Shared test helpers: build oracle problems from product-side setup objects."""

import numpy as np

from femvf_b200 import mesh as M, meshgen
from femvf_b200.residuals import solid as slr, fluid as flr
from oracle import fem, model as om


def mesh_tuples():
    return {
        'square5': lambda: M.fixture_mesh_tuple(M.unit_square_mesh(5, 5)),
        'cube332': lambda: M.fixture_mesh_tuple(M.unit_cube_mesh(3, 3, 2)),
        'm5': lambda: meshgen.m5_cb_mesh(0.05),
    }


def oracle_problem(residual) -> fem.SolidProblem:
    mesh = residual.mesh()
    fids, pf_cell, _ = residual.pressure_facets()
    return fem.SolidProblem(mesh.coordinates(), mesh.cells(), mesh.facets[fids], pf_cell,
                            residual.fixed_dofs())


def random_solid_prop(prob, rng, membrane=False, contact_offset=-0.01, kcontact=1e8):
    d, ne = prob.d, prob.ne
    prop = dict(
        rho=np.full(ne, 1.0), eta=rng.uniform(1, 5, ne), emod=rng.uniform(2.5e4, 1e5, ne),
        nu=0.45, ncontact=np.eye(d)[1], ycontact=prob.coords[:, 1].max() + contact_offset,
        kcontact=kcontact)
    if membrane:
        prop.update(emod_membrane=rng.uniform(1e4, 5e4, ne), nu_membrane=np.full(ne, 0.45),
                    th_membrane=np.full(ne, 0.005))
    return prop


def set_model_prop(model_prop, prop):
    """Copy an oracle property dict into a model BlockVector (labels that exist)."""
    for key, value in prop.items():
        if key in model_prop:
            model_prop[key][:] = value


def random_state(N, rng):
    return (rng.uniform(-1e-2, 1e-2, N), rng.uniform(-1e-2, 1e-2, N), rng.uniform(-1, 1, N),
            rng.uniform(-1e3, 1e3, N))


def rel_row_err(vals, J_ref):
    """max |a - b| / (max |row|) over CSR entries."""
    rowmax = np.maximum.reduceat(np.abs(J_ref.data), J_ref.indptr[:-1])
    scale = np.repeat(rowmax, np.diff(J_ref.indptr))
    return float(np.max(np.abs(vals - J_ref.data) / scale))
