"""GPU parity: P2 triangle residual + Jacobian (csrc/p2.cu through vf_p2_*) vs the quadrature
oracle (oracle/fem_p2.py): entries <= 1e-12 relative, CSR pattern bit-exact, bit-reproducible."""

import numpy as np
import pytest
import scipy.sparse as sp

from helpers import mesh_tuples, oracle_problem, rel_row_err
from oracle import fem_p2

pytestmark = pytest.mark.gpu


def _setup(mesh_name, levels=0):
    from femvf_b200 import meshgen
    from femvf_b200.residuals import solid as slr
    mt = meshgen.m5_cb_refined(0.05, levels) if mesh_name == 'm5r' else mesh_tuples()[mesh_name]()
    res = slr.KelvinVoigt(*mt)
    p1 = oracle_problem(res)
    mesh = res.mesh()
    fids = res.facet_ids('fixed') if hasattr(res, 'facet_ids') else None
    return mesh.coordinates(), mesh.cells(), p1, res


def _fixed_edges(p1prob, coords, cells):
    """Boundary edges whose two vertices are Dirichlet vertices of the P1 problem."""
    fv = np.unique(p1prob.fixed_dofs // 2)
    e = np.sort(np.concatenate([cells[:, [1, 2]], cells[:, [0, 2]], cells[:, [0, 1]]]), axis=1)
    key, count = np.unique(e[:, 0] * len(coords) + e[:, 1], return_counts=True)
    be = np.stack([key // len(coords), key % len(coords)], axis=1)[count == 1]
    return be[np.isin(be[:, 0], fv) & np.isin(be[:, 1], fv)]


@pytest.mark.parametrize('mesh_name,levels,interleave', [('square5', 0, False), ('square5', 0, True),
                                                         ('m5', 0, True), ('m5r', 2, True)])
def test_p2_assembly_parity(mesh_name, levels, interleave):
    import torch
    from femvf_b200.p2 import P2Assembler
    coords, cells, p1prob, res = _setup(mesh_name, levels)
    fe = _fixed_edges(p1prob, coords, cells)
    asm = P2Assembler(coords, cells, p1prob.pfacets, p1prob.pfacet_cells, fe,
                      interleave=interleave)
    # oracle numbering: vertices first, mid-edge nodes appended in edge order
    prob0 = fem_p2.SolidProblemP2(coords, cells, p1prob.pfacets, p1prob.pfacet_cells, [])
    fixed_old = prob0.closure_nodes(fe)
    prob = fem_p2.SolidProblemP2(coords, cells, p1prob.pfacets, p1prob.pfacet_cells, fixed_old)
    new_of_old = np.concatenate([asm.vertex_ids, asm.edge_node])
    assert np.array_equal(np.sort(new_of_old[fixed_old]), asm.fixed_nodes)
    rng = np.random.default_rng(7)
    N, nn, ne = prob.N, prob.nn, prob.ne
    prop = dict(rho=rng.uniform(0.9, 1.1, ne), eta=rng.uniform(1, 5, ne),
                emod=rng.uniform(2.5e4, 1e5, ne), nu=0.45)
    u1, u0 = rng.uniform(-1e-2, 1e-2, N), rng.uniform(-1e-2, 1e-2, N)
    v0, a0 = rng.uniform(-1, 1, N), rng.uniform(-1e3, 1e3, N)
    p1 = rng.uniform(0, 8e3, nn)
    dt = 1e-4
    F_ref = fem_p2.assemble_res_u(prob, u1, u0, v0, a0, dt, prop, p1)
    J_ref = fem_p2.assemble_jac_uu(prob, u1, dt, prop, p1)
    # permute the oracle to the assembler's numbering
    dof_new_of_old = (2 * new_of_old[:, None] + np.arange(2)[None, :]).ravel()
    old_of_new = np.argsort(dof_new_of_old)
    F_ref_n = F_ref[old_of_new]
    J_ref_n = J_ref[old_of_new][:, old_of_new].tocsr()
    J_ref_n.sort_indices()

    def dev_nodal(v, width):
        out = np.empty_like(v)
        out.reshape(-1, width)[new_of_old] = v.reshape(-1, width)
        return torch.as_tensor(out, device='cuda')
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device='cuda')
    F, J = asm.assemble(dev_nodal(u1, 2), dev_nodal(u0, 2), dev_nodal(v0, 2), dev_nodal(a0, 2),
                        dev_nodal(p1, 1), t(prop['emod']), t(prop['eta']), t(prop['rho']), 0.45, dt)
    F, vals = F.cpu().numpy(), J.cpu().numpy()
    indptr, indices = asm.csr_pattern()
    # the permuted oracle pattern keeps explicit zeros only where its own pattern had them:
    # compare as matrices on the assembler's pattern
    assert np.max(np.abs(F - F_ref_n)) <= 1e-12 * np.max(np.abs(F_ref_n))
    Jg = sp.csr_matrix((vals, indices, indptr), shape=(N, N))
    assert np.array_equal(indptr, J_ref_n.indptr) and np.array_equal(indices, J_ref_n.indices)
    assert rel_row_err(vals, J_ref_n) <= 1e-12
    # bit-reproducible
    F2, J2 = asm.assemble(dev_nodal(u1, 2), dev_nodal(u0, 2), dev_nodal(v0, 2), dev_nodal(a0, 2),
                          dev_nodal(p1, 1), t(prop['emod']), t(prop['eta']), t(prop['rho']), 0.45, dt)
    assert np.array_equal(J2.cpu().numpy(), vals) and np.array_equal(F2.cpu().numpy(), F)
    del Jg
