"""The arithmetic the CUDA kernels execute (csrc/elem.cuh + node_assembly.cuh), run in a CPU
loop by the test-only harness tests/hostcheck, against the oracle: <= 1e-12 relative, and the
CSR pattern built by the product's setup bit-exact against the oracle's independent builder."""

import ctypes
import os
import subprocess

import numpy as np
import pytest

from helpers import mesh_tuples, oracle_problem, random_solid_prop, random_state, rel_row_err
from femvf_b200 import tables
from femvf_b200.residuals import solid as slr
from oracle import model as om

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def lib():
    hc = os.path.join(HERE, 'hostcheck')
    so = os.path.join(hc, 'libhostcheck.so')
    srcs = [os.path.join(hc, 'hostcheck.cpp'),
            os.path.join(HERE, '..', 'vf-fem_b200', 'csrc', 'elem.cuh'),
            os.path.join(HERE, '..', 'vf-fem_b200', 'csrc', 'node_assembly.cuh'),
            os.path.join(HERE, '..', 'vf-fem_b200', 'csrc', 'fan_assembly.cuh'),
            os.path.join(HERE, '..', 'vf-fem_b200', 'csrc', 'p2_node.cuh'),
            os.path.join(HERE, '..', 'vf-fem_b200', 'csrc', 'tet_tables.h'),
            os.path.join(HERE, '..', 'vf-fem_b200', 'csrc', 'p2_tables.h')]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(['g++', '-O2', '-shared', '-fPIC', '-o', so, srcs[0]], check=True)
    return ctypes.CDLL(so)


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize('mesh_name', ['square5', 'cube332', 'm5'])
@pytest.mark.parametrize('contact,membrane,damping', [(0, 0, 0), (1, 0, 0), (1, 1, 0), (1, 0, 1)])
def test_device_element_math_on_cpu(lib, mesh_name, contact, membrane, damping):
    rng = np.random.default_rng(7)
    Residual = slr.KelvinVoigtWEpithelium if membrane else slr.KelvinVoigt
    res = Residual(*mesh_tuples()[mesh_name]())
    mesh = res.mesh()
    d = mesh.topology().dim()
    fids, pfc, pfo = res.pressure_facets()
    T = tables.build_tables(mesh.coordinates(), mesh.cells(), pfc, pfo, res.fixed_dofs())
    prob = oracle_problem(res)
    assert np.array_equal(prob.rowptr, T['rowptr']) and np.array_equal(prob.colidx, T['colidx'])
    N, ne, nn = prob.N, prob.ne, prob.nn
    prop = random_solid_prop(prob, rng, membrane=True)
    if damping:
        prop.update(rayleigh_m=rng.uniform(5, 20), rayleigh_k=rng.uniform(1e-5, 1e-4))
    so = om.SolidOracle(prob, contact=bool(contact), membrane=bool(membrane))
    u1, u0, v0, a0 = random_state(N, rng)
    p1 = rng.uniform(0, 8e3, nn)
    dt = 1e-4
    Jo = so.jac(u1, dt, prop, p1)
    Fo = so.res(u1, (u0, v0, a0), dt, prop, p1)
    scal = np.zeros(10)
    scal[7], scal[8] = prop.get('rayleigh_m', 0.0), prop.get('rayleigh_k', 0.0)
    scal[0], scal[1], scal[2] = 0.45, prop['ycontact'], prop['kcontact']
    scal[3:3 + d] = prop['ncontact']
    J = np.zeros(len(T['colidx']))
    F = np.zeros(N)
    rc = lib.hostcheck_assemble(
        d, nn, ne, T['nfp'], P(T['xyz']), P(T['cells']), P(T['brptr']), P(T['bcol']),
        P(T['n2e_ptr']), P(T['n2e']), P(T['n2f_ptr']), P(T['n2f']), P(T['pf_cell']),
        P(T['pf_opp']), P(T['bc']), P(prop['rho']), P(prop['eta']), P(prop['emod']), P(scal),
        P(prop['emod_membrane']), P(prop['nu_membrane']), P(prop['th_membrane']),
        contact, membrane, damping, P(u1), P(u0), P(v0), P(a0), P(p1), ctypes.c_double(dt), P(J),
        P(F))
    assert rc == 0
    assert rel_row_err(J, Jo) <= 1e-12
    assert np.max(np.abs(F - Fo)) <= 1e-12 * np.max(np.abs(Fo))
    if d == 3:
        # table-driven form (csrc/tet_tables.h + assemble_node_tet): same cells, same order, same
        # arithmetic -> the same bits
        J2 = np.full(len(T['colidx']), np.nan)
        F2 = np.full(N, np.nan)
        rc = lib.hostcheck_assemble_tet_tables(
            nn, ne, T['nfp'], P(T['xyz']), P(T['cells']), P(T['brptr']), P(T['bcol']),
            P(T['n2e_ptr']), P(T['n2e']), P(T['n2f_ptr']), P(T['n2f']), P(T['pf_cell']),
            P(T['pf_opp']), P(T['bc']), P(prop['rho']), P(prop['eta']), P(prop['emod']), P(scal),
            P(prop['emod_membrane']), P(prop['nu_membrane']), P(prop['th_membrane']),
            contact, membrane, damping, P(u1), P(u0), P(v0), P(a0), P(p1), ctypes.c_double(dt),
            P(J2), P(F2))
        assert rc == 0
        assert np.array_equal(J2, J) and np.array_equal(F2, F)


@pytest.mark.parametrize('mesh_name,nodes_per_tile', [('square5', 7), ('m5', 16), ('m5', 96)])
@pytest.mark.parametrize('contact,membrane,damping', [(0, 0, 0), (1, 1, 0), (0, 0, 1)])
def test_two_phase_tile_algorithm_on_cpu(lib, mesh_name, nodes_per_tile, contact, membrane,
                                         damping):
    """CPU emulation of asm_tile2_kernel + facet_bc_kernel with the product's tile tables."""
    rng = np.random.default_rng(11)
    Residual = slr.KelvinVoigtWEpithelium if membrane else slr.KelvinVoigt
    res = Residual(*mesh_tuples()[mesh_name]())
    mesh = res.mesh()
    fids, pfc, pfo = res.pressure_facets()
    T = tables.build_tables(mesh.coordinates(), mesh.cells(), pfc, pfo, res.fixed_dofs())
    ts = tables.tile_partition(T['brptr'], 2, nodes_per_tile, 64 * nodes_per_tile)
    TT = tables.build_tile_elem_tables(T, ts)
    assert TT is not None
    # every (tile, cell) incidence is listed once; pair slots point at the right columns
    assert len(np.unique(TT['te_elem'][TT['te_ptr'][0]:TT['te_ptr'][1]])) == TT['te_ptr'][1]
    prob = oracle_problem(res)
    N, ne, nn = prob.N, prob.ne, prob.nn
    prop = random_solid_prop(prob, rng, membrane=True)
    if damping:
        prop.update(rayleigh_m=rng.uniform(5, 20), rayleigh_k=rng.uniform(1e-5, 1e-4))
    so = om.SolidOracle(prob, contact=bool(contact), membrane=bool(membrane))
    u1, u0, v0, a0 = random_state(N, rng)
    p1 = rng.uniform(0, 8e3, nn)
    dt = 1e-4
    Jo = so.jac(u1, dt, prop, p1)
    Fo = so.res(u1, (u0, v0, a0), dt, prop, p1)
    scal = np.zeros(10)
    scal[7], scal[8] = prop.get('rayleigh_m', 0.0), prop.get('rayleigh_k', 0.0)
    scal[0], scal[1], scal[2] = 0.45, prop['ycontact'], prop['kcontact']
    scal[3:5] = prop['ncontact']
    J = np.full(len(T['colidx']), np.nan)
    F = np.full(N, np.nan)
    rc = lib.hostcheck_assemble_tile2(
        nn, ne, T['nfp'], P(T['xyz']), P(T['cells']), P(T['brptr']), P(T['bcol']),
        P(T['n2e_ptr']), P(T['n2e']), P(T['n2f_ptr']), P(T['n2f']), P(T['pf_cell']),
        P(T['pf_opp']), P(T['bc']), P(prop['rho']), P(prop['eta']), P(prop['emod']), P(scal),
        P(prop['emod_membrane']), P(prop['nu_membrane']), P(prop['th_membrane']),
        contact, membrane, damping, P(u1), P(u0), P(v0), P(a0), P(p1), ctypes.c_double(dt),
        len(ts) - 1, P(ts), P(TT['te_ptr']), P(TT['te_elem']), P(TT['pair_info']),
        TT['max_tile_elems'], P(J), P(F))
    assert rc == 0
    assert rel_row_err(J, Jo) <= 1e-12
    assert np.max(np.abs(F - Fo)) <= 1e-12 * np.max(np.abs(Fo))


@pytest.mark.parametrize('mesh_name,tile_nodes', [('square5', 64), ('m5', 64), ('m5', 128)])
@pytest.mark.parametrize('contact,membrane,damping,is_static',
                         [(0, 0, 0, 0), (1, 1, 0, 0), (0, 0, 1, 0), (1, 0, 0, 1)])
def test_fan_walk_algorithm_on_cpu(lib, mesh_name, tile_nodes, contact, membrane, damping,
                                   is_static):
    """CPU emulation of asm_fan_kernel + facet_bc_kernel with the product's fan tables."""
    rng = np.random.default_rng(13)
    Residual = slr.KelvinVoigtWEpithelium if membrane else slr.KelvinVoigt
    res = Residual(*mesh_tuples()[mesh_name]())
    mesh = res.mesh()
    fids, pfc, pfo = res.pressure_facets()
    T = tables.build_tables(mesh.coordinates(), mesh.cells(), pfc, pfo, res.fixed_dofs())
    assert T['fan_ok']
    FT = tables.build_fan_tables(T, tile_nodes)
    assert FT is not None
    prob = oracle_problem(res)
    N, ne, nn = prob.N, prob.ne, prob.nn
    prop = random_solid_prop(prob, rng, membrane=True)
    if damping:
        prop.update(rayleigh_m=rng.uniform(5, 20), rayleigh_k=rng.uniform(1e-5, 1e-4))
    so = om.SolidOracle(prob, contact=bool(contact), membrane=bool(membrane))
    u1, u0, v0, a0 = random_state(N, rng)
    p1 = rng.uniform(0, 8e3, nn)
    dt = 1e-4
    if is_static:   # oracle.model.static_solid_configuration: no inertia / damping
        sp = dict(prop, rho=np.zeros(ne), eta=np.zeros(ne))
        Jo = so.jac(u1, 1.0, sp, p1)
        Fo = so.res(u1, (u1, np.zeros(N), np.zeros(N)), 1.0, sp, p1)
    else:
        Jo = so.jac(u1, dt, prop, p1)
        Fo = so.res(u1, (u0, v0, a0), dt, prop, p1)
    scal = np.zeros(10)
    scal[7], scal[8] = prop.get('rayleigh_m', 0.0), prop.get('rayleigh_k', 0.0)
    scal[0], scal[1], scal[2] = 0.45, prop['ycontact'], prop['kcontact']
    scal[3:5] = prop['ncontact']
    J = np.full(len(T['colidx']), np.nan)
    F = np.full(N, np.nan)
    rc = lib.hostcheck_assemble_fan(
        nn, ne, T['nfp'], P(T['xyz']), P(T['cells']), P(T['brptr']), P(T['bcol']),
        P(T['n2e_ptr']), P(T['n2e']), P(T['n2f_ptr']), P(T['n2f']), P(T['pf_cell']),
        P(T['pf_opp']), P(T['bc']), P(prop['rho']), P(prop['eta']), P(prop['emod']), P(scal),
        P(prop['emod_membrane']), P(prop['nu_membrane']), P(prop['th_membrane']),
        contact, membrane, damping, P(u1), P(u0), P(v0), P(a0), P(p1), ctypes.c_double(dt),
        is_static, FT['tile_nodes'], FT['ntiles'], P(FT['desc']), P(FT['ring']), P(FT['halo']),
        P(FT['tcell']), FT['max_verts'], FT['max_rows'], FT['max_cells'], FT['max_blocks'],
        P(J), P(F))
    assert rc == 0
    assert not np.any(np.isnan(J)) and not np.any(np.isnan(F))
    assert rel_row_err(J, Jo) <= 1e-12
    assert np.max(np.abs(F - Fo)) <= 1e-12 * np.max(np.abs(Fo))


@pytest.mark.parametrize('mesh_name,interleave', [('square5', False), ('square5', True),
                                                  ('m5', True)])
@pytest.mark.parametrize('flags', [3, 2, 1, 7, 5])   # +4: packed nodal state of the pre-pass
def test_p2_second_kernel_arithmetic_on_cpu(lib, mesh_name, interleave, flags):
    """vf::p2_node_row (the per-node code of p2_assemble_warp_kernel: structural zeros of the
    reference tensor skipped, rows in CSR layout) in a CPU loop with the product's P2 tables,
    against the quadrature oracle."""
    from femvf_b200.p2 import build_p2_tables
    from oracle import fem_p2
    import scipy.sparse as sp
    res = slr.KelvinVoigt(*mesh_tuples()[mesh_name]())
    mesh = res.mesh()
    coords, cells = mesh.coordinates(), mesh.cells()
    p1prob = oracle_problem(res)
    # Dirichlet edges: boundary edges between two fixed vertices of the P1 problem
    fv = np.unique(p1prob.fixed_dofs // 2)
    e = np.sort(np.concatenate([cells[:, [1, 2]], cells[:, [0, 2]], cells[:, [0, 1]]]), axis=1)
    key, cnt = np.unique(e[:, 0] * len(coords) + e[:, 1], return_counts=True)
    be = np.stack([key // len(coords), key % len(coords)], axis=1)[cnt == 1]
    fe = be[np.isin(be[:, 0], fv) & np.isin(be[:, 1], fv)]
    T = build_p2_tables(coords, cells, p1prob.pfacets, p1prob.pfacet_cells, fe, interleave)
    keep = T['keep']
    prob0 = fem_p2.SolidProblemP2(coords, cells, p1prob.pfacets, p1prob.pfacet_cells, [])
    fixed_old = prob0.closure_nodes(fe)
    prob = fem_p2.SolidProblemP2(coords, cells, p1prob.pfacets, p1prob.pfacet_cells, fixed_old)
    new_of_old = np.concatenate([T['vertex_ids'], T['edge_node']])
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
    dof_new_of_old = (2 * new_of_old[:, None] + np.arange(2)[None, :]).ravel()
    old_of_new = np.argsort(dof_new_of_old)
    F_ref_n = F_ref[old_of_new]
    J_ref_n = J_ref[old_of_new][:, old_of_new].tocsr()
    J_ref_n.sort_indices()

    def nodal(v, width):
        out = np.empty_like(v)
        out.reshape(-1, width)[new_of_old] = v.reshape(-1, width)
        return out
    indptr, indices = tables.scalar_csr_from_graph(T['brptr'].astype(np.int32),
                                                   T['bcol'].astype(np.int32), 2)
    assert np.array_equal(indptr, J_ref_n.indptr) and np.array_equal(indices, J_ref_n.indices)
    J = np.full(len(indices), np.nan)
    F = np.full(N, np.nan)
    U1, U0, V0, A0, P1 = nodal(u1, 2), nodal(u0, 2), nodal(v0, 2), nodal(a0, 2), nodal(p1, 1)
    rc = lib.hostcheck_p2_assemble(
        nn, P(keep[0]), P(keep[1]), P(keep[2]), P(keep[3]), P(keep[4]), P(keep[5]), P(keep[6]),
        P(keep[7]), P(keep[8]), P(keep[9]), P(keep[10]), P(keep[11]), P(keep[12]), P(keep[13]),
        P(keep[14]), T['nv'], P(prop['emod']), P(prop['eta']), P(prop['rho']),
        ctypes.c_double(0.45), P(U1), P(U0), P(V0), P(A0), P(P1), ctypes.c_double(dt), flags,
        P(J), P(F))
    assert rc == 0
    if flags & 2:
        assert rel_row_err(J, J_ref_n) <= 1e-12
    else:
        assert np.all(np.isnan(J))
    if flags & 1:
        assert np.max(np.abs(F - F_ref_n)) <= 1e-12 * np.max(np.abs(F_ref_n))
    else:
        assert np.all(np.isnan(F))


def test_shape_parameter_tables_on_cpu(lib):
    """KelvinVoigtWShape after ``set_prop`` with a non-zero 'umesh': the rebuilt device tables
    (``FenicsModel.assembly_tables``) run through the kernels' element code give the oracle's
    residual and Jacobian on the displaced mesh (the CPU half of
    tests/test_gpu_assembly.py::test_shape_parameter_moves_the_device_mesh)."""
    from femvf_b200.models import transient
    model = transient.FenicsModel(slr.KelvinVoigtWShape(*mesh_tuples()['m5']()))
    mesh = model.residual.mesh()
    ref = mesh.coordinates().copy()
    rng = np.random.default_rng(5)
    prop_bv = model.prop.copy()
    du = 2e-3 * rng.standard_normal(ref.shape)
    prop_bv['umesh'][:] = du.ravel()
    model.set_prop(prop_bv)
    T = model.assembly_tables
    assert np.array_equal(T['xyz'], (ref + du).T)
    prob = oracle_problem(model.residual)
    assert np.array_equal(prob.coords, ref + du)
    N, ne, nn = prob.N, prob.ne, prob.nn
    prop = random_solid_prop(prob, rng, membrane=True)
    so = om.SolidOracle(prob, contact=False, membrane=False)
    u1, u0, v0, a0 = random_state(N, rng)
    p1 = rng.uniform(0, 8e3, nn)
    dt = 1e-4
    Jo = so.jac(u1, dt, prop, p1)
    Fo = so.res(u1, (u0, v0, a0), dt, prop, p1)
    scal = np.zeros(10)
    scal[0], scal[1], scal[2] = 0.45, prop['ycontact'], prop['kcontact']
    scal[3:5] = prop['ncontact']
    J = np.zeros(len(T['colidx']))
    F = np.zeros(N)
    rc = lib.hostcheck_assemble(
        2, nn, ne, T['nfp'], P(T['xyz']), P(T['cells']), P(T['brptr']), P(T['bcol']),
        P(T['n2e_ptr']), P(T['n2e']), P(T['n2f_ptr']), P(T['n2f']), P(T['pf_cell']),
        P(T['pf_opp']), P(T['bc']), P(prop['rho']), P(prop['eta']), P(prop['emod']), P(scal),
        P(prop['emod_membrane']), P(prop['nu_membrane']), P(prop['th_membrane']),
        0, 0, 0, P(u1), P(u0), P(v0), P(a0), P(p1), ctypes.c_double(dt), P(J), P(F))
    assert rc == 0
    assert rel_row_err(J, Jo) <= 1e-12
    assert np.max(np.abs(F - Fo)) <= 1e-12 * np.max(np.abs(Fo))
    # moving back restores the reference mesh and its tables
    model.set_prop(_with(model.prop, 'umesh', 0.0))
    assert np.array_equal(mesh.coordinates(), ref)
    assert np.array_equal(model.assembly_tables['xyz'], ref.T)


def _with(prop, key, value):
    out = prop.copy()
    out[key][:] = value
    return out
