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
            os.path.join(HERE, '..', 'vf-fem_b200', 'csrc', 'fan_assembly.cuh')]
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
