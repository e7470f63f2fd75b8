"""GPU parity: device assembly / SpMV / linear solve (through the C ABI) vs the oracle."""

import numpy as np
import pytest
import scipy.sparse.linalg as spla

from helpers import mesh_tuples, oracle_problem, random_solid_prop, set_model_prop, \
    random_state, rel_row_err
from oracle import model as om

pytestmark = pytest.mark.gpu

TOL = 1e-12  # BASELINE.json north_star: assembled entries within 1e-12 relative


@pytest.fixture(scope='module')
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    return torch


@pytest.mark.parametrize('mesh_name', ['square5', 'cube332', 'm5'])
@pytest.mark.parametrize('variant', ['kv', 'kv_contact', 'epithelium_contact', 'rayleigh'])
def test_assembly_parity(torch_cuda, mesh_name, variant):
    from femvf_b200.models import transient
    from femvf_b200.residuals import solid as slr
    rng = np.random.default_rng(42)
    mt = mesh_tuples()[mesh_name]()
    membrane = variant.startswith('epithelium')
    contact = variant.endswith('contact')
    Residual = slr.KelvinVoigtWEpithelium if membrane else slr.KelvinVoigt
    if variant == 'rayleigh':
        Residual = slr.Rayleigh
    Model = transient.NodalContactModel if contact else transient.FenicsModel
    model = Model(Residual(*mt))
    prob = oracle_problem(model.residual)
    N = prob.N
    prop = random_solid_prop(prob, rng, membrane=membrane)
    if variant == 'rayleigh':
        del prop['eta']
        prop.update(rayleigh_m=12.5, rayleigh_k=4e-5)
    mprop = model.prop.copy()
    set_model_prop(mprop, prop)
    model.set_prop(mprop)
    u1, u0, v0, a0 = random_state(N, rng)
    p1 = rng.uniform(0, 8e3, prob.nn)
    dt = 1e-4
    model.dt = dt
    s0 = model.state0.copy(); s0['u'][:] = u0; s0['v'][:] = v0; s0['a'][:] = a0
    s1 = model.state1.copy(); s1['u'][:] = u1
    model.set_ini_state(s0); model.set_fin_state(s1)
    ctl = model.control.copy(); ctl['p'][:] = p1
    model.set_control(ctl)

    so = om.SolidOracle(prob, contact=contact, membrane=membrane)
    F_ref = so.res(u1, (u0, v0, a0), dt, prop, p1)
    J_ref = so.jac(u1, dt, prop, p1)

    res = model.assem_res()
    assert np.max(np.abs(res['u'] - F_ref)) <= TOL * np.max(np.abs(F_ref))
    # nodal Newmark residuals (transient.py:374-377), computed on the device
    from femvf_b200.equations import newmark
    v1, a1 = np.asarray(s1['v']), np.asarray(s1['a'])
    rv_ref = v1 - newmark.newmark_v(u1, u0, v0, a0, dt)
    ra_ref = a1 - newmark.newmark_a(u1, u0, v0, a0, dt)
    assert np.max(np.abs(res['v'] - rv_ref)) <= TOL * np.max(np.abs(rv_ref))
    assert np.max(np.abs(res['a'] - ra_ref)) <= TOL * np.max(np.abs(ra_ref))
    dres = model.assem_dres_dstate1()
    J = dres.sub['u', 'state/u1']
    dv = dres.sub['v', 'state/u1']
    assert np.allclose(dv.diagonal(), -newmark.newmark_v_du1(dt), rtol=0, atol=0)
    assert np.allclose(dres.sub['a', 'state/a1'].diagonal(), 1.0, rtol=0, atol=0)
    # CSR sparsity pattern bit-exact
    assert np.array_equal(J.indptr, J_ref.indptr)
    assert np.array_equal(J.indices, J_ref.indices)
    assert rel_row_err(J.data, J_ref) <= TOL

    # run-to-run bit reproducibility (no atomics)
    J2 = model.assem_dres_dstate1().sub['u', 'state/u1']
    assert np.array_equal(J.data, J2.data)

    # SpMV
    torch = torch_cuda
    x = rng.standard_normal(N)
    xt = torch.as_tensor(x, device='cuda'); yt = torch.empty_like(xt)
    model.engine.spmv(xt, yt)
    y_ref = J_ref @ x
    assert np.max(np.abs(yt.cpu().numpy() - y_ref)) <= 1e-13 * np.max(np.abs(y_ref)) * 10

    # linear solve vs LU
    b = rng.standard_normal(N)
    bt = torch.as_tensor(b, device='cuda'); st = torch.empty_like(bt)
    info = model.engine.linear_solve(bt, st)
    x_ref = spla.splu(J_ref.tocsc()).solve(b)
    err = np.linalg.norm(st.cpu().numpy() - x_ref) / np.linalg.norm(x_ref)
    assert err < 1e-9, (err, info)


def _assemble_and_compare(model, rng, contact=False, membrane=False):
    prob = oracle_problem(model.residual)
    N = prob.N
    prop = random_solid_prop(prob, rng, membrane=membrane)
    mprop = model.prop.copy()
    set_model_prop(mprop, prop)
    model.set_prop(mprop)
    u1, u0, v0, a0 = random_state(N, rng)
    p1 = rng.uniform(0, 8e3, prob.nn)
    dt = 1e-4
    model.dt = dt
    s0 = model.state0.copy(); s0['u'][:] = u0; s0['v'][:] = v0; s0['a'][:] = a0
    s1 = model.state1.copy(); s1['u'][:] = u1
    model.set_ini_state(s0); model.set_fin_state(s1)
    ctl = model.control.copy(); ctl['p'][:] = p1
    model.set_control(ctl)
    so = om.SolidOracle(prob, contact=contact, membrane=membrane)
    F_ref = so.res(u1, (u0, v0, a0), dt, prop, p1)
    J_ref = so.jac(u1, dt, prop, p1)
    res = model.assem_res()
    assert np.max(np.abs(res['u'] - F_ref)) <= TOL * np.max(np.abs(F_ref))
    J = model.assem_dres_dstate1().sub['u', 'state/u1']
    assert np.array_equal(J.indptr, J_ref.indptr)
    assert np.array_equal(J.indices, J_ref.indices)
    assert rel_row_err(J.data, J_ref) <= TOL
    return model


@pytest.mark.parametrize('tile_nodes', ['24', '48', '64', '96'])
def test_assembly_parity_many_tiles(torch_cuda, monkeypatch, tile_nodes):
    """Entry-wise parity on a mesh of ~90-350 tiles (Morton-ordered M5_CB refined 3x, 15.7 k
    triangles): exercises the per-tile halo lists, local vertex slots, the L2 prefetch of the
    far tile and every launch-bounds variant of the tile kernel (128/192/256/320 threads)."""
    from femvf_b200 import meshgen
    from femvf_b200.models import transient
    from femvf_b200.residuals import solid as slr
    monkeypatch.setenv('VF_FAN', '0')              # the round-1 tile kernel (fallback path)
    monkeypatch.setenv('VF_TILE_NODES', tile_nodes)
    monkeypatch.setenv('VF_PF_DIST', '7')          # a far tile that exists on a small grid
    mt = meshgen.m5_cb_refined(0.05, 3)
    model = transient.FenicsModel(slr.KelvinVoigt(*mt))
    _assemble_and_compare(model, np.random.default_rng(int(tile_nodes)))
    info = model.engine.tile_info
    assert info['two_phase'] and info['ntiles'] >= 80
    assert info['nodes_per_tile'] == int(tile_nodes)


@pytest.mark.parametrize('grid,groups,pool_kb', [('5', '3', '0'), ('3', '2', '0'), ('7', '3', '104'),
                                                  ('1', '3', '0')])
def test_assembly_parity_pipeline(torch_cuda, monkeypatch, grid, groups, pool_kb):
    """Entry-wise parity of the persistent producer / consumer pipeline (asm_fan_pipe_kernel) when
    every CTA walks many tiles: 62 tiles of 128 nodes (M5_CB refined 3x) on 1-7 CTAs, two and
    three consumer groups, and a small stage pool that wraps around every few tiles."""
    from femvf_b200 import meshgen
    from femvf_b200.models import transient
    from femvf_b200.residuals import solid as slr
    monkeypatch.setenv('VF_PIPE_GRID', grid)
    monkeypatch.setenv('VF_PIPE_GROUPS', groups)
    if pool_kb != '0':
        monkeypatch.setenv('VF_PIPE_POOL_KB', pool_kb)
    mt = meshgen.m5_cb_refined(0.05, 3)
    model = transient.FenicsModel(slr.KelvinVoigt(*mt))
    model = _assemble_and_compare(model, np.random.default_rng(int(grid)))
    assert model.engine.fan_info is not None and model.engine.fan_info['ntiles'] >= 60
    # Jacobian-only and residual-only launches of the same pipeline
    eng = model.engine
    J = eng.view('J').clone(); F = eng.view('F').clone()
    eng.view('J').zero_(); eng.assemble(0, res=False, jac=True, dt=model.dt)
    assert torch_cuda.equal(J, eng.view('J'))
    eng.view('F').zero_(); eng.assemble(0, res=True, jac=False, dt=model.dt)
    assert torch_cuda.equal(F, eng.view('F'))


@pytest.mark.parametrize('fan_nodes', ['64', '96', '128'])
def test_assembly_parity_fan_kernel_per_tile(torch_cuda, monkeypatch, fan_nodes):
    """The non-persistent fan kernel (asm_fan_kernel, one CTA per tile): the fallback when the
    pipeline's shared-memory plan does not fit, and the only path for 64- / 96-node tiles."""
    from femvf_b200 import meshgen
    from femvf_b200.models import transient
    from femvf_b200.residuals import solid as slr
    monkeypatch.setenv('VF_FAN_PIPE', '0')
    monkeypatch.setenv('VF_FAN_NODES', fan_nodes)
    monkeypatch.setenv('VF_PF_DIST', '5')
    mt = meshgen.m5_cb_refined(0.05, 3)
    model = transient.FenicsModel(slr.KelvinVoigt(*mt))
    model = _assemble_and_compare(model, np.random.default_rng(int(fan_nodes)))
    assert model.engine.fan_info['tile_nodes'] == int(fan_nodes)


def test_assembly_parity_degenerate_meshes(torch_cuda):
    """Smallest inputs: two triangles; a mesh with no Dirichlet and no pressure facets."""
    from femvf_b200 import mesh as M
    from femvf_b200.models import transient
    from femvf_b200.residuals import solid as slr
    rng = np.random.default_rng(5)
    mt = M.fixture_mesh_tuple(M.unit_square_mesh(1, 1))
    _assemble_and_compare(transient.FenicsModel(slr.KelvinVoigt(*mt)), rng)
    # strip every facet tag: free-floating body, no traction
    mesh, mfs, labels = M.fixture_mesh_tuple(M.unit_square_mesh(4, 3))
    mfs[1].array()[:] = 7          # neither 'fixed' (1) nor 'pressure' (0)
    model = transient.FenicsModel(slr.KelvinVoigt(mesh, mfs, labels))
    assert len(model.residual.fixed_dofs()) == 0
    assert len(model.residual.pressure_facets()[0]) == 0
    _assemble_and_compare(model, rng)


@pytest.mark.parametrize('mesh_name', ['m5', 'cube332'])
@pytest.mark.parametrize('variant', ['kv', 'rayleigh'])
def test_state0_and_control_sensitivities(torch_cuda, mesh_name, variant):
    """assem_dres_dstate0 / assem_dres_dcontrol (transient.py:408-435).  F_u is affine in
    (u0, v0, a0) and linear in p1, so a finite difference of the ORACLE residual is exact to
    round-off and pins every block without new oracle code."""
    from femvf_b200.equations import newmark
    from femvf_b200.models import transient
    from femvf_b200.residuals import solid as slr
    rng = np.random.default_rng(11)
    mt = mesh_tuples()[mesh_name]()
    Residual = slr.Rayleigh if variant == 'rayleigh' else slr.KelvinVoigt
    model = transient.FenicsModel(Residual(*mt))
    prob = oracle_problem(model.residual)
    N = prob.N
    prop = random_solid_prop(prob, rng)
    if variant == 'rayleigh':
        del prop['eta']
        prop.update(rayleigh_m=12.5, rayleigh_k=4e-5)
    mprop = model.prop.copy()
    set_model_prop(mprop, prop)
    model.set_prop(mprop)
    u1, u0, v0, a0 = random_state(N, rng)
    p1 = rng.uniform(0, 8e3, prob.nn)
    dt = 1e-4
    model.dt = dt
    s0 = model.state0.copy(); s0['u'][:] = u0; s0['v'][:] = v0; s0['a'][:] = a0
    s1 = model.state1.copy(); s1['u'][:] = u1
    model.set_ini_state(s0); model.set_fin_state(s1)
    ctl = model.control.copy(); ctl['p'][:] = p1
    model.set_control(ctl)
    so = om.SolidOracle(prob, contact=False, membrane=False)

    def F(u0_, v0_, a0_, p_):
        # raw residual rows: the reference applies no Dirichlet condition to these blocks, so
        # compare on the free rows and check the fixed rows separately below
        return so.res(u1, (u0_, v0_, a0_), dt, prop, p_)
    free = np.ones(N, dtype=bool)
    free[model.residual.fixed_dofs()] = False
    base = F(u0, v0, a0, p1)
    d0 = model.assem_dres_dstate0()
    scales = {'u': 1e-2, 'v': 1.0, 'a': 1e3}
    for key, args in (('u', 0), ('v', 1), ('a', 2)):
        dx = rng.uniform(-1, 1, N) * scales[key]
        pert = [u0, v0, a0]
        pert[args] = pert[args] + dx
        fd = F(pert[0], pert[1], pert[2], p1) - base
        A = d0.sub['u', 'state/' + key + '0']
        got = A @ dx
        ref_scale = np.max(np.abs(fd))
        assert np.max(np.abs(got[free] - fd[free])) <= 1e-9 * ref_scale, key
        # pattern is the Jacobian's pattern
        J = model.assem_dres_dstate1().sub['u', 'state/u1']
        assert np.array_equal(A.indptr, J.indptr) and np.array_equal(A.indices, J.indices)
    # nodal blocks: F_v = v1 - v_nmk, F_a = a1 - a_nmk
    assert np.all(d0.sub['v', 'state/u0'].diagonal() == -newmark.newmark_v_du0(dt))
    assert np.all(d0.sub['v', 'state/v0'].diagonal() == -newmark.newmark_v_dv0(dt))
    assert np.all(d0.sub['a', 'state/a0'].diagonal() == -newmark.newmark_a_da0(dt))
    # Dirichlet rows are NOT identity rows here (bc.apply is commented out, transient.py:415)
    fixed = model.residual.fixed_dofs()
    if len(fixed):
        A = d0.sub['u', 'state/a0']
        assert np.any(A[fixed].toarray() != 0.0)

    dc = model.assem_dres_dcontrol()
    B = dc.sub['u', 'control/p1'] if 'control/p1' in dc.labels[1] else dc.sub['u', 0]
    dp = rng.uniform(-1e3, 1e3, prob.nn)
    fd = F(u0, v0, a0, p1 + dp) - base
    got = B @ dp
    assert np.max(np.abs(got[free] - fd[free])) <= 1e-9 * np.max(np.abs(fd))
    assert B.shape == (N, prob.nn)


@pytest.mark.xfail(strict=False, reason="first GPU run pending: written after the GPU budget of "
                   "round 2 was spent.  The host logic (mesh moved, tables rebuilt, engine "
                   "dropped) is covered on the CPU in tests/test_host_logic.py; the kernels are "
                   "the ones of the tests above")
def test_shape_parameter_moves_the_device_mesh(torch_cuda):
    """KelvinVoigtWShape: 'umesh' displaces the mesh (reference models/transient.py:347-360); the
    device assembles on the moved mesh: residual and Jacobian against the oracle built on the
    displaced coordinates, before and after the move, and after moving back."""
    from femvf_b200.models import transient
    from femvf_b200.residuals import solid as slr
    model = transient.FenicsModel(slr.KelvinVoigtWShape(*mesh_tuples()['m5']()))
    ref = model.residual.mesh().coordinates().copy()
    _assemble_and_compare(model, np.random.default_rng(21))
    engine0 = model.engine
    rng = np.random.default_rng(22)
    prop = model.prop.copy()
    du = 2e-3 * rng.standard_normal(ref.shape)
    prop['umesh'][:] = du.ravel()
    model.set_prop(prop)
    assert np.array_equal(model.residual.mesh().coordinates(), ref + du)
    _assemble_and_compare(model, np.random.default_rng(23))
    assert model.engine is not engine0
    prop = model.prop.copy()
    prop['umesh'][:] = 0.0
    model.set_prop(prop)
    assert np.array_equal(model.residual.mesh().coordinates(), ref)
    _assemble_and_compare(model, np.random.default_rng(24))
