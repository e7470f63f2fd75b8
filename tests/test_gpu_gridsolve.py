"""GPU parity: whole-GPU solves for meshes that do not fit one CTA (femvf_b200/gridsolve.py):
multicolour block ILU(0) (csrc/ilu.cu) against a host restatement, ILU(0)-GMRES against a sparse
LU of the oracle matrix, and the grid-wide Newton loop behind FenicsModel.solve_state1 /
static_solid_configuration against the oracle's Newton loop (transient.py:441-491, static.py:68-168)."""

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from helpers import mesh_tuples, oracle_problem, random_solid_prop, set_model_prop, random_state
from oracle import model as om

pytestmark = pytest.mark.gpu


def _setup_model(mt, rng, contact=False):
    from femvf_b200.models import transient
    from femvf_b200.residuals import solid as slr
    Model = transient.NodalContactModel if contact else transient.FenicsModel
    model = Model(slr.KelvinVoigt(*mt))
    prob = oracle_problem(model.residual)
    prop = random_solid_prop(prob, rng)
    mprop = model.prop.copy()
    set_model_prop(mprop, prop)
    model.set_prop(mprop)
    u1, u0, v0, a0 = random_state(prob.N, rng)
    u1 = u0 + 1e-4 * v0           # a plausible guess
    p1 = rng.uniform(0, 8e3, prob.nn)
    model.dt = 1e-4
    s0 = model.state0.copy(); s0['u'][:] = u0; s0['v'][:] = v0; s0['a'][:] = a0
    s1 = model.state1.copy(); s1['u'][:] = u1
    model.set_ini_state(s0); model.set_fin_state(s1)
    ctl = model.control.copy(); ctl['p'][:] = p1
    model.set_control(ctl)
    return model, prob, prop, (u1, u0, v0, a0), p1


def _ilu0_reference(A: sp.csr_matrix, order: np.ndarray):
    """Scalar ILU(0) of A in the elimination order `order` (dense loops: small matrices only).
    On a pattern made of full d x d blocks it defines the same M = L U as the block version."""
    P = A[order][:, order].toarray()
    pat = P != 0
    n = P.shape[0]
    for i in range(n):
        for k in np.nonzero(pat[i, :i])[0]:
            P[i, k] /= P[k, k]
            js = np.nonzero(pat[i, k + 1:] & pat[k, k + 1:])[0] + k + 1
            P[i, js] -= P[i, k] * P[k, js]
    L = np.tril(P, -1) + np.eye(n)
    U = np.triu(P)
    return L, U


def test_ilu0_apply_matches_host_restatement():
    import torch
    rng = np.random.default_rng(5)
    model, prob, prop, (u1, u0, v0, a0), p1 = _setup_model(mesh_tuples()['m5'](), rng)
    eng = model.engine
    model._push_all()
    eng.assemble(0, res=False, jac=True, dt=model.dt)
    ncol = eng.ilu_setup()
    assert 3 <= ncol <= 12
    eng.ilu_factor()
    r = rng.standard_normal(prob.N)
    rt = torch.as_tensor(r, device='cuda')
    zt = torch.empty_like(rt)
    eng.ilu_apply(rt, zt)
    # host: same colouring (deterministic seed), DOF order = nodes grouped by colour
    from femvf_b200 import tables
    color, rows, cptr = tables.color_node_graph(eng.tables['brptr'], eng.tables['bcol'], 0, eng.nn)
    assert np.array_equal(np.sort(rows), np.arange(eng.nn))
    # explicit zeros of the pattern must take part: build A from the pattern with 1e-300 floor
    J = om.SolidOracle(prob).jac(u1, model.dt, prop, p1).tocsr()
    J.data[J.data == 0] = 1e-300
    order = (2 * rows[:, None] + np.arange(2)[None, :]).ravel()
    L, U = _ilu0_reference(J, order)
    z_ref = np.empty(prob.N)
    z_ref[order] = np.linalg.solve(U, np.linalg.solve(L, r[order]))
    z = zt.cpu().numpy()
    assert np.max(np.abs(z - z_ref)) <= 1e-9 * np.max(np.abs(z_ref))


def test_grid_linear_solve_matches_sparse_lu():
    """63 k DOF (M5_CB refined 4x): restarted GMRES on the whole GPU vs splu of the oracle's J,
    with both preconditioners; ILU(0) needs clearly fewer iterations than block-Jacobi."""
    import torch
    from femvf_b200 import meshgen
    from femvf_b200.gridsolve import GridSolver
    rng = np.random.default_rng(11)
    model, prob, prop, (u1, u0, v0, a0), p1 = _setup_model(meshgen.m5_cb_refined(0.05, 4), rng)
    eng = model.engine
    model._push_all()
    eng.assemble(0, res=False, jac=True, dt=model.dt)
    b = rng.standard_normal(prob.N)
    b[model.residual.fixed_dofs()] = 0.0
    bt = torch.as_tensor(b, device='cuda')
    J_ref = om.SolidOracle(prob).jac(u1, model.dt, prop, p1)
    x_ref = spla.splu(J_ref.tocsc()).solve(b)
    iters = {}
    for precond in ('ilu0', 'jacobi'):
        gs = GridSolver(eng, precond=precond)
        xt = torch.empty_like(bt)
        info = gs.linear_solve(bt, xt, rtol=1e-13)
        err = np.linalg.norm(xt.cpu().numpy() - x_ref) / np.linalg.norm(x_ref)
        assert err <= 1e-9, (precond, err, info)
        iters[precond] = info['iterations']
    print(f"\n63k DOF transient J, rtol 1e-13: GMRES(40) iterations {iters}")
    assert iters['ilu0'] < 0.6 * iters['jacobi']


def test_grid_newton_behind_model_api(monkeypatch):
    """FenicsModel.solve_state1 and solve_dres_dstate1 on the grid-wide path (forced by a low
    VF_GRID_MIN_DOF) reproduce the oracle's Newton loop and sparse LU."""
    from femvf_b200 import meshgen
    monkeypatch.setenv('VF_GRID_MIN_DOF', '1000')
    rng = np.random.default_rng(3)
    model, prob, prop, (u1, u0, v0, a0), p1 = _setup_model(meshgen.m5_cb_refined(0.05, 3), rng)
    assert model._grid_solver() is not None
    s1 = model.state1.copy()
    x, info = model.solve_state1(s1)
    (u_ref, v_ref, a_ref), info_ref = om.SolidOracle(prob).solve_state1(
        (u0, v0, a0), model.dt, prop, p1, u_guess=u1)
    assert info['num_iter'] == info_ref['num_iter']
    assert np.max(np.abs(x['u'] - u_ref)) <= 1e-8 * np.max(np.abs(u_ref))   # north_star 1e-8
    assert np.max(np.abs(x['v'] - v_ref)) <= 1e-5 * np.max(np.abs(v_ref))
    assert np.max(np.abs(x['a'] - a_ref)) <= 1e-2 * np.max(np.abs(a_ref))
    # linearised solve with the (device-resident) Jacobian of the API
    model.set_fin_state(x)
    dres = model.assem_dres_dstate1()
    b = model.state1.copy()
    b['u'][:] = rng.standard_normal(prob.N); b['v'][:] = rng.standard_normal(prob.N)
    b['a'][:] = rng.standard_normal(prob.N)
    out = model.solve_dres_dstate1(dres, model.state1.copy(), b)
    J_ref = om.SolidOracle(prob).jac(np.asarray(x['u']), model.dt, prop, p1)
    xu_ref = spla.splu(J_ref.tocsc()).solve(np.asarray(b['u']))
    assert np.linalg.norm(out['u'] - xu_ref) <= 1e-9 * np.linalg.norm(xu_ref)
    assert dres.sub['u', 'state/u1'].on_device        # nothing was downloaded on the way


def test_grid_static_contact_solve(monkeypatch):
    """static_solid_configuration with contact at N = 4 118 (M5_CB refined 2x) on the grid path:
    Newton iterations and solution of the oracle."""
    from femvf_b200 import meshgen, static
    from femvf_b200.models import transient
    from femvf_b200.residuals import solid as slr
    monkeypatch.setenv('VF_GRID_MIN_DOF', '1000')
    model = transient.NodalContactModel(slr.KelvinVoigt(*meshgen.m5_cb_refined(0.05, 2)))
    prob = oracle_problem(model.residual)
    ymax = prob.coords[:, 1].max()
    prop = model.prop.copy()
    prop['emod'][:] = 1e5; prop['nu'][:] = 0.45; prop['eta'][:] = 5.0; prop['rho'][:] = 1.0
    prop['kcontact'][:] = 1e13; prop['ycontact'][:] = ymax - 0.01
    prop['ncontact'][:] = [0.0, 1.0]
    control = model.control.copy(); control['p'][:] = 0.0
    state, info = static.static_solid_configuration(model, control, prop)
    oprop = {k: np.array(v) for k, v in prop.items()}
    oprop['ycontact'] = float(prop['ycontact'][0]); oprop['kcontact'] = 1e13; oprop['nu'] = 0.45
    u_ref, info_ref = om.static_solid_configuration(om.SolidOracle(prob, contact=True), oprop,
                                                    np.zeros(prob.nn))
    assert info['num_iter'] == info_ref['num_iter']
    assert np.max(np.abs(state['u'] - u_ref)) <= 1e-7 * np.max(np.abs(u_ref))


def test_forward_integrate_on_the_grid_path(monkeypatch):
    """forward.integrate of the coupled model on a mesh routed to the whole-GPU solver (per-step
    API: grid Newton for the solid, fluid kernel, FSI maps) reproduces the in-kernel time loop."""
    from femvf_b200 import forward, meshgen
    from femvf_b200.load import load_fsi_model
    from femvf_b200.residuals import solid as slr, fluid as flr

    def build():
        mt = meshgen.m5_cb_refined(0.05, 1)
        return load_fsi_model(mt, slr.KelvinVoigt, flr.BernoulliAreaRatioSep,
                              {'dirichlet_bcs': {'state/u1': [(np.zeros(2), 'facet', 'fixed')]}}, {})

    def setup(model):
        state0 = model.state0.copy(); state0[:] = 0
        control = model.control.copy(); control[:] = 0; control['psub'][:] = 8e3
        prop = model.prop.copy()
        ymax = model.solid.residual.mesh().coordinates()[:, 1].max()
        prop['emod'][:] = 5e4; prop['rho'][:] = 1; prop['eta'][:] = 3; prop['nu'][:] = 0.45
        prop['ycontact'][:] = ymax + 0.05; prop['kcontact'][:] = 1e8; prop['ymid'][:] = 1.0
        return state0, control, prop
    times = 1e-4 * np.arange(9)
    ref = build()
    s0, c, p = setup(ref)
    fin_ref, _ = forward.integrate(ref, None, s0, [c], p, times, write=False)
    monkeypatch.setenv('VF_GRID_MIN_DOF', '500')
    model = build()
    assert model.solid._grid_solver() is not None
    s0, c, p = setup(model)
    fin, info = forward.integrate(model, None, s0, [c], p, times, write=False)
    for key in ('u', 'q', 'p'):
        scale = max(np.max(np.abs(fin_ref[key])), 1e-300)
        assert np.max(np.abs(fin[key] - fin_ref[key])) <= 1e-8 * scale, key
    assert info['num_iter'] >= 1


def test_banded_lu_matches_sparse_lu():
    """csrc/band.cu: banded LU in reverse Cuthill-McKee ordering + one refinement step against
    splu of the oracle matrix, transient and static (stiffness + contact penalty) matrices."""
    import torch
    from femvf_b200 import meshgen
    from femvf_b200.gridsolve import GridSolver
    rng = np.random.default_rng(13)
    model, prob, prop, (u1, u0, v0, a0), p1 = _setup_model(meshgen.m5_cb_refined(0.05, 2), rng,
                                                           contact=True)
    eng = model.engine
    model._push_all()
    gs = GridSolver(eng)
    assert gs.direct and gs.half_bandwidth < prob.N // 4
    for is_static in (False, True):
        eng.assemble(0, res=False, jac=True, dt=model.dt, is_static=is_static)
        vals = eng.download('J')
        rowptr, colidx = eng.csr_pattern()
        J = sp.csr_matrix((vals, colidx, rowptr), shape=(prob.N, prob.N))
        b = rng.standard_normal(prob.N)
        bt = torch.as_tensor(b, device='cuda')
        xt = torch.empty_like(bt)
        info = gs.linear_solve(bt, xt)
        x_ref = spla.splu(J.tocsc()).solve(b)
        err = np.linalg.norm(xt.cpu().numpy() - x_ref) / np.linalg.norm(x_ref)
        assert err <= 1e-10, (is_static, err, info)
