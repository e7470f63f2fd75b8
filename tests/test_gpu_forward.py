"""GPU parity: forward.integrate trajectories (device-resident loop) vs the oracle."""

import numpy as np
import pytest

from helpers import mesh_tuples, oracle_problem
from oracle import model as om

pytestmark = pytest.mark.gpu

TRAJ_TOL = 1e-8  # BASELINE.json north_star: displacement and glottal flow within 1e-8 relative


def build_fsi(mesh_name, Fluid=None, zs=None):
    from femvf_b200.load import load_fsi_model
    from femvf_b200.residuals import solid as slr, fluid as flr
    mt = mesh_tuples()[mesh_name]()
    d = mt[0].topology().dim()
    solid_kwargs = {'dirichlet_bcs': {'state/u1': [(np.zeros(d), 'facet', 'fixed')]}}
    return load_fsi_model(mt, slr.KelvinVoigt, Fluid or flr.BernoulliAreaRatioSep,
                          solid_kwargs, {}, zs=zs)


def benchmark_setup(model):
    """benchmarks/setup.py:34-49 settings (config 1)."""
    state0 = model.state0.copy(); state0[:] = 0
    control = model.control.copy(); control[:] = 0; control['psub'][:] = 8e3
    prop = model.prop.copy()
    ymax = model.solid.residual.mesh().coordinates()[:, 1].max()
    prop['emod'][:] = 5e4; prop['rho'][:] = 1; prop['eta'][:] = 3; prop['nu'][:] = 0.45
    prop['ycontact'][:] = ymax + 0.05; prop['kcontact'][:] = 1e8
    prop['ymid'][:] = ymax + 0.05 if False else 1.0
    return state0, control, prop


def oracle_run(model, state0, control, prop, times, fluid_kind='area_ratio'):
    prob = oracle_problem(model.solid.residual)
    co = om.CoupledOracle(om.SolidOracle(prob), model.fluid.residual.mesh(),
                          model.fsimap.dofs_solid, model.fsimap.dofs_fluid, fluid_kind)
    oprop = {k: np.array(v) for k, v in prop.items()}
    ctl = {'psub': control['psub'], 'psup': control['psup']}
    hist, infos = co.integrate(tuple(state0.vecs), [ctl], oprop, times)
    return hist, infos


# ('m5', 100) is BASELINE config 1 in full: times = 1e-4 arange(100), 99 steps
# (benchmarks/setup.py:34-49)
@pytest.mark.parametrize('mesh_name,ntimes', [('m5', 30), ('square5', 30), ('m5', 100)])
def test_integrate_matches_oracle(mesh_name, ntimes, tmp_path):
    import torch
    assert torch.cuda.is_available()
    from femvf_b200 import forward, statefile as sf
    model = build_fsi(mesh_name)
    state0, control, prop = benchmark_setup(model)
    if mesh_name == 'square5':
        prop['ymid'][:] = 1.05
    times = 1e-4 * np.arange(ntimes)
    path = str(tmp_path / 'out.h5')
    with sf.StateFile(model, path, mode='w') as f:
        fin_state, info = forward.integrate(model, f, state0, [control], prop, times)
        hist, infos = oracle_run(model, state0, control, prop, times)
        assert f.size == len(times)
        for n in (1, len(times) // 2, len(times) - 1):
            st = f.get_state(n)
            for k, key in enumerate(('u', 'v', 'a', 'q', 'p')):
                ref = hist[n][k]
                scale = max(np.max(np.abs(ref)), 1e-300)
                # v, a amplify the u error by 2/dt, 4/dt^2 (Newmark); the reference's own
                # Newton tolerance (abs 1e-8) bounds them no tighter than this
                tol = {'v': 1e-5, 'a': 1e-2}.get(key, TRAJ_TOL)
                assert np.max(np.abs(st[key] - ref)) <= tol * scale, (n, key)
        # glottal flow at EVERY step
        q_ref = np.array([h[3][0] for h in hist])
        q = np.array([f.get_state(n)['q'][0] for n in range(len(times))])
        assert np.max(np.abs(q - q_ref)) <= TRAJ_TOL * np.max(np.abs(q_ref))
    assert abs(fin_state['q'][0] - q_ref[-1]) <= TRAJ_TOL * abs(q_ref[-1])
    assert info['num_iter'] >= 1


def test_per_step_api_equals_device_loop():
    """integrate_step through the host-buffer model API (set_* + solve_state1 per step) gives
    the same states as the device-resident loop used by forward.integrate."""
    from femvf_b200 import forward
    model = build_fsi('m5')
    state0, control, prop = benchmark_setup(model)
    times = 1e-4 * np.arange(8)
    fin_dev, _ = forward.integrate(model, None, state0, [control], prop, times, write=False)
    model.set_prop(prop)
    st = state0
    for n in range(len(times) - 1):
        st, info = forward.integrate_step(model, st, control, prop, times[n + 1] - times[n])
    for key in ('u', 'q', 'p'):
        scale = max(np.max(np.abs(fin_dev[key])), 1e-300)
        assert np.max(np.abs(st[key] - fin_dev[key])) <= 1e-12 * scale, key
    assert set(info) >= {'num_iter', 'abs_err', 'rel_err'}


def test_golden_forward_fixture_on_gpu():
    import os
    from femvf_b200 import forward
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'forward_m5.npz'))
    model = build_fsi('m5')
    state0, control, prop = benchmark_setup(model)
    model.set_prop(prop)
    model.set_ini_state(state0)
    model.push_to_device()
    states, infos = model.device_integrate(np.diff(z['times']), [control])
    N = model.solid.state0['u'].size
    q = states[:, 3 * N]
    assert np.max(np.abs(q - z['q'])) <= TRAJ_TOL * np.max(np.abs(z['q']))
    assert np.max(np.abs(states[-1, :N] - z['u_last'])) <= TRAJ_TOL * np.max(np.abs(z['u_last']))
    # glottal-width series (min fluid area) written by the kernel
    gw = infos[1:, 3]
    assert np.max(np.abs(gw - z['min_area'][1:])) <= TRAJ_TOL * np.max(np.abs(z['min_area'][1:]))


@pytest.mark.parametrize('variant', ['contact', 'epithelium'])
def test_coupled_variants_match_oracle(variant):
    """NodalContactModel solid (active contact) and the membrane residual inside the coupled loop."""
    from femvf_b200 import forward
    from femvf_b200.load import derive_1D_interface_from_facet_subdomain
    from femvf_b200.models import transient
    from femvf_b200.residuals import solid as slr, fluid as flr
    mt = mesh_tuples()['m5']()
    Residual = slr.KelvinVoigtWEpithelium if variant == 'epithelium' else slr.KelvinVoigt
    residual = Residual(*mt)
    solid = transient.NodalContactModel(residual) if variant == 'contact' \
        else transient.FenicsModel(residual)
    s, sdofs, fdofs = derive_1D_interface_from_facet_subdomain(
        residual.mesh(), None, residual.mesh_function('facet'),
        {residual.mesh_subdomain('facet')['pressure']})
    fluid = transient.JaxModel(flr.BernoulliAreaRatioSep(s))
    model = transient.ExplicitFSIModel(solid, fluid, sdofs, fdofs)
    state0, control, prop = benchmark_setup(model)
    ymax = residual.mesh().coordinates()[:, 1].max()
    prop['ymid'][:] = ymax + 0.02
    if variant == 'contact':
        prop['ycontact'][:] = ymax + 0.001   # the surface bulges into the plane within a few steps
        prop['kcontact'][:] = 1e11
    else:
        prop['emod_membrane'][:] = 2e5; prop['th_membrane'][:] = 0.005; prop['nu_membrane'][:] = 0.45
    times = 1e-4 * np.arange(25)
    fin, info = forward.integrate(model, None, state0, [control], prop, times, write=False)
    prob = oracle_problem(residual)
    co = om.CoupledOracle(om.SolidOracle(prob, contact=variant == 'contact',
                                         membrane=variant == 'epithelium'), s, sdofs, fdofs)
    oprop = {k: np.array(v) for k, v in prop.items()}
    hist, infos = co.integrate(tuple(state0.vecs),
                               [{'psub': control['psub'], 'psup': control['psup']}], oprop, times)
    if variant == 'contact':
        gap = (prob.coords[:, 1] + hist[-1][0][1::2]) - float(prop['ycontact'][0])
        assert gap.max() > 0, "the test should exercise active contact"
    for k, key in ((0, 'u'), (3, 'q'), (4, 'p')):
        ref = hist[-1][k]
        assert np.max(np.abs(fin[key] - ref)) <= TRAJ_TOL * np.max(np.abs(ref)), key


def test_3d_coupled_steps_match_oracle():
    """Unit-cube fixture with three fluid planes: tetrahedral kernels + nonlinear 3D pressure."""
    from femvf_b200 import forward
    zs = np.linspace(0, 1, 3)
    model = build_fsi('cube332', zs=zs)
    state0 = model.state0.copy(); state0[:] = 0
    control = model.control.copy(); control['psub'][:] = 8e3; control['psup'][:] = 0
    prop = model.prop.copy()
    prop['emod'][:] = 1e5; prop['rho'][:] = 1; prop['eta'][:] = 4e-3; prop['nu'][:] = 0.45
    prop['ymid'][:] = 1.05
    times = 2e-5 * np.arange(10)
    fin, info = forward.integrate(model, None, state0, [control], prop, times, write=False)
    prob = oracle_problem(model.solid.residual)
    co = om.CoupledOracle(om.SolidOracle(prob), model.fluid.residual.mesh(),
                          model.fsimap.dofs_solid, model.fsimap.dofs_fluid)
    oprop = {k: np.array(v) for k, v in prop.items()}
    hist, infos = co.integrate(tuple(state0.vecs),
                               [{'psub': control['psub'], 'psup': control['psup']}], oprop, times)
    for k, key in ((0, 'u'), (3, 'q'), (4, 'p')):
        ref = hist[-1][k]
        assert np.max(np.abs(fin[key] - ref)) <= TRAJ_TOL * max(np.max(np.abs(ref)), 1e-300), key
    assert info['num_iter'] >= 1


def test_ensemble_members_and_host_entry_point():
    """Members with identical inputs reproduce each other bit for bit and the single simulation
    to solver tolerance; the
    host-buffer entry point equals the device-resident one; different members differ."""
    from femvf_b200.ensemble import EnsembleRunner
    model = build_fsi('m5')
    state0, control, prop = benchmark_setup(model)
    B = 5
    runner = EnsembleRunner(model, B)
    runner.set_common_prop(prop)
    ne = runner.ne
    rng = np.random.default_rng(0)
    emod = np.tile(np.asarray(prop['emod']), (B, 1)); eta = np.tile(np.asarray(prop['eta']), (B, 1))
    emod[3] *= np.exp(0.3 * rng.standard_normal(ne))
    ini = np.zeros((B, runner.state_size))
    dts = np.full(12, 1e-4)
    ctl = np.array([[[8e3], [0.0]]])
    fin, series = runner.run_host(dts, ctl, ini, emod, eta)
    # the single simulation with bit-identical step sizes
    model.set_prop(prop)
    model.set_ini_state(state0)
    model.push_to_device()
    states, _ = model.device_integrate(dts, [control])
    ref = states[-1]
    # members with identical inputs are bit-identical to each other; the single simulation
    # runs on its own engine, whose time loop preconditions GMRES differently (dense inverse
    # for single simulations, polynomial for ensembles), so it agrees to solver tolerance
    for b in (1, 2, 4):
        assert np.array_equal(fin[b], fin[0]), b
    scale = np.max(np.abs(ref))
    assert np.max(np.abs(fin[0] - ref)) <= 1e-9 * scale
    assert np.max(np.abs(fin[3] - ref)) > 1e-6 * scale
    # device-resident path gives the same numbers
    runner.upload_members(ini, emod, eta)
    hs, hi = runner.run_device(dts, ctl, store_states=True)
    assert np.array_equal(hs[:, -1, :].cpu().numpy(), fin)
    assert np.array_equal(hi.cpu().numpy(), series)


def test_rayleigh_coupled_matches_oracle():
    """slr.Rayleigh is one of the solid residuals of the reference's test_forward.py:36-41."""
    from femvf_b200 import forward
    from femvf_b200.load import load_fsi_model
    from femvf_b200.residuals import solid as slr, fluid as flr
    mt = mesh_tuples()['m5']()
    model = load_fsi_model(mt, slr.Rayleigh, flr.BernoulliAreaRatioSep,
                           {'dirichlet_bcs': {'state/u1': [(np.zeros(2), 'facet', 'fixed')]}}, {})
    assert model.solid.prop.keys() == ['rho', 'emod', 'nu', 'rayleigh_m', 'rayleigh_k',
                                       'ycontact', 'ncontact', 'kcontact']
    state0 = model.state0.copy(); state0[:] = 0
    control = model.control.copy(); control[:] = 0; control['psub'][:] = 8e3
    prop = model.prop.copy()
    prop['emod'][:] = 5e4; prop['rho'][:] = 1; prop['nu'][:] = 0.45
    prop['rayleigh_m'][:] = 10.0; prop['rayleigh_k'][:] = 3e-5; prop['ymid'][:] = 1.0
    times = 1e-4 * np.arange(20)
    fin, info = forward.integrate(model, None, state0, [control], prop, times, write=False)
    prob = oracle_problem(model.solid.residual)
    co = om.CoupledOracle(om.SolidOracle(prob), model.fluid.residual.mesh(),
                          model.fsimap.dofs_solid, model.fsimap.dofs_fluid)
    oprop = {k: np.array(v) for k, v in prop.items()}
    hist, _ = co.integrate(tuple(state0.vecs),
                           [{'psub': control['psub'], 'psup': control['psup']}], oprop, times)
    for k, key in ((0, 'u'), (3, 'q'), (4, 'p')):
        ref = hist[-1][k]
        assert np.max(np.abs(fin[key] - ref)) <= TRAJ_TOL * np.max(np.abs(ref)), key


def test_integrate_extend_continues_a_statefile(tmp_path):
    """forward.integrate_extend (forward.py:105-136): restarting from the last stored state
    gives the same trajectory as one uninterrupted run."""
    from femvf_b200 import forward, statefile as sf
    model = build_fsi('m5')
    state0, control, prop = benchmark_setup(model)
    times = 1e-4 * np.arange(13)
    full, _ = forward.integrate(model, None, state0, [control], prop, times, write=False)
    path = str(tmp_path / 'restart.h5')
    with sf.StateFile(model, path, mode='w') as f:
        forward.integrate(model, f, state0, [control], prop, times[:7])
    with sf.StateFile(model, path, mode='a') as f:
        assert f.size == 7
        fin, _ = forward.integrate_extend(model, f, [control], times[6:] - times[6])
        assert f.size == 13
        assert np.allclose(f.get_times(), times, rtol=0, atol=1e-18)
    for key in ('u', 'q', 'p'):
        scale = np.max(np.abs(full[key]))
        assert np.max(np.abs(fin[key] - full[key])) <= 1e-9 * scale, key


def test_implicit_coupling_matches_oracle():
    """ImplicitFSIModel: fixed-point coupling, per-step host loop over device solves."""
    from femvf_b200 import forward
    from femvf_b200.load import load_fsi_model
    from femvf_b200.residuals import solid as slr, fluid as flr
    mt = mesh_tuples()['m5']()
    model = load_fsi_model(mt, slr.KelvinVoigt, flr.BernoulliAreaRatioSep,
                           {'dirichlet_bcs': {'state/u1': [(np.zeros(2), 'facet', 'fixed')]}}, {},
                           coupling='implicit')
    state0, control, prop = benchmark_setup(model)
    times = 1e-4 * np.arange(6)
    fin, info = forward.integrate(model, None, state0, [control], prop, times, write=False)
    prob = oracle_problem(model.solid.residual)
    co = om.ImplicitCoupledOracle(om.SolidOracle(prob), model.fluid.residual.mesh(),
                                  model.fsimap.dofs_solid, model.fsimap.dofs_fluid)
    oprop = {k: np.array(v) for k, v in prop.items()}
    hist, infos = co.integrate(tuple(state0.vecs),
                               [{'psub': control['psub'], 'psup': control['psup']}], oprop, times)
    for k, key in ((0, 'u'), (3, 'q'), (4, 'p')):
        ref = hist[-1][k]
        assert np.max(np.abs(fin[key] - ref)) <= 1e-7 * np.max(np.abs(ref)), key
    assert info['num_iter'] == infos[-1]['num_iter'] and info['num_iter'] > 1


def test_dense_inverse_preconditioner_opt_in(monkeypatch):
    """The experimental dense-inverse preconditioner (VF_DENSE_PREC=1) must give the same
    trajectory as the default polynomial one to solver tolerance (it only changes how the
    Newton linear systems are solved)."""
    import numpy as np
    import bench
    from femvf_b200 import forward
    times = 1e-4 * np.arange(12)
    out = {}
    for flag in ('0', '1'):
        monkeypatch.setenv('VF_DENSE_PREC', flag)
        fm = bench.fsi_model()
        state0, control, prop = bench.config1_args(fm)
        fin, info = forward.integrate(fm, None, state0, [control], prop, times, write=False)
        out[flag] = (np.asarray(fin['u']).copy(), float(np.asarray(fin['q'])[0]),
                     fm.engine.download('info')[3])
    u0, q0, it0 = out['0']
    u1, q1, it1 = out['1']
    assert np.max(np.abs(u1 - u0)) <= 1e-8 * np.max(np.abs(u0))
    assert abs(q1 - q0) <= 1e-8 * abs(q0)
    assert it1 < it0          # fewer Krylov iterations with the inverse


def test_coupled_model_linearisation_matches_finite_differences():
    """ExplicitFSIModel.assem_dres_dstate1 / assem_dres_dstate0 / solve_dres_dstate1
    (transient.py:873-896, 922-937) against central differences of assem_res."""
    from femvf_b200 import forward
    model = build_fsi('m5')
    state0, control, prop = benchmark_setup(model)
    model.set_prop(prop)
    model.set_control(control)
    model.dt = 1e-4
    # a state on the trajectory (non-trivial area, flow and pressure)
    fin, _ = forward.integrate(model, None, state0, [control], prop, 1e-4 * np.arange(12),
                               write=False)
    st0 = fin.copy()
    st1, _ = forward.integrate_step(model, st0, control, prop, 1e-4)
    rng = np.random.default_rng(0)
    st1['u'][:] += 1e-6 * rng.standard_normal(st1['u'].size)   # off the solution: F != 0
    model.set_ini_state(st0)
    model.set_fin_state(st1)
    d1 = model.assem_dres_dstate1()
    d0 = model.assem_dres_dstate0()
    keys = ('u', 'v', 'a', 'q', 'p')

    def res_at(s0, s1):
        model.set_ini_state(s0)
        model.set_fin_state(s1)
        r = model.assem_res()
        return {k: np.array(r[k]) for k in keys}

    def column(dres, ckey, j, suffix):
        return {k: np.asarray(dres.sub[k, f'state/{ckey}{suffix}'].tocsr()[:, [j]].todense()).ravel()
                for k in keys}
    ndim = 2
    fsi_y = np.asarray(model.fsimap.dofs_solid) * ndim + 1
    # state1: displacement of FSI surface nodes (moves the area -> q and p), an interior DOF,
    # and the fluid state itself
    for ckey, j, h in [('u', int(fsi_y[len(fsi_y) // 2]), 1e-7), ('u', int(fsi_y[3]), 1e-7),
                       ('u', 5, 1e-7), ('v', 7, 1e-4), ('a', 9, 1e-1), ('q', 0, 1e-3),
                       ('p', 4, 1e-2)]:
        sp_, sm = st1.copy(), st1.copy()
        sp_[ckey][j] += h; sm[ckey][j] -= h
        rp, rm = res_at(st0, sp_), res_at(st0, sm)
        col = column(d1, ckey, j, '1')
        for k in keys:
            fd = (rp[k] - rm[k]) / (2 * h)
            scale = max(np.abs(col[k]).max(), np.abs(fd).max(), 1e-30)
            assert np.max(np.abs(fd - col[k])) <= 2e-5 * scale + 1e-9, (ckey, j, k)
    # state0: u0, v0, a0 through the Newmark relations, p0 through the pressure map
    for ckey, j, h in [('u', 11, 1e-7), ('v', 12, 1e-4), ('a', 13, 1e-1), ('p', 10, 1e-2),
                       ('q', 0, 1e-3)]:
        sp_, sm = st0.copy(), st0.copy()
        sp_[ckey][j] += h; sm[ckey][j] -= h
        rp, rm = res_at(sp_, st1), res_at(sm, st1)
        col = column(d0, ckey, j, '0')
        # as in the reference (transient.py:408-421) no Dirichlet condition is applied to the
        # state0 blocks, while assem_res zeroes F_u on the fixed DOFs: compare the free rows
        col['u'][model.solid.residual.fixed_dofs()] = 0.0
        for k in keys:
            fd = (rp[k] - rm[k]) / (2 * h)
            scale = max(np.abs(col[k]).max(), np.abs(fd).max(), 1e-30)
            assert np.max(np.abs(fd - col[k])) <= 2e-5 * scale + 1e-9, (ckey, j, k)
    # linear solve: x = (dF/dstate1)^-1 b, checked by multiplying back block by block
    model.set_ini_state(st0)
    model.set_fin_state(st1)
    d1 = model.assem_dres_dstate1()
    b = model.state1.copy()
    for k in keys:
        b[k][:] = rng.standard_normal(b[k].size)
    x = model.solve_dres_dstate1(b, d1)
    for k in keys:
        acc = np.zeros(b[k].size)
        for c in keys:
            acc += d1.sub[k, f'state/{c}1'] @ np.asarray(x[c])
        err = float(np.max(np.abs(acc - b[k])))
        # the device GMRES stops on the left-preconditioned residual (1e-13): the true residual of
        # the u rows is ~1e-8 of |b|
        assert err <= 1e-7 * max(float(np.abs(b[k]).max()), 1.0), (k, err)


def test_ensemble_statefiles_match_single_runs(tmp_path):
    """EnsembleRunner.run_to_statefiles: every member's StateFile equals the file written by
    forward.integrate for that member's properties (the reference's output layout, statefile.py)."""
    from femvf_b200 import forward, statefile as sf
    from femvf_b200.ensemble import EnsembleRunner
    model = build_fsi('m5')
    state0, control, prop = benchmark_setup(model)
    times = 1e-4 * np.arange(13)
    B = 3
    runner = EnsembleRunner(model, B)
    rng = np.random.default_rng(1)
    emod = 5e4 * np.exp(0.3 * rng.standard_normal((B, runner.ne)))
    eta = 3.0 * np.exp(0.3 * rng.standard_normal((B, runner.ne)))
    ini = np.zeros((B, runner.state_size))
    names = [str(tmp_path / f'member{b}.h5') for b in range(B)]
    infos = runner.run_to_statefiles(names, times, [control], prop, ini, emod, eta, nchunk=5)
    assert infos.shape == (B, len(times), 4)
    single = build_fsi('m5')
    for b in range(B):
        p = prop.copy(); p['emod'][:] = emod[b]; p['eta'][:] = eta[b]
        ref_name = str(tmp_path / f'ref{b}.h5')
        with sf.StateFile(single, ref_name, mode='w') as f:
            forward.integrate(single, f, state0, [control], p, times)
        with sf.StateFile(single, ref_name, mode='r') as fr, \
                sf.StateFile(single, names[b], mode='r') as fe:
            assert fe.size == fr.size == len(times)
            for n in (0, 1, 6, len(times) - 1):
                a, r = fe.get_state(n), fr.get_state(n)
                for key in ('u', 'q', 'p'):
                    scale = max(np.max(np.abs(r[key])), 1e-300)
                    assert np.max(np.abs(a[key] - r[key])) <= 1e-12 * scale, (b, n, key)
            assert np.allclose(fe.get_prop()['emod'], emod[b])
