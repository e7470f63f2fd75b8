"""GPU parity: Bernoulli fluid kernels and the static solid solve vs the oracle / fixtures."""

import os

import numpy as np
import pytest

from helpers import mesh_tuples, oracle_problem
from oracle import fem, fluid as ofl, model as om

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def make_fluid(Fluid, s, **kw):
    from femvf_b200.load import load_jax_model
    return load_jax_model(s, Fluid, **kw)


def run_fluid(model, area, psub, psup, prop_updates):
    ctl = model.control.copy()
    ctl['area'][:] = np.ravel(area); ctl['psub'][:] = psub; ctl['psup'][:] = psup
    model.set_control(ctl)
    prop = model.prop.copy()
    for k, v in prop_updates.items():
        prop[k][:] = v
    model.set_prop(prop)
    qp, _ = model.solve_state1(model.state1)
    return qp['q'].copy(), qp['p'].copy()


def test_bernoulli_reference_fixture_all_kinds():
    from femvf_b200.residuals import fluid as flr
    z = np.load(os.path.join(GOLDEN, 'bernoulli.npz'))
    s, area = z['s'], z['area']
    q, p = run_fluid(make_fluid(flr.BernoulliAreaRatioSep, s), area, 100.0, 0.0, {'r_sep': 1.0})
    assert np.allclose(q, np.ravel(z['q_area_ratio']), rtol=1e-13)
    assert np.allclose(p, np.ravel(z['p_area_ratio']), rtol=1e-12, atol=1e-10)
    q, p = run_fluid(make_fluid(flr.BernoulliAreaRatioSep, s), area, 100.0, 0.0, {'r_sep': 1.2})
    assert np.allclose(q, np.ravel(z['q_area_ratio_r12']), rtol=1e-13)
    assert np.allclose(p, np.ravel(z['p_area_ratio_r12']), rtol=1e-12, atol=1e-10)
    q, p = run_fluid(make_fluid(flr.BernoulliFixedSep, s, idx_sep=5), area, 100.0, 0.0, {})
    assert np.allclose(q, np.ravel(z['q_fixed']), rtol=1e-13)
    assert np.allclose(p, np.ravel(z['p_fixed']), rtol=1e-12, atol=1e-10)
    q, p = run_fluid(make_fluid(flr.BernoulliSmoothMinSep, s), area, 100.0, 0.0,
                     {'zeta_min': 1e-2, 'zeta_sep': 1e-2})
    assert np.allclose(q, np.ravel(z['q_smooth']), rtol=1e-12)
    assert np.allclose(p, np.ravel(z['p_smooth']), rtol=1e-11, atol=1e-9)


@pytest.mark.parametrize('ns', [2, 31, 32, 33, 200])
def test_bernoulli_random_and_edge_sizes(ns):
    """ragged sizes around the warp width, ties in the minimum, reversed flow, lower bound"""
    from femvf_b200.residuals import fluid as flr
    rng = np.random.default_rng(ns)
    s = np.cumsum(rng.uniform(0.01, 0.05, ns)); s -= s[0]
    area = rng.uniform(0.05, 1.0, ns)
    if ns > 4:
        area[ns // 2] = area[ns // 2 + 1] = area.min() * 0.5  # tie: first index wins
    for psub, psup, r_sep, alb in [(800.0, 0.0, 1.0, 0.0), (0.0, 500.0, 1.3, 0.0),
                                   (300.0, 100.0, 1.1, 0.2)]:
        model = make_fluid(flr.BernoulliAreaRatioSep, s)
        q, p = run_fluid(model, area, psub, psup, {'r_sep': r_sep, 'area_lb': alb, 'rho_air': 1.2e-3})
        qo, po = ofl.bernoulli_area_ratio_sep(s, area, np.array([psub]), np.array([psup]),
                                              np.array([1.2e-3]), np.array([r_sep]), np.array([alb]))
        assert np.allclose(q, qo, rtol=1e-13), (ns, psub)
        assert np.allclose(p, po, rtol=1e-12, atol=1e-12 * max(abs(psub), abs(psup))), (ns, psub)
        # res = state1 - (q, p): the reference's residual definition (fluid.py:286-294)
        res = model.assem_res()
        assert np.allclose(res['q'], model.state1['q'] - q)


def test_bernoulli_multi_plane():
    from femvf_b200.residuals import fluid as flr
    rng = np.random.default_rng(5)
    s1 = np.linspace(0, 1, 17)
    S = np.stack([s1, s1 * 1.1, s1 * 0.9])
    A = rng.uniform(0.1, 1.0, S.shape)
    model = make_fluid(flr.BernoulliAreaRatioSep, S)
    q, p = run_fluid(model, A, np.array([800.0, 600.0, 400.0]), np.array([0.0, 10.0, 20.0]),
                     {'r_sep': 1.1})
    col = lambda v: np.reshape(v, (3, 1))
    qo, po = ofl.bernoulli_area_ratio_sep(S, A, col([800.0, 600.0, 400.0]), col([0.0, 10.0, 20.0]),
                                          col([1.0] * 3), col([1.1] * 3), col([0.0] * 3))
    assert np.allclose(q, np.ravel(qo), rtol=1e-13)
    assert np.allclose(p, np.ravel(po), rtol=1e-12, atol=1e-9)


# (-0.1, 1e13) is SURVEY.md section 8(d) config 2 as written (examples/prephonatory_gap.py:46)
@pytest.mark.parametrize('offset,kcontact,utol', [(-0.01, 1e13, 1e-8), (-0.002, 1e11, 1e-8),
                                                  (-0.1, 1e13, 1e-6)])
def test_static_solve_with_contact(offset, kcontact, utol):
    """config 2: static prephonatory solve with the cubic contact penalty (static.py:68-168)."""
    from femvf_b200 import static
    from femvf_b200.models import transient
    from femvf_b200.residuals import solid as slr
    model = transient.NodalContactModel(slr.KelvinVoigt(*mesh_tuples()['m5']()))
    prob = oracle_problem(model.residual)
    ymax = prob.coords[:, 1].max()
    prop = model.prop.copy()
    prop['emod'][:] = 1e5; prop['nu'][:] = 0.45; prop['eta'][:] = 5.0; prop['rho'][:] = 1.0
    prop['kcontact'][:] = kcontact; prop['ycontact'][:] = ymax + offset
    prop['ncontact'][:] = [0.0, 1.0]
    control = model.control.copy(); control['p'][:] = 0.0
    state, info = static.static_solid_configuration(model, control, prop)
    oprop = {k: np.array(v) for k, v in prop.items()}
    oprop['ycontact'] = float(prop['ycontact'][0]); oprop['kcontact'] = float(prop['kcontact'][0])
    oprop['nu'] = 0.45
    u_ref, info_ref = om.static_solid_configuration(om.SolidOracle(prob, contact=True), oprop,
                                                    np.zeros(prob.nn))
    assert info['abs_err'] <= 1e-8 or info['rel_err'] <= 1e-10
    scale = np.max(np.abs(u_ref))
    # deep contact (10% of the fold height pressed into a k = 1e13 penalty): both Newton loops stop
    # on the RELATIVE criterion with a residual of ~6e-2 dyn, which bounds u no tighter than 1e-6
    assert np.max(np.abs(state['u'] - u_ref)) <= utol * scale
    assert info['num_iter'] == info_ref['num_iter']
    if utol > 1e-8:
        # ... so check the device solution against the equations themselves: its oracle residual
        # is as small as the oracle's own at convergence
        so = om.SolidOracle(prob, contact=True)
        sprop = dict(oprop); sprop['rho'] = np.zeros(prob.ne); sprop['eta'] = np.zeros(prob.ne)
        u = np.asarray(state['u'])
        zero = np.zeros(prob.N)
        r_gpu = so.res(u, (u, zero, zero), 1.0, sprop, np.zeros(prob.nn))
        assert np.linalg.norm(r_gpu) <= 10 * max(info_ref['abs_err'], 1e-8)
