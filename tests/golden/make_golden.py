"""
Generate the regression fixtures under tests/golden/ (run from the repo root:
``python tests/golden/make_golden.py``).

These vectors are produced by the ORACLE (numpy restatement), not by the reference: the
reference's FEniCS/PETSc/JAX stack cannot be imported in this container and
its tests hold no known answers (F6).  They pin the oracle against accidental change and
give the GPU tests a fixed target that does not depend on importing ``oracle`` at all.
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, 'vf-fem_b200'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

from helpers import mesh_tuples, oracle_problem, random_solid_prop, random_state  # noqa: E402
from femvf_b200.residuals import solid as slr  # noqa: E402
from oracle import fem, fluid as ofl, model as om  # noqa: E402

ASSEMBLY_CASES = {
    'square5_kv': ('square5', False, False),
    'cube332_kv': ('cube332', False, False),
    'm5_epi_contact': ('m5', True, True),
}


def assembly_inputs(name):
    mesh_name, membrane, contact = ASSEMBLY_CASES[name]
    rng = np.random.default_rng(abs(hash(name)) % 1000 if False else len(name))
    Residual = slr.KelvinVoigtWEpithelium if membrane else slr.KelvinVoigt
    res = Residual(*mesh_tuples()[mesh_name]())
    prob = oracle_problem(res)
    prop = random_solid_prop(prob, rng, membrane=membrane)
    u1, u0, v0, a0 = random_state(prob.N, rng)
    p1 = rng.uniform(0, 8e3, prob.nn)
    return res, prob, prop, (u1, u0, v0, a0), p1, 1e-4, membrane, contact


def assembly_case(name):
    res, prob, prop, (u1, u0, v0, a0), p1, dt, membrane, contact = assembly_inputs(name)
    so = om.SolidOracle(prob, contact=contact, membrane=membrane)
    J = so.jac(u1, dt, prop, p1)
    return {'rowptr': prob.rowptr, 'colidx': prob.colidx, 'J': J.data,
            'F': so.res(u1, (u0, v0, a0), dt, prop, p1)}


def bernoulli_case():
    """The reference's fluid fixture (tests/residuals/test_fluid.py:12-45)."""
    s = np.linspace(0, 1, 11)
    area = np.abs(s - 0.5)
    area[area < 0.1] = 0.1
    psub, psup, rho = np.array([100.0]), np.array([0.0]), np.array([1.0])
    out = {'s': s, 'area': area}
    q, p = ofl.bernoulli_area_ratio_sep(s, area, psub, psup, rho, np.array([1.0]), np.array([0.0]))
    out['q_area_ratio'], out['p_area_ratio'] = q, p
    q, p = ofl.bernoulli_area_ratio_sep(s, area, psub, psup, rho, np.array([1.2]), np.array([0.0]))
    out['q_area_ratio_r12'], out['p_area_ratio_r12'] = q, p
    q, p = ofl.bernoulli_fixed_sep(s, area, psub, psup, rho, 5)
    out['q_fixed'], out['p_fixed'] = q, p
    q, p = ofl.bernoulli_smooth_min_sep(s, area, psub, psup, rho, np.array([1e-2]), np.array([1e-2]))
    out['q_smooth'], out['p_smooth'] = q, p
    return out


def forward_case():
    """30 coupled steps of config 1 on the M5_CB mesh (benchmarks/setup.py:34-49)."""
    from test_gpu_forward import build_fsi, benchmark_setup, oracle_run
    model = build_fsi('m5')
    state0, control, prop = benchmark_setup(model)
    times = 1e-4 * np.arange(31)
    hist, infos = oracle_run(model, state0, control, prop, times)
    return {'times': times,
            'u_last': hist[-1][0], 'q': np.array([h[3][0] for h in hist]),
            'p_last': hist[-1][4],
            'min_area': np.array([np.min(i['area']) if 'area' in i else np.nan for i in infos])}


if __name__ == '__main__':
    for name in ASSEMBLY_CASES:
        np.savez_compressed(os.path.join(HERE, f'assembly_{name}.npz'), **assembly_case(name))
    np.savez_compressed(os.path.join(HERE, 'bernoulli.npz'), **bernoulli_case())
    np.savez_compressed(os.path.join(HERE, 'forward_m5.npz'), **forward_case())
    print('golden fixtures written to', HERE)
