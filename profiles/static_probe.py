"""Static contact solve (config 2 settings) on refined M5_CB meshes through the model API:
Newton iterations, GMRES iterations and time of the one-CTA and the whole-GPU paths."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'vf-fem_b200')]
import numpy as np, torch
from femvf_b200 import meshgen, static
from femvf_b200.models import transient
from femvf_b200.residuals import solid as slr
for levels, min_dof, direct in ((1, 10**9, '1'), (1, 500, '1'), (2, 500, '0'), (2, 500, '1'),
                                (3, 500, '0'), (3, 500, '1')):
    os.environ['VF_GRID_MIN_DOF_STATIC'] = str(min_dof)
    os.environ['VF_GRID_DIRECT'] = direct
    model = transient.NodalContactModel(slr.KelvinVoigt(*meshgen.m5_cb_refined(0.05, levels)))
    ymax = model.residual.mesh().coordinates()[:, 1].max()
    prop = model.prop.copy()
    prop['emod'][:] = 1e5; prop['nu'][:] = 0.45; prop['eta'][:] = 5.0; prop['rho'][:] = 1.0
    prop['kcontact'][:] = 1e13; prop['ycontact'][:] = ymax - 0.01; prop['ncontact'][:] = [0.0, 1.0]
    control = model.control.copy(); control['p'][:] = 0.0
    static.static_solid_configuration(model, control, prop)          # warm-up
    torch.cuda.synchronize(); t0 = time.perf_counter()
    state, info = static.static_solid_configuration(model, control, prop)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    gs = model._grid_solver(static=True)
    print(json.dumps({'levels': levels, 'dof': int(state['u'].size),
                      'path': ('banded LU' if gs.direct else 'whole GPU ILU(0)-GMRES') if gs is not None else 'one CTA',
                      'newton_iterations': info['num_iter'], 'ms': round(1e3 * dt, 2),
                      'abs_err': info['abs_err']}), flush=True)
