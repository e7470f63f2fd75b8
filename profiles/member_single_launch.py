"""One 99-step launch of member_kernel on BASELINE configs[0] (target of the ncu capture)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'vf-fem_b200'), os.path.join(ROOT, 'tests')]
import numpy as np, torch
import bench
fm = bench.fsi_model()
state0, control, prop = bench.config1_args(fm)
fm.set_prop(prop); fm.set_ini_state(state0); fm.push_to_device()
for _ in range(3):
    fm.engine.integrate(np.full(99, 1e-4), np.array([[[8e3], [0.0]]]), None, False, True)
torch.cuda.synchronize()
print('ok', fm.engine.download('info')[8:16])
