# scratch driver: forward + ensemble timing
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [ROOT, os.path.join(ROOT, 'vf-fem_b200'), os.path.join(ROOT, 'tests')]
import numpy as np, torch
import bench
from femvf_b200 import forward
from femvf_b200.ensemble import EnsembleRunner
fm = bench.fsi_model()
state0, control, prop = bench.config1_args(fm)
times = 1e-4 * np.arange(100)
forward.integrate(fm, None, state0, [control], prop, times, write=False)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3):
    fin, info = forward.integrate(fm, None, state0, [control], prop, times, write=False)
torch.cuda.synchronize(); print('forward steps/s', 297 / (time.perf_counter() - t0), info, fm.engine.download('info')[:6])
# device-only time of the kernel
fm.push_to_device()
ms = bench.time_events(lambda: fm.engine.integrate(np.full(99, 1e-4), np.array([[[8e3], [0.0]]]), None, False, True), 3, 1) / 3
print('device-only 99 steps ms', ms, 'steps/s', 99 / ms * 1e3)
info = fm.engine.download('info'); print('cycles/step: asm %.0f spmv %.0f orth %.0f givens %.0f tail %.0f fluid %.0f total %.0f; inverse total %.0f; gmres its %s' % (*tuple(info[8:15] / 99), info[15], info[3]))
B = int(os.environ.get('B', '1024'))
runner = EnsembleRunner(fm, B)
rng = np.random.default_rng(0)
emod = 5e4 * np.exp(0.3 * rng.standard_normal((B, runner.ne))); eta = 3.0 * np.exp(0.3 * rng.standard_normal((B, runner.ne)))
ini = np.zeros((B, runner.state_size)); dts = np.full(99, 1e-4); ctl = np.array([[[8e3], [0.0]]])
runner.set_common_prop(prop)
runner.run_host(dts, ctl, ini, emod, eta)
t0 = time.perf_counter(); fin, series = runner.run_host(dts, ctl, ini, emod, eta); dt = time.perf_counter() - t0
print('ensemble e2e member-steps/s', B * 99 / dt, 'max newton', series[:, :, 0].max())
runner.upload_members(ini, emod, eta)
ms = bench.time_events(lambda: runner.run_device(dts, ctl), 1, 0)
print('ensemble device member-steps/s', B * 99 / ms * 1e3, 'ms', ms)
