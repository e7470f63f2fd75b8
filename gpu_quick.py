# scratch driver for the first GPU bring-up
import sys, time
sys.path.insert(0, 'vf-fem_b200'); sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np, torch
from test_gpu_forward import build_fsi, benchmark_setup, oracle_run
from femvf_b200 import forward
model = build_fsi('m5')
state0, control, prop = benchmark_setup(model)
times = 1e-4 * np.arange(100)
t0 = time.time()
fin, info = forward.integrate(model, None, state0, [control], prop, times, write=False)
torch.cuda.synchronize(); t1 = time.time()
print('gpu integrate 99 steps', t1 - t0, info)
t0 = time.time()
fin, info = forward.integrate(model, None, state0, [control], prop, times, write=False)
torch.cuda.synchronize(); t1 = time.time()
print('gpu integrate 99 steps (2nd)', t1 - t0, info)
print('gmres info', model.engine.download('info'))
t0 = time.time()
hist, infos = oracle_run(model, state0, control, prop, times)
print('oracle', time.time() - t0, infos[-1]['num_iter'])
for k, key in enumerate(('u', 'v', 'a', 'q', 'p')):
    ref = hist[-1][k]
    print(key, np.max(np.abs(fin[key] - ref)) / max(np.max(np.abs(ref)), 1e-300))
