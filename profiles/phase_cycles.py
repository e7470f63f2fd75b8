import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [ROOT, os.path.join(ROOT, 'vf-fem_b200'), os.path.join(ROOT, 'tests')]
import numpy as np, torch
import bench
os.environ['VF_DEBUG_SKIP'] = '32'
model = bench.build_big_model(7, 0)
eng = model.engine if hasattr(model, 'engine') else model._engine
model._push_all()
eng = model._engine
eng.upload('info', np.zeros(16))
for _ in range(5): eng.assemble(0, True, True, model.dt)
torch.cuda.synchronize()
info = eng.download('info')
n = info[12]
print('staging split: desc %.0f nodes %.0f props+cpasync %.0f' % tuple(info[13:16] / n))
print('CTAs', n, 'cycles/CTA: staging %.0f phase1 %.0f phase2 %.0f writeout %.0f total %.0f' % (*(info[8:12] / n), info[8:12].sum() / n))
