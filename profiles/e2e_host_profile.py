import sys, os, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [ROOT, os.path.join(ROOT, 'vf-fem_b200'), os.path.join(ROOT, 'tests')]
import numpy as np, torch
import bench
model = bench.build_big_model(7, 0)
model.trust_setters = True
s1 = model.state1.copy()
def step():
    model.set_fin_state(s1)
    r = model.assem_res()
    J = model.assem_dres_dstate1().sub['u', 'state/u1']
    return r, J
step(); step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(3): step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
