"""Cycle probes of asm_fan_pipe_kernel (library built with -DVF_PIPE_PROF, see ab_builds.sh):
    VF_LIB_PATH=/path/to/prof/libvffem_b200.so python profiles/pipe_cycles.py [levels]
Prints, per CTA, where the producer warp and the consumer warps spend their cycles."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'vf-fem_b200')]
import numpy as np, torch
import bench
levels = int(sys.argv[1]) if len(sys.argv) > 1 else 7
model = bench.build_big_model(levels, 0)
eng = model.engine
model._push_all()
for mode, (r, j) in (('res+jac', (True, True)), ('jac', (False, True)), ('res', (True, False))):
    eng.assemble(0, r, j, model.dt)
    eng.upload('info', np.zeros(16))
    reps = 5
    for _ in range(reps):
        eng.assemble(0, r, j, model.dt)
    torch.cuda.synchronize()
    info = eng.download('info')
    nwarp = 148 * (12 if os.environ.get('VF_PIPE_GROUPS', '3') == '3' else 8)
    c = info[12:16] / (reps * nwarp)
    print(f"{mode}: consumer warp total {c[0]:.0f} wait-ready {c[1]:.0f} wait-slice {c[2]:.0f} walk {c[3]:.0f}")
