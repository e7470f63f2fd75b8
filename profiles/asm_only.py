"""Residual + Jacobian assembly of the bench mesh, nothing else (the ncu target of round 2):
    python profiles/asm_only.py [levels] [reps]
Prints the CUDA-event time per assembly; tunables come from the environment."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'vf-fem_b200')]
import bench

levels = int(sys.argv[1]) if len(sys.argv) > 1 else 7
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
model = bench.build_big_model(levels, 0)
eng = model.engine
model._push_all()
out = {}
for name, (r, j) in (('ms', (True, True)), ('ms_jac', (False, True)), ('ms_res', (True, False))):
    fn = lambda: eng.assemble(0, r, j, model.dt)
    out[name] = round(bench.time_events(fn, reps, 3) / reps, 4)
B = bench.assembly_bytes(2, eng.nn, eng.ne, eng.nnz)
out['frac'] = round(B / out['ms'] / 1e6 / 6560, 4)
print(json.dumps(out), flush=True)
