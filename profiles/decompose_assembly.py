"""Phase decomposition of asm_tile2_kernel by skipping phases (VF_DEBUG_SKIP bit mask:
1 = phase 1 records, 2 = phase 2 rows, 4 = write-out).  Measurement aid only."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'vf-fem_b200')]
import bench
model = bench.build_big_model(7, 0)
eng = model.engine
model._push_all()
for name, (res, jac) in {'res+jac': (True, True), 'jac': (False, True), 'res': (True, False)}.items():
    for skip in (0, 1, 2, 4, 3, 6, 7):
        os.environ['VF_DEBUG_SKIP'] = str(skip)
        ms = bench.time_events(lambda: eng.assemble(0, res, jac, model.dt), 20, 3) / 20
        print(json.dumps({'mode': name, 'skip_mask': skip, 'ms': round(ms, 4)}), flush=True)
