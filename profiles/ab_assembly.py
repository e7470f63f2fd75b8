"""Round-2 A/B sweep of the 2D assembly kernels on one box (run under gpurun):
    python profiles/ab_assembly.py [levels]
One mesh, one set of tables; the engine is re-created per configuration (the tunables are
environment variables read at engine creation / launch).  Every configuration is checked
entry-wise against the first one (the round-1 two-phase tile kernel) before it is timed."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'vf-fem_b200')]
import numpy as np, torch
import bench
from femvf_b200.engine import Engine

levels = int(sys.argv[1]) if len(sys.argv) > 1 else 7
model = bench.build_big_model(levels, 0)
T = model.assembly_tables
TUNABLES = ('VF_FAN', 'VF_FAN_NODES', 'VF_FAN_MINB', 'VF_PF_DIST', 'VF_FAN_PIPE', 'VF_PIPE_GROUPS',
            'VF_PIPE_POOL_KB', 'VF_PIPE_PF', 'VF_PIPE_GRID')
configs = [dict(VF_FAN='0'), dict(VF_FAN_PIPE='0'), dict()]
configs += [dict(VF_PIPE_PF=pf) for pf in ('0', '1', '4')]
configs += [dict(VF_PIPE_GROUPS='2'), dict(VF_PIPE_POOL_KB='100'), dict(VF_PIPE_POOL_KB='70')]
ref = None
for cfg in configs:
    for k in TUNABLES:
        os.environ.pop(k, None)
    os.environ.update(cfg)
    try:
        eng = Engine(T)
        model._engine = eng
        model._push_all()
        out = dict(cfg)
        eng.assemble(0, True, True, model.dt)
        J = eng.view('J').clone(); F = eng.view('F').clone()
        if ref is None:
            ref = (J, F)
        else:
            out['dJ'] = float(((J - ref[0]).abs().max() / ref[0].abs().max()).item())
            out['dF'] = float(((F - ref[1]).abs().max() / ref[1].abs().max()).item())
        del J, F
        for name, (r, j) in (('ms', (True, True)), ('ms_jac', (False, True)), ('ms_res', (True, False))):
            fn = lambda: eng.assemble(0, r, j, model.dt)
            out[name] = round(bench.time_events(fn, 20, 3) / 20, 4)
        B = bench.assembly_bytes(2, eng.nn, eng.ne, eng.nnz)
        out['frac'] = round(B / out['ms'] / 1e6 / 6560, 4)
        print(json.dumps(out), flush=True)
        del eng
        model._engine = None
        torch.cuda.empty_cache()
    except Exception as ex:
        print(cfg, 'FAILED', repr(ex), flush=True)
