"""Tuning sweep for the tile assembly kernels (run on the GPU box):
python profiles/sweep_assembly.py [levels]"""
import itertools, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'vf-fem_b200')]
import numpy as np, torch
import bench
from femvf_b200.engine import Engine

levels = int(sys.argv[1]) if len(sys.argv) > 1 else 7
model = bench.build_big_model(levels, 0)
T = model.assembly_tables
configs = []
for tn in (40, 48, 56, 64, 72, 80, 88, 96, 104, 112, 128):
    configs.append(dict(VF_TILE2='1', VF_TILE_NODES=str(tn)))
configs.append(dict(VF_TILE2='0', VF_TILE_NODES='128'))
for cfg in configs:
    os.environ.update(cfg)
    try:
        eng = Engine(T)
        model._engine = eng
        model._push_all()
        fn = lambda: eng.assemble(0, True, True, model.dt)
        ms = bench.time_events(fn, 20, 3) / 20
        B = bench.assembly_bytes(2, eng.nn, eng.ne, eng.nnz)
        print(json.dumps(dict(cfg, ms=round(ms, 4), frac=round(B / ms / 1e6 / 6560, 4), **eng.tile_info)), flush=True)
        del eng
        model._engine = None
        torch.cuda.empty_cache()
    except Exception as ex:
        print(cfg, 'FAILED', ex, flush=True)
