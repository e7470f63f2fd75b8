"""A/B timing of the kernel variants behind VF_NODE_WARP (tetrahedra: block rows in global memory
vs in a warp's shared-memory slice) and VF_P2_WARP (P2 triangles: version 1 / 2), CUDA events:
    python profiles/variants_ab.py [tets] [p2_levels] [reps]
Prints one JSON line per workload (also checks that the two settings give the same numbers)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'vf-fem_b200')]
import numpy as np
import torch
import bench

tets = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0e6
p2_levels = int(sys.argv[2]) if len(sys.argv) > 2 else 6
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
PEAK = 6560.0


def tet_case():
    from femvf_b200 import meshgen
    from femvf_b200.models import transient
    from femvf_b200.residuals import solid as slr
    t0 = time.perf_counter()
    mt2 = meshgen.m5_cb_refined(bench.BASE_H, 3, renumber=False)
    nz = max(int(round(tets / (3.0 * mt2[0].num_cells()))), 1)
    mt3 = meshgen.renumber_for_locality(meshgen.extrude_to_tets(mt2, 1.5, nz))
    model = transient.FenicsModel(slr.KelvinVoigt(*mt3))
    rng = np.random.default_rng(0)
    N, ne = model.state0['u'].size, model.prop['emod'].size
    prop = model.prop.copy()
    prop['emod'][:] = rng.uniform(2.5e4, 1e5, ne); prop['eta'][:] = rng.uniform(1, 5, ne)
    prop['rho'][:] = 1.0
    model.set_prop(prop)
    s1 = model.state1.copy(); s1['u'][:] = rng.uniform(-1e-4, 1e-4, N)
    s0 = model.state0.copy(); s0['u'][:] = rng.uniform(-1e-4, 1e-4, N)
    s0['v'][:] = rng.uniform(-1e-2, 1e-2, N); s0['a'][:] = rng.uniform(-1e2, 1e2, N)
    model.set_ini_state(s0); model.set_fin_state(s1)
    ctl = model.control.copy(); ctl['p'][:] = rng.uniform(0, 8e3, N // 3)
    model.set_control(ctl)
    model.dt = 1e-4
    eng = model.engine
    model._push_all()
    out = {'workload': f'{ne} tetrahedra, {N} DOF, {eng.nnz} non-zeros', 'setup_s': time.perf_counter() - t0}
    B = bench.assembly_bytes(3, eng.nn, eng.ne, eng.nnz)
    ref = None
    for setting, name in (('00', 'asm_node_global_kernel'), ('10', 'asm_node_warp_kernel'),
                          ('01', 'asm_node_global_kernel + gather tables'),
                          ('11', 'asm_node_warp_kernel + gather tables')):
        os.environ['VF_NODE_WARP'] = setting[0]
        os.environ['VF_TET_TABLES'] = setting[1]
        fn = lambda: eng.assemble(0, True, True, model.dt)
        ms = bench.time_events(fn, reps, 3) / reps
        msj = bench.time_events(lambda: eng.assemble(0, False, True, model.dt), reps, 1) / reps
        J = eng.view('J').clone()
        if ref is None:
            ref = J
        else:
            out['max_rel_diff'] = max(out.get('max_rel_diff', 0.0),
                                      float((J - ref).abs().max() / ref.abs().max()))
        out[name] = {'ms': round(ms, 4), 'ms_jac_only': round(msj, 4), 'hbm_frac': round(B / ms / 1e6 / PEAK, 4)}
    os.environ.pop('VF_NODE_WARP')
    os.environ.pop('VF_TET_TABLES')
    print(json.dumps(out), flush=True)


def p2_case():
    from femvf_b200 import meshgen
    from femvf_b200.p2 import P2Assembler
    t0 = time.perf_counter()
    mesh2 = meshgen.m5_cb_refined(bench.BASE_H, p2_levels)[0]
    asm = P2Assembler(mesh2.coordinates(), mesh2.cells())
    g = torch.Generator(device='cuda'); g.manual_seed(0)
    rnd = lambda n, lo, hi: lo + (hi - lo) * torch.rand(n, dtype=torch.float64, device='cuda', generator=g)
    Np, nnp, nep = asm.N, asm.nn, asm.ne
    vec = [rnd(Np, -1e-3, 1e-3), rnd(Np, -1e-3, 1e-3), rnd(Np, -1e-2, 1e-2), rnd(Np, -1e2, 1e2),
           rnd(nnp, 0.0, 8e3), rnd(nep, 2.5e4, 1e5), rnd(nep, 1, 5),
           torch.ones(nep, dtype=torch.float64, device='cuda')]
    B = 8 * asm.nnz + 8 * Np + 32 * Np + 16 * nnp + (4 * 6 + 24) * nep + 8 * nnp
    out = {'workload': f'{nep} P2 triangles, {Np} DOF, {asm.nnz} non-zeros', 'setup_s': time.perf_counter() - t0}
    ref = None
    for setting, name in (('00', 'p2_assemble_kernel'), ('10', 'p2_assemble_warp_kernel'),
                          ('11', 'p2_pack_state_kernel + p2_assemble_warp_kernel')):
        os.environ['VF_P2_WARP'] = setting[0]
        os.environ['VF_P2_PACK'] = setting[1]
        ms = bench.time_events(lambda: asm.assemble(*vec, 0.45, 1e-4), reps, 3) / reps
        J = asm.J.clone()
        if ref is None:
            ref = J
        else:
            out['max_rel_diff'] = max(out.get('max_rel_diff', 0.0),
                                      float((J - ref).abs().max() / ref.abs().max()))
        out[name] = {'ms': round(ms, 4), 'hbm_frac': round(B / ms / 1e6 / PEAK, 4)}
    os.environ.pop('VF_P2_WARP')
    os.environ.pop('VF_P2_PACK')
    print(json.dumps(out), flush=True)


if __name__ == '__main__':
    for case in (tet_case, p2_case):
        try:
            case()
        except Exception as ex:  # keep going: the other workload is independent
            print(json.dumps({'case': case.__name__, 'error': repr(ex)}), flush=True)
        torch.cuda.empty_cache()
