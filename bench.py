#!/usr/bin/env python
"""
bench.py -- residual + Jacobian assembly throughput (DOF/s) of the B200 hot path.

Contract (one JSON line on stdout, rank 0):
  python bench.py --gpus N --steps K --warmup W [--impl reference]
  N > 1: launched by torchrun, one rank per GPU; every rank assembles its own mesh shard of
  the same size (owner-computes rows with duplicated ghost cells need no collective), so
  scaling is weak and `value` is the aggregate DOF/s over all ranks.

Workload (config.workload): BASELINE.json configs[2] -- residual + Jacobian assembly
microbenchmark on the M5_CB outline red-refined 7x (4.0e6 P1 triangles, 4.0e6 DOF,
5.6e7 non-zeros).  The reference's elements are P1 only; the P2 variant named
in BASELINE.json has no reference behaviour: it is measured as written (1.0 M P2 triangles) in
the `p2` object of the same line.  One step = one assembly of F_u and J_uu (Dirichlet rows
applied) from state/properties resident in HBM.

Also measured in the same run and attached to the line: the SpMV kernel's roofline
(`spmv`), the single-simulation forward step rate on config 1 (`forward`), the 1024-member
ensemble of config 4 through the device-resident time loop and through the host-buffer C ABI
(`ensemble`, strong-scaled over the ranks), forward.integrate on a refined mesh through the
whole-GPU Newton (`forward_refined`), the P2 triangle assembly (`p2`, both kernel versions in
`ms_per_step_by_kernel`), the mesh-partitioned tetrahedral assembly + GMRES of config 5 on the
same ranks (`partition`, both node kernels in `assembly_ms_per_step_by_kernel`, and at N > 1
the check that the N-rank iterate equals the single-rank one), and the CPU oracle timed on a
bounded sample on all host cores (`cpu_baseline`).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, 'vf-fem_b200'), os.path.join(ROOT, 'tests')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = 'jacobian_assembly_dof_per_s'
UNIT = 'DOF/s'
REFINE_LEVELS = 7
SAMPLE_LEVELS = 4  # CPU baseline sample: the same outline refined 4x (1/64 of the elements)
BASE_H = 0.05


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits',
                 '-lms', '100', '-i', str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': float(np.max(smax)) if smax else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# --- workloads ----------------------------------------------------------------------------

def build_big_model(levels, seed):
    """config 3 (P1): refined M5_CB solid with random state and per-cell properties."""
    from femvf_b200 import meshgen
    from femvf_b200.models import transient
    from femvf_b200.residuals import solid as slr
    mt = meshgen.m5_cb_refined(BASE_H, levels)
    model = transient.FenicsModel(slr.KelvinVoigt(*mt))
    rng = np.random.default_rng(seed)
    N = model.state0['u'].size
    ne = model.prop['emod'].size
    nn = N // 2
    prop = model.prop.copy()
    prop['emod'][:] = rng.uniform(2.5e4, 1e5, ne)
    prop['eta'][:] = rng.uniform(1, 5, ne)
    prop['rho'][:] = 1.0
    prop['nu'][:] = 0.45
    model.set_prop(prop)
    s1 = model.state1.copy()
    s1['u'][:] = rng.uniform(-1e-3, 1e-3, N)
    s0 = model.state0.copy()
    s0['u'][:] = rng.uniform(-1e-3, 1e-3, N)
    s0['v'][:] = rng.uniform(-1e-2, 1e-2, N)
    s0['a'][:] = rng.uniform(-1e2, 1e2, N)
    model.set_ini_state(s0)
    model.set_fin_state(s1)
    ctl = model.control.copy()
    ctl['p'][:] = rng.uniform(0, 8e3, nn)
    model.set_control(ctl)
    model.dt = 1e-4
    return model


def assembly_bytes(d, nn, ne, nnz):
    """Algorithmic bytes of one residual + Jacobian assembly."""
    N = d * nn
    nen = d + 1
    return 8 * nnz + 8 * N + 32 * N + 8 * d * nn + (4 * nen + 24) * ne + 8 * nn


def spmv_bytes(nnz, nrows):
    return 12 * nnz + 20 * nrows


def time_events(fn, steps, warmup, barrier=None):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if barrier:
        barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if barrier:
        barrier()
    return e0.elapsed_time(e1)  # ms


def fsi_model(h=BASE_H):
    from femvf_b200 import meshgen
    from femvf_b200.load import load_fsi_model
    from femvf_b200.residuals import solid as slr, fluid as flr
    mt = meshgen.m5_cb_mesh(h)
    return load_fsi_model(mt, slr.KelvinVoigt, flr.BernoulliAreaRatioSep,
                          {'dirichlet_bcs': {'state/u1': [(np.zeros(2), 'facet', 'fixed')]}}, {})


def config1_args(model):
    """benchmarks/setup.py:34-49."""
    state0 = model.state0.copy(); state0[:] = 0
    control = model.control.copy(); control[:] = 0; control['psub'][:] = 8e3
    prop = model.prop.copy()
    ymax = model.solid.residual.mesh().coordinates()[:, 1].max()
    prop['emod'][:] = 5e4; prop['rho'][:] = 1; prop['eta'][:] = 3; prop['nu'][:] = 0.45
    prop['ycontact'][:] = ymax + 0.05; prop['kcontact'][:] = 1e8; prop['ymid'][:] = 1.0
    return state0, control, prop


def oracle_for(model):
    from oracle import fem, model as om
    res = model.solid.residual if hasattr(model, 'solid') else model.residual
    mesh = res.mesh()
    fids, pf_cell, _ = res.pressure_facets()
    prob = fem.SolidProblem(mesh.coordinates(), mesh.cells(), mesh.facets[fids], pf_cell,
                            res.fixed_dofs())
    return prob, om.SolidOracle(prob)


def cpu_assembly_sample(levels, steps):
    """Oracle residual + Jacobian assembly on the sample mesh; returns (DOF/s, description)."""
    model = build_big_model(levels, seed=0)
    prob, so = oracle_for(model)
    prop = {k: np.array(v) for k, v in model.prop.items()}
    u1 = model.state1['u']; st0 = tuple(model.state0.vecs); p1 = model.control['p']
    so.res(u1, st0, model.dt, prop, p1)  # warm caches
    t0 = time.perf_counter()
    for _ in range(steps):
        so.res(u1, st0, model.dt, prop, p1)
        so.jac(u1, model.dt, prop, p1)
    dt = time.perf_counter() - t0
    desc = (f"M5_CB refined {levels}x ({prob.ne} P1 triangles, {prob.N} DOF = 1/"
            f"{4 ** (REFINE_LEVELS - levels)} of the workload), {steps} residual+Jacobian "
            "assemblies, numpy/scipy oracle (CPU restatement of the FEniCS path)")
    return prob.N * steps / dt, desc, dt / steps


def _reference_worker(levels, steps, barrier, queue):
    """One host process of the reference arm: own model, same sample, timed after a barrier."""
    os.environ['OMP_NUM_THREADS'] = '1'
    model = build_big_model(levels, seed=0)
    prob, so = oracle_for(model)
    prop = {k: np.array(v) for k, v in model.prop.items()}
    u1 = model.state1['u']; st0 = tuple(model.state0.vecs); p1 = model.control['p']
    so.res(u1, st0, model.dt, prop, p1)
    so.jac(u1, model.dt, prop, p1)
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(steps):
        so.res(u1, st0, model.dt, prop, p1)
        so.jac(u1, model.dt, prop, p1)
    queue.put((prob.N, prob.ne, time.perf_counter() - t0))


def reference_rate(steps, procs=None):
    """The CPU restatement (oracle) on ALL host cores: one process per core (the numpy/scipy
    oracle is single-threaded), each assembling the bounded sample, started together.
    Returns (DOF/s = total DOF assembled / slowest process time, seconds of the slowest process,
    processes, description)."""
    import multiprocessing as mp
    procs = procs or int(os.environ.get('VF_REF_PROCS', '0')) or min(os.cpu_count() or 1, 64)
    ctx = mp.get_context('fork')
    barrier = ctx.Barrier(procs)
    queue = ctx.Queue()
    workers = [ctx.Process(target=_reference_worker, args=(SAMPLE_LEVELS, steps, barrier, queue))
               for _ in range(procs)]
    for w in workers:
        w.start()
    results = [queue.get() for _ in workers]
    for w in workers:
        w.join()
    N, ne = results[0][0], results[0][1]
    slowest = max(r[2] for r in results)
    value = N * steps * procs / slowest
    desc = (f"M5_CB refined {SAMPLE_LEVELS}x ({ne} P1 triangles, {N} DOF = 1/"
            f"{4 ** (REFINE_LEVELS - SAMPLE_LEVELS)} of the workload) per process, {steps} "
            f"residual+Jacobian assemblies each, {procs} processes started together, "
            "numpy/scipy oracle (CPU restatement of the FEniCS path)")
    return value, slowest, procs, desc


def run_reference(args, rank, world):
    """Reference arm: the reference's FEniCS/PETSc path cannot be installed, so
    the CPU restatement (oracle) is timed for the same metric on all host cores."""
    if rank != 0:
        return
    steps = min(max(args.steps, 1), 10)
    value, slowest, procs, desc = reference_rate(steps)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': steps, 'warmup': 1,
        'ms_per_step': slowest / steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': procs, 'kind': 'port',
                         'sample': desc},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def workload_config():
    return {
        'workload': ('BASELINE configs[2]: residual+Jacobian assembly on M5_CB red-refined '
                     f'{REFINE_LEVELS}x, P1 triangles (reference elements are P1 only)'),
        'element': 'P1 triangle, plane strain, Kelvin-Voigt + Newmark + follower pressure',
        'l2_policy': 'inputs+outputs (>= 0.6 GB per step) exceed the 126 MB L2; no explicit flush',
    }


def run_partition(args, rank, world, local_rank, standalone=True, tets=None):
    """BASELINE configs[4]: extruded M5 tetrahedral mesh partitioned over the ranks; local
    (communication-free) assembly and a fixed number of GMRES iterations with NCCL halo
    exchange.  ``--workload partition`` prints it as its own JSON line; the default workload
    carries it (on a smaller mesh, --partition-tets) as the ``partition`` object of its line."""
    import torch
    import torch.distributed as dist
    from femvf_b200 import meshgen
    from femvf_b200.distributed import DistributedSolid
    from femvf_b200.residuals import solid as slr
    tets = args.tets if tets is None else tets
    torch.cuda.set_device(local_rank)
    if world > 1 and standalone:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    t0 = time.perf_counter()
    mt2 = meshgen.m5_cb_refined(BASE_H, args.levels2d, renumber=False)
    nz = max(int(round(tets / (3.0 * mt2[0].num_cells()))), 1)
    mt3 = meshgen.renumber_for_locality(meshgen.extrude_to_tets(mt2, 1.5, nz))
    res = slr.KelvinVoigt(*mt3)
    ds = DistributedSolid(res, rank, world, restart=30)
    setup_s = time.perf_counter() - t0
    mesh = res.mesh()
    nn, ne = mesh.num_vertices(), mesh.num_cells()
    rng = np.random.default_rng(0)
    prop = dict(rho=np.full(ne, 1.0), eta=rng.uniform(1, 5, ne), emod=rng.uniform(2.5e4, 1e5, ne))
    N = 3 * nn
    state = dict(u1=rng.uniform(-1e-4, 1e-4, N), u0=rng.uniform(-1e-4, 1e-4, N),
                 v0=rng.uniform(-1e-2, 1e-2, N), a0=rng.uniform(-1e2, 1e2, N))
    p1 = rng.uniform(0, 8e3, nn)
    scal = np.zeros(10); scal[0] = 0.45; scal[1] = np.inf; scal[2] = 1.0; scal[4] = 1.0
    ds.upload_global(prop, state, p1, scal)
    barrier = dist.barrier if world > 1 else None
    dt = 1e-4
    ms_asm = time_events(lambda: ds.assemble(dt), args.steps, args.warmup, barrier)
    # the two thread-per-node kernels side by side (VF_NODE_WARP selects)
    asm_by_kernel = {}
    try:
        for setting, name in (('00', 'asm_node_global_kernel, generic gathers (round 1)'),
                              ('11', 'asm_node_warp_kernel + gather tables (default)')):
            os.environ['VF_NODE_WARP'], os.environ['VF_TET_TABLES'] = setting[0], setting[1]
            asm_by_kernel[name] = time_events(lambda: ds.assemble(dt), 10, 2, barrier) / 10
    except Exception as ex:
        asm_by_kernel = {'error': repr(ex)}
    finally:
        os.environ.pop('VF_NODE_WARP', None)
        os.environ.pop('VF_TET_TABLES', None)
    ds.assemble(dt)
    b = ds.owned('F').clone()
    x = torch.empty_like(b)
    iters = args.gmres_iters

    def solve():
        return ds.solve(b, x, rtol=0.0, atol=0.0, maxiter=iters)
    solve()
    ds.gmres.spmv_count = 0
    ms_solve = time_events(solve, 1, 0, barrier)
    t = torch.tensor([ms_asm, ms_solve], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    eng = ds.engine
    own_nnz = int(9 * (ds.tables['brptr'][ds.part.n_own] - ds.tables['brptr'][0]))
    line = {
        'metric': 'partitioned_assembly_dof_per_s', 'unit': 'DOF/s',
        'value': N * args.steps / (float(t[0]) * 1e-3), 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': float(t[0]) / args.steps,
        'higher_is_better': True, 'scaling': 'strong', 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'BASELINE configs[4]: M5_CB triangles extruded to tetrahedra, '
                               'vertex ranges partitioned over ranks, owner-computes assembly + '
                               'GMRES with NCCL halo exchange',
                   'tets': ne, 'nn': nn, 'dof': N, 'nz': nz,
                   'rank0': {'owned_nodes': ds.part.n_own, 'ghost_nodes': len(ds.part.ghost_global),
                             'local_cells': len(ds.part.cell_ids), 'owned_nnz': own_nnz,
                             'halo_send_bytes_per_spmv': ds.halo.halo_bytes}},
        'gmres': {'iterations': iters, 'ms_total': float(t[1]),
                  'iterations_per_s': iters / (float(t[1]) * 1e-3),
                  'operator_applications': ds.gmres.spmv_count,
                  'halo_exchanges_per_iteration': 1 if world > 1 else 0,
                  'allreduces_per_iteration': 2 if world > 1 else 0,
                  'note': 'host-launched, device-resident Arnoldi data (one device->host read per 8 '
                          'iterations); CGS2, restart 30, left block-Jacobi'},
        'setup_s': setup_s, 'gpu_launches': int(eng.launch_count),
        'assembly_ms_per_step_by_kernel': asm_by_kernel,
    }
    if world > 1:
        # the check of tests/test_gpu_partition.py::test_two_gpu_distributed_solve_matches_lu,
        # moved into the multi-GPU run itself (the GPU test box has one GPU): block-Jacobi GMRES
        # is partition-independent, so the iterate after `iters` iterations across `world` ranks
        # must equal the single-rank iterate to round-off
        pieces = [None] * world
        dist.all_gather_object(pieces, (ds.part.n0, x.cpu().numpy()))
        if rank == 0:
            ds1 = DistributedSolid(res, 0, 1, restart=30)
            ds1.upload_global(prop, state, p1, scal)
            ds1.assemble(dt)
            b1 = ds1.owned('F').clone()
            x1 = torch.empty_like(b1)
            ds1.solve(b1, x1, rtol=0.0, atol=0.0, maxiter=iters)
            x_all = np.concatenate([p[1] for p in sorted(pieces, key=lambda p: p[0])])
            x1 = x1.cpu().numpy()
            diff = float(np.linalg.norm(x_all - x1) / np.linalg.norm(x1))
            line['check'] = {'what': f'{world}-rank iterate after {iters} GMRES iterations vs the '
                                     'single-rank iterate (NCCL halo exchange + all-reduces)',
                             'rel_diff': diff, 'ok': bool(diff <= 1e-8)}
            del ds1
        dist.barrier()
    del ds
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    if not standalone:
        return line
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--levels', type=int, default=REFINE_LEVELS)
    ap.add_argument('--skip-extras', action='store_true')
    ap.add_argument('--workload', default='assembly', choices=['assembly', 'partition'])
    ap.add_argument('--tets', type=float, default=5.0e6)
    ap.add_argument('--levels2d', type=int, default=3)
    ap.add_argument('--gmres-iters', type=int, default=60)
    ap.add_argument('--partition-tets', type=float, default=1.0e6,
                    help='size of the mesh-partition run carried by the default line (0: skip)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if args.workload == 'partition':
        run_partition(args, rank, world, local_rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: femvf_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    barrier = None
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        barrier = dist.barrier
    peak, peak_src = load_peaks()

    # ---- primary: assembly on the large mesh --------------------------------------------------
    model = build_big_model(args.levels, seed=rank)
    eng = model.engine
    model._push_all()
    d, nn, ne, nnz, N = eng.dim, eng.nn, eng.ne, eng.nnz, eng.N
    launches0 = eng.launch_count

    def step():
        eng.assemble(0, res=True, jac=True, dt=model.dt)

    # clocks are sampled from before the warm-up (same kernel, same load) to the end of the timed
    # region: nvidia-smi cannot sample faster than ~100 ms and the timed region may be shorter
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_w = time.perf_counter()
    n_w = 0
    while n_w < args.warmup or time.perf_counter() - t_w < 0.6:
        step()
        n_w += 1
    torch.cuda.synchronize()
    launches0 = eng.launch_count
    ms = time_events(step, args.steps, 0, barrier)
    launches = eng.launch_count - launches0
    clocks = sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = N * world * args.steps / (ms_max * 1e-3)
    B_asm = assembly_bytes(d, nn, ne, nnz)
    asm_gbs = B_asm / (ms / args.steps * 1e-3) / 1e9

    # ---- SpMV on the same matrix ------------------------------------------------------------
    x = torch.randn(N, dtype=torch.float64, device='cuda')
    y = torch.empty_like(x)
    ms_spmv = time_events(lambda: eng.spmv(x, y), max(args.steps, 20), 3)
    n_spmv = max(args.steps, 20)
    B_spmv = spmv_bytes(nnz, N)
    spmv_gbs = B_spmv / (ms_spmv / n_spmv * 1e-3) / 1e9

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_max / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': dict(workload_config(), nn=nn, ne=ne, dof=N, nnz=nnz,
                       parallelism=f'{world} independent mesh shards, no collective'),
        'roofline': {'kernel': 'asm_fan_pipe_kernel<true,true,3,false> (+ facet_bc_fast_kernel on boundary nodes)',
                     'bound': 'hbm',
                     'achieved': asm_gbs, 'peak': peak, 'unit': 'GB/s', 'frac': asm_gbs / peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel on
                     # this workload, from the ncu --set full capture summarised in
                     # profiles/r2_ncu_full_asm_fan_pipe.csv (not re-measured in this run)
                     'traffic': 792462080 if args.levels == REFINE_LEVELS else None,
                     'algorithmic_bytes': B_asm, 'peak_source': peak_src},
        'spmv': {'kernel': 'spmv_kernel<2,4>', 'bound': 'hbm', 'achieved': spmv_gbs,
                 'peak': peak, 'unit': 'GB/s', 'frac': spmv_gbs / peak,
                 'algorithmic_bytes': B_spmv, 'ms': ms_spmv / n_spmv},
        'clocks': clocks,
        'gpu_launches': int(launches),
        'kernel_config': dict(model.engine.tile_info),
    }

    # ---- e2e: the public API with host buffers (set_fin_state -> assem_res + assem_dres_dstate1)
    if rank == 0 or world > 1:
        e2e_steps = 5
        s1 = model.state1.copy()
        # the per-Newton-iteration pattern: only u1 changes between calls
        model.trust_setters = True

        def e2e_step():
            model.set_fin_state(s1)
            r = model.assem_res()
            # dF_u/du1 stays on the device (femvf_b200.devmat.DeviceCSR): a Newton loop hands it
            # straight back to solve_dres_dstate1; nothing but its handle crosses PCIe
            J = model.assem_dres_dstate1().sub['u', 'state/u1']
            assert J.on_device
            return r, J
        e2e_step()
        torch.cuda.synchronize()
        if barrier:
            barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt_e2e = time.perf_counter() - t0
        te = torch.tensor([dt_e2e], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        line['e2e'] = {'value': N * world * e2e_steps / float(te.item()), 'unit': UNIT,
                       'h2d_bytes_per_step': 8 * 3 * N, 'd2h_bytes_per_step': 8 * 3 * N,
                       'api': 'FenicsModel.set_fin_state + assem_res + assem_dres_dstate1 '
                              '(host BlockVector in; host F_u, F_v, F_a out; dF_u/du1 returned '
                              'as a device-resident DeviceCSR, downloaded only if the caller '
                              'reads its values; trust_setters=True: only the state changed '
                              'through the setter is re-uploaded)'}

    # release the big model now (engine arena, page-locked staging buffers): left to the
    # cyclic garbage collector it would be freed at a random point inside a later timed loop
    del model, eng, x, y
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    torch.cuda.synchronize()

    # ---- forward.integrate on config 1 and the ensemble of config 4 ------------------------------
    if not args.skip_extras:
        from femvf_b200 import forward
        fm = fsi_model()
        state0, control, prop = config1_args(fm)
        times = 1e-4 * np.arange(100)
        for _ in range(2):  # warm-up (engine creation, pinned buffers, kernel load)
            forward.integrate(fm, None, state0, [control], prop, times, write=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            forward.integrate(fm, None, state0, [control], prop, times, write=False)
        torch.cuda.synchronize()
        fwd = 99 * reps / (time.perf_counter() - t0)
        line['forward'] = {'workload': 'BASELINE configs[0]: M5_CB (245 P1 triangles, 296 DOF), '
                                       '99 steps dt=1e-4, forward.integrate(write=False)',
                           'steps_per_s': fwd}

        # ensemble (SURVEY.md 8d cfg 4): 1024 members IN TOTAL, randomised emod / eta fields,
        # sharded contiguously over the ranks (strong scaling, no data-path collective); the
        # weak-scaled variant (1024 members per GPU) is reported beside it
        from femvf_b200.ensemble import EnsembleRunner, shard_members
        dts = np.full(99, 1e-4)
        ctl = np.array([[[8e3], [0.0]]])

        def run_ensemble(first, count):
            runner = EnsembleRunner(fm, count)
            emod = np.empty((count, runner.ne)); eta = np.empty((count, runner.ne))
            for b in range(count):
                g = np.random.default_rng([4, first + b])
                emod[b] = 5e4 * np.exp(0.3 * g.standard_normal(runner.ne))
                eta[b] = 3.0 * np.exp(0.3 * g.standard_normal(runner.ne))
            ini = np.zeros((count, runner.state_size))
            runner.set_common_prop(prop)
            runner.run_host(dts, ctl, ini, emod, eta)  # warm-up
            if barrier:
                barrier()
            dt_hs = []
            for _ in range(3):          # median of three: a single host-timed run is noisy
                t0 = time.perf_counter()
                fin, series = runner.run_host(dts, ctl, ini, emod, eta)
                dt_hs.append(time.perf_counter() - t0)
            # device-resident timing of the same work
            ms_devs = []
            for _ in range(3):
                runner.upload_members(ini, emod, eta)   # every repetition starts from the same state
                ms_devs.append(time_events(lambda: runner.run_device(dts, ctl), 1, 0, barrier))
            tt = torch.tensor([sorted(ms_devs)[1], sorted(dt_hs)[1] * 1e3], dtype=torch.float64,
                              device='cuda')
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            out = {'ms_device': float(tt[0]), 'ms_host_buffers': float(tt[1]),
                   'h2d_bytes': int(ini.nbytes + emod.nbytes + eta.nbytes),
                   'd2h_bytes': int(fin.nbytes + series.nbytes),
                   'max_newton_iters': float(series[:, :, 0].max()),
                   'e2e_run_seconds': [round(x, 4) for x in dt_hs]}
            del runner
            return out
        B = 1024
        lo, hi = shard_members(B, rank, world)
        strong = run_ensemble(lo, hi - lo)
        line['ensemble'] = {
            'workload': f'BASELINE configs[3]: {B} members in total x 99 steps, emod/eta '
                        f'log-normal per cell, seed = member id; {hi - lo} members on rank 0',
            'scaling': 'strong',
            'member_steps_per_s': B * 99 / (strong['ms_device'] * 1e-3),
            'e2e_member_steps_per_s': B * 99 / (strong['ms_host_buffers'] * 1e-3),
            'h2d_bytes': strong['h2d_bytes'], 'd2h_bytes': strong['d2h_bytes'],
            'max_newton_iters': strong['max_newton_iters'],
            'e2e_run_seconds': strong['e2e_run_seconds'],
        }
        if world > 1:
            weak = run_ensemble(rank * B, B)
            line['ensemble']['weak'] = {
                'workload': f'{B} members PER GPU',
                'member_steps_per_s': B * world * 99 / (weak['ms_device'] * 1e-3),
                'e2e_member_steps_per_s': B * world * 99 / (weak['ms_host_buffers'] * 1e-3)}

    # ---- forward.integrate on a REFINED mesh: the coupled model on a solid that does not fit one
    # CTA takes the per-step path (whole-GPU Newton: pipelined assembly + ILU(0)-GMRES) ------------
    if not args.skip_extras and rank == 0:
        try:
            from femvf_b200 import forward, meshgen
            from femvf_b200.load import load_fsi_model
            from femvf_b200.residuals import solid as slr, fluid as flr
            lv = 4
            fr = load_fsi_model(meshgen.m5_cb_refined(BASE_H, lv), slr.KelvinVoigt,
                                flr.BernoulliAreaRatioSep,
                                {'dirichlet_bcs': {'state/u1': [(np.zeros(2), 'facet', 'fixed')]}}, {})
            st0, ctl_r, prop_r = config1_args(fr)
            nst = 10
            tms = 1e-4 * np.arange(nst + 1)
            forward.integrate(fr, None, st0, [ctl_r], prop_r, tms[:3], write=False)   # warm-up
            gs = fr.solid._grid_solver()
            it0 = gs.gmres.spmv_count
            l0 = fr.engine.launch_count
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            forward.integrate(fr, None, st0, [ctl_r], prop_r, tms, write=False)
            torch.cuda.synchronize()
            dt_r = time.perf_counter() - t0
            e_r = fr.engine
            ops = gs.gmres.spmv_count - it0
            # bytes of the dominant kernels: one SpMV + one ILU(0) application (~1.3 SpMV) per
            # operator application, the Krylov basis read twice (CGS2); assemblies are a few per step
            B_op = 2.3 * spmv_bytes(e_r.nnz, e_r.N) + 8 * e_r.N * (gs.gmres.m + 4)
            line['forward_refined'] = {
                'workload': f'M5_CB refined {lv}x coupled to the Bernoulli fluid: {e_r.ne} P1 '
                            f'triangles, {e_r.N} DOF, {nst} steps dt=1e-4, forward.integrate '
                            '(per-step API, whole-GPU Newton with ILU(0)-GMRES(40))',
                'steps_per_s': nst / dt_r, 'ms_per_step': 1e3 * dt_r / nst,
                'operator_applications_per_step': ops / nst,
                'gpu_launches_per_step': (e_r.launch_count - l0) / nst,
                'roofline': {'bound': 'hbm', 'achieved': ops * B_op / dt_r / 1e9, 'peak': peak,
                             'unit': 'GB/s', 'frac': ops * B_op / dt_r / 1e9 / peak,
                             'bytes_model': '2.3 B_spmv + 8 N (m + 4) per operator application'}}
            del fr, gs
            torch.cuda.empty_cache()
        except Exception as ex:
            line['forward_refined'] = {'error': repr(ex)}

    # ---- configs[2] as written: ~1 M P2 triangles (the reference itself is P1 only; P2 is the
    # extension BASELINE.json names).  First P2 kernel: node-gather, not yet tiled / pipelined ----
    if not args.skip_extras and args.levels >= REFINE_LEVELS and rank == 0:
        try:
            from femvf_b200 import meshgen
            from femvf_b200.p2 import P2Assembler
            mt = meshgen.m5_cb_refined(BASE_H, REFINE_LEVELS - 1)
            mesh2 = mt[0]
            asm = P2Assembler(mesh2.coordinates(), mesh2.cells())
            g = torch.Generator(device='cuda'); g.manual_seed(0)
            rnd = lambda n, lo, hi: lo + (hi - lo) * torch.rand(n, dtype=torch.float64,
                                                                device='cuda', generator=g)
            Np, nnp, nep = asm.N, asm.nn, asm.ne
            vec = [rnd(Np, -1e-3, 1e-3), rnd(Np, -1e-3, 1e-3), rnd(Np, -1e-2, 1e-2),
                   rnd(Np, -1e2, 1e2), rnd(nnp, 0.0, 8e3), rnd(nep, 2.5e4, 1e5), rnd(nep, 1, 5),
                   torch.ones(nep, dtype=torch.float64, device='cuda')]
            fn = lambda: asm.assemble(*vec, 0.45, 1e-4)
            ms_p2 = time_events(fn, 20, 3) / 20
            B_p2 = 8 * asm.nnz + 8 * Np + 32 * Np + 16 * nnp + (4 * 6 + 24) * nep + 8 * nnp
            line['p2'] = {
                'workload': f'BASELINE configs[2]: residual+Jacobian assembly, {nep} P2 triangles '
                            f'(M5_CB refined {REFINE_LEVELS - 1}x), {Np} DOF, {asm.nnz} non-zeros',
                'kernel': 'p2_assemble_warp_kernel (node gather, closed-form reference tensors, '
                          'structural zeros skipped, coalesced row write-out)',
                'ms_per_step': ms_p2, 'value': Np / (ms_p2 * 1e-3), 'unit': UNIT,
                'roofline': {'bound': 'hbm', 'algorithmic_bytes': B_p2,
                             'achieved': B_p2 / (ms_p2 * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                             'frac': B_p2 / (ms_p2 * 1e-3) / 1e9 / peak}}
            # both versions of the kernel side by side (VF_P2_WARP selects)
            try:
                both = {}
                for setting, name in (('0', 'p2_assemble_kernel'),
                                      ('1', 'p2_assemble_warp_kernel (default)')):
                    os.environ['VF_P2_WARP'] = setting
                    both[name] = time_events(fn, 10, 2) / 10
                line['p2']['ms_per_step_by_kernel'] = both
            except Exception as ex:
                line['p2']['ms_per_step_by_kernel'] = {'error': repr(ex)}
            finally:
                os.environ.pop('VF_P2_WARP', None)
            del asm, vec
            torch.cuda.empty_cache()
        except Exception as ex:
            line['p2'] = {'error': repr(ex)}

    # ---- mesh partition (BASELINE configs[4]) on the same ranks: owner-computes assembly and a
    # fixed number of GMRES iterations with halo exchange + all-reduces over NCCL -----------------
    if args.partition_tets > 0 and not args.skip_extras:
        try:
            line['partition'] = run_partition(args, rank, world, local_rank, standalone=False,
                                              tets=args.partition_tets)
            line['config']['parallelism'] = (
                f'assembly: {world} mesh shards, no collective; ensemble: members sharded over '
                f'{world} ranks, no collective; partition: {world} vertex-range partitions, NCCL '
                'halo exchange + all-reduce per GMRES iteration')
        except Exception as ex:  # keep the primary line if the extra fails
            line['partition'] = {'error': repr(ex)}

    # ---- CPU baseline (rank 0, N = 1 only): the oracle on a bounded sample, all host cores
    # (the same measurement as the reference arm, --impl reference) -------------------------------
    if rank == 0 and world == 1:
        cpu_val, _, procs, desc = reference_rate(5)
        line['cpu_baseline'] = {'value': cpu_val, 'unit': UNIT, 'cores': procs, 'kind': 'port',
                                'sample': desc}
        if not args.skip_extras:
            from oracle import model as om
            prob, so = oracle_for(fm)
            co = om.CoupledOracle(so, fm.fluid.residual.mesh(), fm.fsimap.dofs_solid,
                                  fm.fsimap.dofs_fluid)
            t0 = time.perf_counter()
            co.integrate(tuple(state0.vecs), [{'psub': control['psub'], 'psup': control['psup']}],
                         {k: np.array(v) for k, v in prop.items()}, times)
            line['forward']['cpu_oracle_steps_per_s'] = 99 / (time.perf_counter() - t0)

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
