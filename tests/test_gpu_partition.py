"""GPU tests of the mesh-partitioned path: owner-computes assembly without communication, the
grid-wide GMRES (world = 1) against LU, and -- when two GPUs are visible -- the NCCL halo
exchange + distributed solve."""

import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from helpers import mesh_tuples, oracle_problem
from oracle import model as om

pytestmark = pytest.mark.gpu


def problem(levels=3, dim=2):
    from femvf_b200 import meshgen
    from femvf_b200.residuals import solid as slr
    mt = meshgen.m5_cb_refined(0.05, levels) if dim == 2 else \
        meshgen.renumber_for_locality(meshgen.extrude_to_tets(meshgen.m5_cb_mesh(0.05), 1.5, 6))
    res = slr.KelvinVoigt(*mt)
    prob = oracle_problem(res)
    rng = np.random.default_rng(0)
    prop = dict(rho=np.full(prob.ne, 1.0), eta=rng.uniform(1, 5, prob.ne),
                emod=rng.uniform(2.5e4, 1e5, prob.ne), nu=0.45)
    N = prob.N
    state = dict(u1=rng.uniform(-1e-3, 1e-3, N), u0=rng.uniform(-1e-3, 1e-3, N),
                 v0=rng.uniform(-1e-2, 1e-2, N), a0=rng.uniform(-1e2, 1e2, N))
    p1 = rng.uniform(0, 8e3, prob.nn)
    scal = np.zeros(10); scal[0] = 0.45; scal[1] = np.inf; scal[2] = 1.0; scal[4] = 1.0
    return res, prob, prop, state, p1, scal


@pytest.mark.parametrize('dim', [2, 3])
def test_owner_computes_rows_match_global_assembly(dim):
    """Every rank's owned rows of J and F, assembled from its local cells only, equal the
    corresponding rows of the oracle's global matrix: no communication is needed."""
    from femvf_b200.distributed import DistributedSolid
    res, prob, prop, state, p1, scal = problem(2 if dim == 2 else 0, dim)
    dt = 1e-4
    so = om.SolidOracle(prob)
    J = so.jac(state['u1'], dt, prop, p1).tocsr()
    F = so.res(state['u1'], (state['u0'], state['v0'], state['a0']), dt, prop, p1)
    world = 3
    d = prob.d
    for r in range(world):
        ds = DistributedSolid(res, r, world)
        ds.upload_global(prop, state, p1, scal)
        ds.assemble(dt)
        p = ds.part
        Fl = ds.owned('F').cpu().numpy()
        assert np.max(np.abs(Fl - F[d * p.n0:d * p.n1])) <= 1e-12 * np.max(np.abs(F))
        rowptr, colidx = ds.engine.csr_pattern()
        vals = ds.engine.download('J')
        nloc = d * p.n_local
        Jl = sp.csr_matrix((vals, colidx, rowptr), shape=(nloc, nloc))[:d * p.n_own]
        # map local columns back to global and compare with the owned rows of J
        gcol = (d * p.local_nodes[:, None] + np.arange(d)).reshape(-1)
        Jl = sp.csr_matrix((Jl.data, gcol[Jl.indices], Jl.indptr), shape=(d * p.n_own, prob.N))
        ref = J[d * p.n0:d * p.n1]
        diff = abs(Jl - ref)
        assert diff.max() <= 1e-12 * abs(ref).max()
        assert Jl.nnz == ref.nnz


def test_grid_gmres_single_rank_matches_lu():
    import torch
    from femvf_b200.distributed import DistributedSolid
    res, prob, prop, state, p1, scal = problem(4)   # 62k triangles, 63k DOF
    dt = 1e-4
    ds = DistributedSolid(res, 0, 1, restart=40)
    ds.upload_global(prop, state, p1, scal)
    ds.assemble(dt)
    b = ds.owned('F').clone()
    x = torch.empty_like(b)
    info = ds.solve(b, x, rtol=1e-13)
    J = om.SolidOracle(prob).jac(state['u1'], dt, prop, p1)
    x_ref = spla.splu(J.tocsc()).solve(b.cpu().numpy())
    err = np.linalg.norm(x.cpu().numpy() - x_ref) / np.linalg.norm(x_ref)
    assert err < 1e-9, (err, info)
    assert info['iterations'] < 1000


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [here, os.path.dirname(here), os.path.join(os.path.dirname(here), 'vf-fem_b200')]
    from femvf_b200.distributed import DistributedSolid
    res, prob, prop, state, p1, scal = problem(4)
    dt = 1e-4
    ds = DistributedSolid(res, rank, world, restart=40)
    ds.upload_global(prop, state, p1, scal)
    ds.assemble(dt)
    b = ds.owned('F').clone()
    x = torch.empty_like(b)
    info = ds.solve(b, x, rtol=1e-13)
    gathered = [None] * world
    dist.all_gather_object(gathered, (ds.part.n0, ds.part.n1, x.cpu().numpy(), b.cpu().numpy(), info))
    if rank == 0:
        out.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_distributed_solve_matches_lu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = out.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    res, prob, prop, state, p1, scal = problem(4)
    x = np.concatenate([g[2] for g in gathered])
    b = np.concatenate([g[3] for g in gathered])
    J = om.SolidOracle(prob).jac(state['u1'], 1e-4, prop, p1)
    x_ref = spla.splu(J.tocsc()).solve(b)
    err = np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref)
    assert err < 1e-9, (err, gathered[0][4])
    assert gathered[0][4]['iterations'] == gathered[1][4]['iterations']
