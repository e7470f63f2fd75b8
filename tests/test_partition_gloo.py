"""CPU tests of the mesh-partition index logic (LocalPartition, HaloPlan) including a
world_size-2 gloo exchange: after the halo exchange every rank's local vector equals the
global vector at its local vertices."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import mesh_tuples
from femvf_b200 import meshgen
from femvf_b200.distributed import LocalPartition, HaloPlan, partition_starts
from femvf_b200.residuals import solid as slr


def _setup(name='m5'):
    mt = meshgen.renumber_for_locality(mesh_tuples()[name]())
    res = slr.KelvinVoigt(*mt)
    mesh = res.mesh()
    fids, pfc, pfo = res.pressure_facets()
    return mesh, pfc, pfo, res.fixed_dofs()


def test_local_partition_covers_owned_rows():
    mesh, pfc, pfo, fixed = _setup()
    nn = mesh.num_vertices()
    cells = mesh.cells()
    world = 3
    assert partition_starts(nn, world)[-1] == nn
    seen_rows = np.zeros(nn, dtype=int)
    for r in range(world):
        p = LocalPartition(mesh.coordinates(), cells, pfc, pfo, fixed, r, world)
        seen_rows[p.n0:p.n1] += 1
        # every cell touching an owned vertex is local, with all of its vertices
        touching = ((cells >= p.n0) & (cells < p.n1)).any(axis=1)
        assert np.array_equal(np.nonzero(touching)[0], p.cell_ids)
        assert np.array_equal(p.local_nodes[p.cells], cells[p.cell_ids])
        assert np.all(np.diff(p.ghost_global) > 0)
        # pressure facets with an owned vertex are all present
        fnodes = np.array([np.delete(cells[c], o) for c, o in zip(pfc, pfo)])
        need = ((fnodes >= p.n0) & (fnodes < p.n1)).any(axis=1).sum()
        got_nodes = np.array([np.delete(p.cells[c], o) for c, o in zip(p.pf_cell, p.pf_opp)])
        got = (got_nodes < p.n_own).any(axis=1).sum()
        assert got == need
        # Dirichlet flags follow the vertices
        d = 2
        gl = d * p.local_nodes[p.fixed_dofs // d] + p.fixed_dofs % d
        assert set(gl) <= set(fixed)
        own_fixed = fixed[(fixed // d >= p.n0) & (fixed // d < p.n1)]
        assert set(own_fixed) <= set(gl)
    assert np.all(seen_rows == 1)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    mesh, pfc, pfo, fixed = _setup()
    p = LocalPartition(mesh.coordinates(), mesh.cells(), pfc, pfo, fixed, rank, world)
    halo = HaloPlan(p).to_device('cpu')
    nn = mesh.num_vertices()
    xg = np.random.default_rng(0).standard_normal(2 * nn)
    xl = torch.zeros(2 * p.n_local, dtype=torch.float64)
    xl[:2 * p.n_own] = torch.as_tensor(xg[2 * p.n0:2 * p.n1])   # only the owned part is known
    halo.exchange(xl)
    ok = bool(np.array_equal(xl.numpy(), p.local_vector(xg)))
    # Krylov-style reduction: global dot product from owned parts
    part = torch.tensor([float(xl[:2 * p.n_own] @ xl[:2 * p.n_own])], dtype=torch.float64)
    dist.all_reduce(part)
    ok = ok and abs(part.item() - xg @ xg) <= 1e-12 * (xg @ xg)
    gathered = [None] * world
    dist.all_gather_object(gathered, (ok, len(p.ghost_global), halo.halo_bytes))
    if rank == 0:
        out.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_halo_exchange_gloo():
    world = 2
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for ok, _, _ in res)
    assert all(ng > 0 for _, ng, _ in res)
    # what one rank sends is what the other receives
    assert res[0][2] == 16 * res[1][1] and res[1][2] == 16 * res[0][1]
