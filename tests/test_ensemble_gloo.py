"""world_size-2 gloo test of the ensemble sharding (host logic of the N > 1 path): members are
split contiguously over ranks with no data-path collective; only the aggregate count is reduced."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from femvf_b200.ensemble import shard_members


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = shard_members(n_total, rank, world)
    # each rank would integrate members [lo, hi); seeds are the global member ids
    seeds = torch.arange(lo, hi, dtype=torch.int64)
    work = torch.tensor([float(hi - lo) * 99], dtype=torch.float64)  # member-steps done
    elapsed = torch.tensor([1.0 + 0.5 * rank], dtype=torch.float64)  # pretend device time
    dist.all_reduce(work, op=dist.ReduceOp.SUM)
    dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    gathered = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, torch.tensor([lo, hi]))
    if rank == 0:
        out.put((work.item(), elapsed.item(), [g.tolist() for g in gathered], seeds.tolist()[:2]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_members_covers_everything():
    for n, w in [(1024, 8), (1000, 3), (5, 8), (1, 1)]:
        ranges = [shard_members(n, r, w) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges[:-1], ranges[1:]))
        sizes = [hi - lo for lo, hi in ranges]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_aggregation():
    world, n_total = 2, 1001
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, out)) for r in range(world)]
    for p in procs:
        p.start()
    work, elapsed, ranges, seeds = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert work == n_total * 99        # every member integrated exactly once
    assert elapsed == 1.5              # max over ranks, as bench.py reports
    assert ranges == [[0, 501], [501, 1001]]
    assert seeds == [0, 1]
