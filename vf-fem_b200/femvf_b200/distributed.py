"""
Mesh-partitioned assembly and Krylov solve over several GPUs (SURVEY.md section 8e,
BASELINE.json configs[4]).

The reference is strictly serial (``/root/reference/src/femvf/models/fsi.py:38-39``); this
module is how its PETSc solve (``models/transient.py:487``) scales past one GPU.

Partitioning: the vertices (ordered along a space-filling curve, ``meshgen.morton_order``)
are split into ``world`` contiguous ranges; rank r owns the rows of its vertices.  It holds the
cells touching an owned vertex and the ghost vertices they bring along, so

* **assembly needs no communication** (owner-computes rows; ghost cells are duplicated);
* an **SpMV needs one halo exchange** of the ghost entries of x (``d`` doubles per interface
  vertex, one NCCL send/recv per neighbouring rank), issued before the local product;
* the **Krylov reductions** (Gram-Schmidt coefficients, norms) are one small all-reduce each.

One process per GPU; ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU
tests of the index logic) is the plumbing, the numerics are the ``vf_*`` kernels.
With ``world == 1`` the same code is the single-GPU grid-wide GMRES for large meshes.
"""

from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.distributed as dist


def partition_starts(nn: int, world: int) -> np.ndarray:
    base, rem = divmod(nn, world)
    sizes = np.full(world, base, dtype=np.int64)
    sizes[:rem] += 1
    return np.concatenate([[0], np.cumsum(sizes)])


class LocalPartition:
    """Rank-local sub-mesh: owned vertices first (global order), then ghosts (ascending)."""

    def __init__(self, coords, cells, pf_cell, pf_opp, fixed_dofs, rank: int, world: int):
        coords = np.asarray(coords)
        cells = np.asarray(cells, dtype=np.int64)
        nn, d = coords.shape
        self.dim, self.rank, self.world = d, rank, world
        self.starts = partition_starts(nn, world)
        n0, n1 = int(self.starts[rank]), int(self.starts[rank + 1])
        self.n0, self.n1, self.n_own = n0, n1, n1 - n0
        owned_cell = ((cells >= n0) & (cells < n1)).any(axis=1)
        self.cell_ids = np.nonzero(owned_cell)[0]
        lcells_g = cells[self.cell_ids]
        nodes = np.unique(lcells_g)
        ghosts = nodes[(nodes < n0) | (nodes >= n1)]
        self.ghost_global = ghosts  # ascending global ids
        self.local_nodes = np.concatenate([np.arange(n0, n1), ghosts])
        g2l = -np.ones(nn, dtype=np.int64)
        g2l[self.local_nodes] = np.arange(len(self.local_nodes))
        self.g2l = g2l
        self.coords = coords[self.local_nodes]
        self.cells = g2l[lcells_g]
        # pressure facets whose parent cell is local
        pf_cell = np.asarray(pf_cell, dtype=np.int64)
        cell_g2l = -np.ones(len(cells), dtype=np.int64)
        cell_g2l[self.cell_ids] = np.arange(len(self.cell_ids))
        keep = cell_g2l[pf_cell] >= 0
        self.pf_cell = cell_g2l[pf_cell[keep]]
        self.pf_opp = np.asarray(pf_opp, dtype=np.int64)[keep]
        fixed_dofs = np.asarray(fixed_dofs, dtype=np.int64)
        fnode, fcomp = fixed_dofs // d, fixed_dofs % d
        fk = g2l[fnode] >= 0
        self.fixed_dofs = d * g2l[fnode[fk]] + fcomp[fk]

    @property
    def n_local(self) -> int:
        return len(self.local_nodes)

    def local_vector(self, x_global: np.ndarray) -> np.ndarray:
        d = self.dim
        return np.ascontiguousarray(x_global.reshape(-1, d)[self.local_nodes].reshape(-1))

    def local_cell_field(self, f_global: np.ndarray) -> np.ndarray:
        return np.ascontiguousarray(f_global[self.cell_ids])


class HaloPlan:
    """Who sends which owned entries to whom so that every rank's ghost entries get filled."""

    def __init__(self, part: LocalPartition, group=None):
        self.part = part
        self.group = group
        world, rank = part.world, part.rank
        owner = np.searchsorted(part.starts, part.ghost_global, side='right') - 1
        # ghosts are stored after the owned vertices in ascending global order, hence grouped by
        # owner already
        self.recv_from = {}
        for r in np.unique(owner):
            sel = np.nonzero(owner == r)[0]
            self.recv_from[int(r)] = (part.n_own + sel, part.ghost_global[sel])
        if world > 1 and dist.is_available() and dist.is_initialized():
            wanted = {r: g.tolist() for r, (_, g) in self.recv_from.items()}
            gathered = [None] * world
            dist.all_gather_object(gathered, wanted, group=group)
            self.send_to = {}
            for r, req in enumerate(gathered):
                if r != rank and rank in req and len(req[rank]):
                    self.send_to[r] = np.asarray(req[rank], dtype=np.int64) - part.n0
        else:
            # no process group (single rank, or ranks emulated one after another in a test):
            # assembly works, halo exchange is unavailable
            self.send_to = {}

    def to_device(self, device):
        d = self.part.dim
        comp = np.arange(d)
        self._send_idx = {r: torch.as_tensor((d * idx[:, None] + comp).reshape(-1), device=device)
                          for r, idx in self.send_to.items()}
        self._recv_idx = {r: torch.as_tensor((d * loc[:, None] + comp).reshape(-1), device=device)
                          for r, (loc, _) in self.recv_from.items()}
        return self

    def exchange(self, x_local: torch.Tensor):
        """Fill the ghost entries of ``x_local`` (owned entries must be current)."""
        if self.part.world == 1:
            return
        ops, bufs = [], []
        for r, idx in self._send_idx.items():
            buf = x_local[idx].contiguous()
            ops.append(dist.P2POp(dist.isend, buf, r, group=self.group))
            bufs.append(buf)
        recvs = []
        for r, idx in self._recv_idx.items():
            buf = torch.empty(idx.numel(), dtype=x_local.dtype, device=x_local.device)
            ops.append(dist.P2POp(dist.irecv, buf, r, group=self.group))
            recvs.append((idx, buf))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        for idx, buf in recvs:
            x_local[idx] = buf

    @property
    def halo_bytes(self) -> int:
        d = self.part.dim
        return 8 * d * sum(len(v) for v in self.send_to.values())


class GridGMRES:
    """
    Left block-Jacobi preconditioned restarted GMRES(m) with CGS2 over the rows owned by this
    rank.  Host-side control flow (as PETSc's KSP), device kernels for every O(N) operation,
    one halo exchange per operator application and one small all-reduce per reduction; the
    host reads results back only every few iterations (see ``solve``).
    """

    def __init__(self, engine, n_own_nodes: int, halo: Optional[HaloPlan] = None,
                 restart: int = 30, group=None, precond: str = 'jacobi'):
        self.e = engine
        self.d = engine.dim
        self.n_own = n_own_nodes
        self.nown = self.d * n_own_nodes          # owned DOFs
        self.nloc = engine.N                      # owned + ghost DOFs
        self.halo = halo
        self.group = group
        self.m = restart
        dev = engine.device
        f64 = torch.float64
        self.V = torch.zeros((restart + 1, self.nown), dtype=f64, device=dev)
        self.xl = torch.zeros(self.nloc, dtype=f64, device=dev)   # operand with ghosts
        self.w = torch.zeros(self.nown, dtype=f64, device=dev)
        self.t = torch.zeros(self.nown, dtype=f64, device=dev)
        self.h = torch.zeros(restart + 2, dtype=f64, device=dev)
        # per-iteration device record: [h (pass 1) | h (pass 2) | ||w||^2]
        self.Hd = torch.zeros((restart, 2 * (restart + 1) + 1), dtype=f64, device=dev)
        self.scratch = torch.zeros(592 * (restart + 2), dtype=f64, device=dev)
        self.distributed = halo is not None and halo.part.world > 1
        import os
        want = os.environ.get('VF_GMRES_GRAPH')
        self.use_graphs = (not self.distributed and str(dev).startswith('cuda')) \
            if want is None else want == '1'
        self._graphs = {}
        self.spmv_count = 0
        # left preconditioner of the owned diagonal block: 'jacobi' (d x d node blocks) or
        # 'ilu0' (multicolour block ILU(0), csrc/ilu.cu; block-Jacobi across ranks)
        if precond not in ('jacobi', 'ilu0'):
            raise ValueError(f"unknown preconditioner '{precond}'")
        self.precond = precond
        if precond == 'ilu0':
            engine.ilu_setup(0, n_own_nodes)
            self.tl = torch.zeros(self.nloc, dtype=f64, device=dev)

    def _prec_setup(self):
        if self.precond == 'ilu0':
            self.e.ilu_factor(0)
        else:
            self.e.block_jacobi_setup(0, self.n_own)

    def _prec(self, r_own: torch.Tensor, z_own: torch.Tensor):
        """z = M^-1 r on the owned DOFs."""
        if self.precond == 'ilu0':
            self.tl[:self.nown].copy_(r_own)
            self.e.ilu_apply(self.tl, self.tl)
            z_own.copy_(self.tl[:self.nown])
        else:
            self.e.block_jacobi_apply(r_own, z_own, 0, self.n_own)

    def _allreduce(self, t: torch.Tensor):
        if self.distributed:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def apply(self, v_own: torch.Tensor, out_own: torch.Tensor):
        """out = Dinv (J v) on the owned rows; v's ghost entries are fetched first."""
        self.xl[:self.nown].copy_(v_own)
        if self.distributed:
            self.halo.exchange(self.xl)
        self.e.spmv_rows(self.xl, self.t, 0, self.n_own)
        self._prec(self.t, out_own)
        self.spmv_count += 1

    def dots(self, nvec: int, w: torch.Tensor) -> torch.Tensor:
        out = self.h[:nvec]
        self.e.multidot(self.V, nvec, w, self.nown, out, self.scratch)
        self._allreduce(out)
        return out

    def norm(self, w: torch.Tensor) -> float:
        out = self.h[self.m + 1:self.m + 2]
        self.e.multidot(w.view(1, -1), 1, w, self.nown, out, self.scratch)
        self._allreduce(out)
        return float(torch.sqrt(out)[0].item())

    def _arnoldi_steps(self, k: int, kend: int):
        e, m, n, V, Hd = self.e, self.m, self.nown, self.V, self.Hd
        for j in range(k, kend):
            row = Hd[j]
            h1, h2, nrm2 = row[:j + 1], row[m + 1:m + 2 + j], row[2 * m + 2:2 * m + 3]
            # the new vector is built in place in V[j+1], so that ONE multidot (and one
            # all-reduce) of the second pass returns both V_i.w (i <= j) and w.w
            w = V[j + 1]
            self.apply(V[j], w)
            e.multidot(V, j + 1, w, n, h1, self.scratch)
            self._allreduce(h1)
            e.multi_axpy(V, j + 1, h1, w, n)
            h2n = row[m + 1:m + 3 + j]                  # [h2 (j+1 entries) | w.w before pass 2]
            e.multidot(V, j + 2, w, n, h2n, self.scratch)
            self._allreduce(h2n)
            e.multi_axpy(V, j + 1, h2, w, n)
            # ||w - V h2||^2 = w.w - |h2|^2 (V orthonormal, h2 = O(eps |w|): no cancellation)
            e.scale_rsqrt(w, h2n[j + 1:j + 2], w, n, sub=h2, s_out=nrm2)

    def _arnoldi_block(self, k: int, kend: int):
        """Arnoldi iterations k..kend-1 on the device.  On a single rank the block's launch
        sequence (about 15 small kernels per iteration, all on fixed buffers) is captured once
        into a CUDA graph and replayed, which removes the per-launch host cost."""
        if not self.use_graphs:
            self._arnoldi_steps(k, kend)
            return
        key = (k, kend)
        g = self._graphs.get(key)
        if g is None:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            count = self.spmv_count
            with torch.cuda.graph(g):
                self._arnoldi_steps(k, kend)
            self.spmv_count = count
            self._graphs[key] = g
        g.replay()
        self.spmv_count += kend - k

    def solve(self, b_own: torch.Tensor, x_own: torch.Tensor, rtol: float = 1e-12,
              atol: float = 0.0, maxiter: int = 2000, check_every: int = 8):
        """
        Solve J x = b on the owned rows; ``x_own`` is overwritten (initial guess 0).

        The Arnoldi loop is asynchronous: the Hessenberg column (both CGS passes) and the
        squared norm of every iteration stay in device memory, the new basis vector is
        normalised by ``vf_scale_rsqrt`` from that device scalar, and the host reads them back
        (one copy) only every ``check_every`` iterations to advance the Givens recurrence and
        test convergence.  Iterations computed past the converged one are discarded.
        """
        e, m, n = self.e, self.m, self.nown
        self._prec_setup()
        x_own.zero_()
        w, V = self.w, self.V
        self._prec(b_own, w)
        bnorm = self.norm(w)
        info = {'iterations': 0, 'bnorm': bnorm, 'residual': bnorm, 'restarts': 0}
        if bnorm == 0.0:
            return info
        tol = max(rtol * bnorm, atol)
        beta = bnorm
        iters = 0
        first = True
        H = np.zeros((m + 1, m))
        Hd = self.Hd                                    # (m, 2 (m+1) + 1) device rows
        while True:
            if not first:
                self.apply(x_own, w)                    # w = Dinv J x
                self._prec(b_own, self.t)
                e.axpby(1.0, self.t, -1.0, w, n)        # w = M^-1 b - M^-1 J x
                beta = self.norm(w)
                info['restarts'] += 1
                if beta <= tol:
                    info['residual'] = beta
                    break
            first = False
            e.axpby(1.0 / beta, w, 0.0, V[0], n)
            g = np.zeros(m + 1)
            g[0] = beta
            cs, sn = np.zeros(m), np.zeros(m)
            k, resid = 0, beta
            done = False
            Hd.zero_()
            while k < m and not done:
                kend = min(k + check_every, m, k + max(maxiter - iters, 1))
                self._arnoldi_block(k, kend)            # device only, no host read
                blk = Hd[k:kend].cpu().numpy()          # the only synchronisation of the block
                k0 = k
                for j in range(k0, kend):
                    r = blk[j - k0]
                    hk1 = float(np.sqrt(max(r[2 * m + 2], 0.0)))
                    col = np.zeros(m + 1)
                    col[:j + 1] = r[:j + 1] + r[m + 1:m + 2 + j]
                    for i in range(j):
                        t0 = cs[i] * col[i] + sn[i] * col[i + 1]
                        col[i + 1] = -sn[i] * col[i] + cs[i] * col[i + 1]
                        col[i] = t0
                    denom = np.hypot(col[j], hk1)
                    c, s_ = (1.0, 0.0) if denom == 0 else (col[j] / denom, hk1 / denom)
                    cs[j], sn[j] = c, s_
                    col[j] = c * col[j] + s_ * hk1
                    g[j + 1] = -s_ * g[j]
                    g[j] = c * g[j]
                    H[:, j] = col
                    resid = abs(g[j + 1])
                    iters += 1
                    k = j + 1
                    if resid <= tol or iters >= maxiter or hk1 == 0:
                        done = True
                        break
            y = np.linalg.solve(np.triu(H[:k, :k]), g[:k]) if k else np.zeros(0)
            yd = torch.as_tensor(-y, device=x_own.device)
            e.multi_axpy(V, k, yd, x_own, n)            # x += V y
            info['iterations'] = iters
            info['residual'] = resid
            if resid <= tol or iters >= maxiter:
                break
        return info


class DistributedSolid:
    """One rank's share of a solid model: local assembly + distributed linear solve."""

    def __init__(self, residual, rank: int = 0, world: int = 1, group=None, restart: int = 30,
                 device=None):
        from . import tables as _tables
        from .engine import Engine
        mesh = residual.mesh()
        fids, pf_cell, pf_opp = residual.pressure_facets()
        self.part = LocalPartition(mesh.coordinates(), mesh.cells(), pf_cell, pf_opp,
                                   residual.fixed_dofs(), rank, world)
        p = self.part
        self.tables = _tables.build_tables(p.coords, p.cells, p.pf_cell, p.pf_opp, p.fixed_dofs)
        self.engine = Engine(self.tables, device=device)
        self.halo = HaloPlan(p, group).to_device(self.engine.device)
        self.gmres = GridGMRES(self.engine, p.n_own, self.halo, restart, group)
        self.dim = p.dim

    def upload_global(self, prop: dict, state: dict, p1: np.ndarray, scal: np.ndarray):
        """Scatter globally defined fields to this rank (setup / test convenience)."""
        e, p = self.engine, self.part
        for name in ('rho', 'eta', 'emod'):
            e.upload(name, p.local_cell_field(np.asarray(prop[name])), 0)
        e.upload('scal', scal, 0)
        for name in ('u1', 'u0', 'v0', 'a0'):
            e.upload(name, p.local_vector(np.asarray(state[name])), 0)
        e.upload('p1', np.ascontiguousarray(np.asarray(p1)[p.local_nodes]), 0)

    def assemble(self, dt: float, res: bool = True, jac: bool = True):
        self.engine.assemble(0, res=res, jac=jac, dt=dt)

    def owned(self, name: str) -> torch.Tensor:
        return self.engine.view(name)[:self.dim * self.part.n_own]

    def solve(self, b_own: torch.Tensor, x_own: torch.Tensor, **kw):
        return self.gmres.solve(b_own, x_own, **kw)
