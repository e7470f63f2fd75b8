"""
Grid-wide Newton / linear solves for ONE large mesh behind the model API.

``FenicsModel.solve_state1`` (``/root/reference/src/femvf/models/transient.py:441-468``: Newton
loop, abs 1e-8 / rel 1e-10 / 50 iterations, ``solverconst.py:1-6``) and ``solve_dres_dstate1``
(``transient.py:470-491``: one PETSc LU solve with J_uu) run inside one CTA per member for the
reference-sized meshes (``csrc/member.cu``).  A mesh that does not fit one CTA is solved here
with whole-GPU kernels instead: the pipelined assembly kernel for F and J, and the restarted
GMRES of ``distributed.GridGMRES`` (SpMV, fused multi-dot / multi-axpy kernels, CUDA-graph replay
of Arnoldi blocks) preconditioned by the multicolour block ILU(0) of ``csrc/krylov.cu`` -- the
stand-in for the reference's sparse LU at these sizes -- or by block-Jacobi.

Everything stays on the device; the host reads one residual norm per Newton iteration and one
small Hessenberg block per eight Krylov iterations.
"""

from __future__ import annotations

import os

import numpy as np
import torch

from .distributed import GridGMRES
from .equations import newmark


def band_limit() -> int:
    """Largest number of solid DOFs solved DIRECTLY by the banded LU (``VF_BAND_MAX_DOF``); above
    it the whole-GPU GMRES with the ILU(0) preconditioner takes over."""
    return int(os.environ.get('VF_BAND_MAX_DOF', '20000'))


def grid_threshold(static: bool = False) -> int:
    """Smallest number of solid DOFs solved by the grid-wide path.  Transient steps
    (``VF_GRID_MIN_DOF``, 20000): below it the in-kernel time loop of one CTA wins (no host round
    trip per step, well-conditioned matrices).  Static solves (``VF_GRID_MIN_DOF_STATIC``, 1025):
    the one-CTA solver has a dense inverse up to 1024 DOF and only block-Jacobi GMRES above, which
    needs thousands of iterations on the stiffness matrix with a contact penalty."""
    if static:
        return int(os.environ.get('VF_GRID_MIN_DOF_STATIC',
                                  os.environ.get('VF_GRID_MIN_DOF', '1025')))
    return int(os.environ.get('VF_GRID_MIN_DOF', '20000'))


class GridSolver:
    """Newton and linear solves of one member of an engine with whole-GPU kernels."""

    def __init__(self, engine, member: int = 0, restart: int = 40, precond: str | None = None):
        if member != 0:
            raise ValueError("the grid-wide solver works on member 0 of an engine")
        self.e = engine
        precond = precond or os.environ.get('VF_GRID_PRECOND', 'ilu0')
        # mid-size meshes: banded LU in reverse Cuthill-McKee ordering (csrc/band.cu), one step of
        # iterative refinement; larger ones: restarted GMRES
        self.direct = precond != 'jacobi' and engine.N <= band_limit() and \
            os.environ.get('VF_GRID_DIRECT', '1') != '0'
        if self.direct:
            self.half_bandwidth = engine.band_setup()
            self.gmres = None
            self.res = torch.zeros(engine.N, dtype=torch.float64, device=engine.device)
        else:
            self.gmres = GridGMRES(engine, engine.nn, None, restart, precond=precond)
        self.dx = torch.zeros(engine.N, dtype=torch.float64, device=engine.device)
        self.rhs = torch.zeros_like(self.dx)

    # --- J x = b with the engine's resident J ------------------------------------------------
    def linear_solve(self, b: torch.Tensor, x: torch.Tensor, rtol: float = 1e-12,
                     atol: float = 0.0, maxiter: int = 4000):
        if not self.direct:
            return self.gmres.solve(b, x, rtol=rtol, atol=atol, maxiter=maxiter)
        e = self.e
        e.band_factor(0)
        e.band_solve(b, x)
        # one step of iterative refinement with the matrix itself (no pivoting in the factor)
        e.spmv(x, self.res, 0)
        torch.sub(b, self.res, out=self.res)
        resid0 = float(torch.linalg.vector_norm(self.res).item())
        e.band_solve(self.res, self.res)
        x.add_(self.res)
        return {'iterations': 1, 'restarts': 0, 'residual': resid0,
                'bnorm': float(torch.linalg.vector_norm(b).item()), 'direct': True}

    # --- Newton loop (oracle/model.py SolidOracle.solve_state1 semantics) -----------------------
    def solve_state1(self, dt: float, options: dict, is_static: bool = False):
        e = self.e
        abs_tol = float(options.get('absolute_tolerance', 1e-8))
        rel_tol = float(options.get('relative_tolerance', 1e-10))
        max_it = int(options.get('maximum_iterations', 50))
        lin_rtol = float(options.get('linear_relative_tolerance', 1e-12))
        u1, F = e.view('u1'), e.view('F')
        k, r0 = 0, None
        gmres_iters = 0
        while True:
            # residual and Jacobian in one launch of the pipelined kernel; the Jacobian of the
            # last (converged) iterate is not used
            e.assemble(0, res=True, jac=True, dt=dt, is_static=is_static)
            abs_err = float(torch.linalg.vector_norm(F).item())
            if r0 is None:
                r0 = abs_err
            rel_err = abs_err / r0 if r0 > 0 else 0.0
            if abs_err <= abs_tol or rel_err <= rel_tol or k >= max_it:
                break
            self.rhs.copy_(F)
            info = self.linear_solve(self.rhs, self.dx, rtol=lin_rtol)
            gmres_iters += info['iterations']
            u1.sub_(self.dx)
            k += 1
        if not is_static:
            u0, v0, a0 = e.view('u0'), e.view('v0'), e.view('a0')
            du = u1 - u0
            # newmark.py:8-29, 57-73 (same expressions as equations/newmark.py)
            e.view('v1').copy_(newmark.newmark_v(u1, u0, v0, a0, dt))
            e.view('a1').copy_(newmark.newmark_a(u1, u0, v0, a0, dt))
            del du
        return {'num_iter': k, 'abs_err': abs_err, 'rel_err': rel_err,
                'gmres_iters': gmres_iters, 'gmres_resid': float('nan')}
