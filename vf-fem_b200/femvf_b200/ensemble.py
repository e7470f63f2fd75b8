"""
Parameter-ensemble sweeps of ``forward.integrate`` (BASELINE.json configs[3]).

The reference is serial (``/root/reference/src/femvf/models/fsi.py:38-39``); a sweep over
material fields is a Python loop over independent ``integrate`` calls.  Members are fully
independent, so they are stored member-major in one device arena and advanced by one CTA
each inside a single persistent kernel (``csrc/member_solver.cuh``); across GPUs the member
range is split contiguously, one process per GPU, with no collective on the data path
(SURVEY.md section 8e).
"""

from __future__ import annotations

import numpy as np
import torch

from .engine import Engine
from .models.transient import ExplicitFSIModel


def shard_members(n_total: int, rank: int, world: int):
    """Contiguous member range [lo, hi) owned by ``rank``; sizes differ by at most one."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


class EnsembleRunner:
    """``n_members`` copies of a coupled model differing in their property fields."""

    def __init__(self, model: ExplicitFSIModel, n_members: int, gmres_restart: int = 16):
        self.model = model
        solid, r = model.solid, model.fluid.residual
        self.engine = Engine(
            solid.assembly_tables, s=r.mesh(), fsi_solid=model.fsimap.dofs_solid,
            fsi_fluid=model.fsimap.dofs_fluid, fluid_kind=r.kind, idx_sep=r.idx_sep,
            contact=solid._CONTACT, membrane=solid.residual.form.terms.get('membrane', False),
            damping=solid.residual.form.terms.get('damping', 'kelvin_voigt'),
            n_members=n_members, gmres_restart=gmres_restart)
        self.n_members = n_members
        self.ne = self.engine.ne
        self.state_size = self.engine.state_size

    def set_common_prop(self, prop):
        """Broadcast one property BlockVector to every member."""
        m, e = self.model, self.engine
        m.set_prop(prop)
        dev = e.device
        for name in ('rho', 'eta', 'emod'):
            e.member_view(name).copy_(torch.as_tensor(np.asarray(m.solid.prop[name]), device=dev))
        if m.solid.residual.form.terms.get('membrane', False):
            for name in ('emod_membrane', 'nu_membrane', 'th_membrane'):
                e.member_view(name).copy_(
                    torch.as_tensor(np.asarray(m.solid.prop[name]), device=dev))
        scal = m.solid._scalar_block(float(m.prop['ymid'][0]))
        e.member_view('scal').copy_(torch.as_tensor(scal, device=dev))
        fp = m.fluid._fprop_block().reshape(-1)
        e.member_view('fprop').copy_(torch.as_tensor(fp, device=dev))
        e.member_view('area').copy_(
            torch.as_tensor(np.asarray(m.fluid.control['area']), device=dev))

    def upload_members(self, ini_state: np.ndarray, emod=None, eta=None):
        e = self.engine
        dev = e.device
        N, nq, npp = e.N, e.n_fluid, e.n_fluid * e.ns
        ini = torch.as_tensor(ini_state, device=dev)
        off = 0
        for name, cnt in (('u0', N), ('v0', N), ('a0', N), ('q0', nq), ('p0', npp)):
            e.member_view(name).copy_(ini[:, off:off + cnt])
            off += cnt
        if emod is not None:
            e.member_view('emod').copy_(torch.as_tensor(emod, device=dev))
        if eta is not None:
            e.member_view('eta').copy_(torch.as_tensor(eta, device=dev))

    def run_device(self, dts, controls, options=None, store_states: bool = False):
        """Advance all members from their resident state0; returns device history tensors."""
        return self.engine.integrate(dts, controls, options, store_states=store_states,
                                     store_info=True)

    def run_host(self, dts, controls, ini_state, emod=None, eta=None, options=None):
        """Host buffers in, host buffers out (one C-ABI call, copies inside)."""
        return self.engine.integrate_host(dts, controls, np.ascontiguousarray(ini_state),
                                          None if emod is None else np.ascontiguousarray(emod),
                                          None if eta is None else np.ascontiguousarray(eta),
                                          options)
