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

    def run_to_statefiles(self, fnames, times, controls, prop, ini_state, emod=None, eta=None,
                          options=None, nchunk: int = 100):
        """
        One ``StateFile`` per member (the drop-in output layout of ``forward.integrate``,
        ``/root/reference/src/femvf/statefile.py:163-339``): the ensemble advances on the device in
        chunks of ``nchunk`` steps; after every chunk the state history of ALL members comes back
        in one device->host copy and is appended to the members' files.

        fnames : one path per member;  times : (nt,) shared time grid
        controls : list of control BlockVectors (control n is used for step n, the last one
        for the remaining steps, as in ``forward.py:169-171``);  prop : the common properties
        (per-member ``emod`` / ``eta`` fields overwrite the common ones)
        Returns the (n_members, nt, 4) info series (num_iter, abs_err, rel_err, min area).
        """
        from . import statefile as sf
        if len(fnames) != self.n_members:
            raise ValueError("one file name per member")
        times = np.asarray(times, dtype=float)
        nsteps = len(times) - 1
        m, e = self.model, self.engine
        self.set_common_prop(prop)
        self.upload_members(np.ascontiguousarray(ini_state), emod, eta)
        n_fluid = e.n_fluid
        ctl_rows = np.array([[np.broadcast_to(c['psub'], (n_fluid,)),
                              np.broadcast_to(c['psup'], (n_fluid,))] for c in controls])
        files = []
        try:
            for b, fname in enumerate(fnames):
                mprop = prop.copy()
                if emod is not None:
                    mprop['emod'][:] = emod[b]
                if eta is not None:
                    mprop['eta'][:] = eta[b]
                m.set_prop(mprop)
                f = sf.StateFile(m, fname, mode='w', NCHUNK=nchunk)
                f.init_layout()                          # forward.py:84-89
                st = m.state0.copy()
                st[:] = ini_state[b]
                f.append_state(st)
                f.append_control(controls[0])
                f.append_time(times[0])
                f.append_solver_info({'num_iter': 0, 'abs_err': 0, 'rel_err': 0})
                f.append_prop(mprop)
                files.append(f)
            m.set_prop(prop)
            infos = np.zeros((self.n_members, nsteps + 1, 4))
            n0 = 0
            while n0 < nsteps:
                n1 = min(n0 + nchunk, nsteps)
                idx = [min(n, len(controls) - 1) for n in range(n0, n1)]
                hs, hi = self.run_device(np.diff(times)[n0:n1], ctl_rows[idx], options,
                                         store_states=True)
                hs, hi = hs.cpu().numpy(), hi.cpu().numpy()
                infos[:, n0 + 1:n1 + 1] = hi[:, 1:]
                for b, f in enumerate(files):
                    for k in range(n1 - n0):
                        st = m.state0.copy()
                        st[:] = hs[b, k + 1]
                        f.append_state(st)
                        f.append_control(controls[idx[k]])
                        f.append_time(times[n0 + k + 1])
                        f.append_solver_info({'num_iter': int(hi[b, k + 1, 0]),
                                              'abs_err': float(hi[b, k + 1, 1]),
                                              'rel_err': float(hi[b, k + 1, 2])})
                n0 = n1
        finally:
            for f in files:
                f.close()
        return infos

