"""
Transient model interface: mirror of ``/root/reference/src/femvf/models/transient.py``.

Class and method names, argument meaning and ``info`` keys follow the reference
(``BaseTransientModel`` :32-154, ``FenicsModel`` :221-513, ``NodalContactModel`` :516-583,
``JaxModel`` :590-672, ``BaseTransientFSIModel`` :678-817, ``ExplicitFSIModel`` :821-920) so
that ``forward.integrate`` and ``StateFile`` consume these objects unchanged.  State, control
and property vectors are host ``BlockVector``s exactly as in the reference; every assembly,
linear solve, Newton loop and fluid evaluation is a CUDA kernel reached through the C ABI
(``engine.Engine``).  Host vectors are pushed to the device before each such call and results
are pulled back, which is the "host buffer" path; ``forward.integrate`` uses the
device-resident time loop instead.

Documented deviations:
  Q1/Q2  v1, a1 are computed from the Newmark relations after the solve for u1; the
         off-diagonal blocks returned by ``assem_dres_dstate1`` are the nodal
         -cv*I / -ca*I consistent with the nodal v/a residual rows; ``num_iter`` counts
         Newton updates of u1 (1 for the affine 2D problem).
  Q3     contact is active only in ``NodalContactModel`` (as in the reference).
"""

from __future__ import annotations

from typing import Any, Optional

import numpy as np
import scipy.sparse as sp

from .. import blockvec as bv
from ..blockvec import BlockVector
from ..engine import Engine, SCAL, SCAL_COUNT, FPROP, FPROP_COUNT
from ..equations import newmark
from ..residuals import solid as slr, fluid as flr
from ..solverconst import DEFAULT_NEWTON_SOLVER_PRM, FIXEDPOINT_SOLVER_PRM
from .. import tables as _tables
from . import fsi

BlockVec = BlockVector


class BlockMatrix:
    """Labelled 2D collection of sparse blocks with ``.sub[row, col]`` access."""

    def __init__(self, mats, shape, labels):
        self.mats = list(mats)
        self.shape = tuple(shape)
        self.labels = (tuple(labels[0]), tuple(labels[1]))

    class _Sub:
        def __init__(self, bm):
            self._bm = bm

        def __getitem__(self, key):
            r, c = key
            i = self._bm.labels[0].index(r) if isinstance(r, str) else int(r)
            j = self._bm.labels[1].index(c) if isinstance(c, str) else int(c)
            return self._bm.mats[i * self._bm.shape[1] + j]

    @property
    def sub(self):
        return BlockMatrix._Sub(self)

    def __bool__(self):
        return True


class BaseTransientModel:
    """One time step of a system: residual F(u1, u0, g, p, dt) (``transient.py:32-154``)."""

    @property
    def dt(self):
        raise NotImplementedError(f"Subclass {type(self)} must implement this function")

    def set_ini_state(self, state0: BlockVec):
        raise NotImplementedError(f"Subclass {type(self)} must implement this function")

    def set_fin_state(self, state1: BlockVec):
        raise NotImplementedError(f"Subclass {type(self)} must implement this function")

    def set_control(self, control: BlockVec):
        raise NotImplementedError(f"Subclass {type(self)} must implement this function")

    def set_prop(self, prop: BlockVec):
        raise NotImplementedError(f"Subclass {type(self)} must implement this function")

    def assem_res(self) -> BlockVec:
        raise NotImplementedError(f"Subclass {type(self)} must implement this function")

    def assem_dres_dstate1(self):
        raise NotImplementedError(f"Subclass {type(self)} must implement this function")

    def assem_dres_dstate0(self):
        raise NotImplementedError(f"Subclass {type(self)} must implement this function")

    def assem_dres_dcontrol(self):
        raise NotImplementedError(f"Subclass {type(self)} must implement this function")

    def assem_dres_dprops(self):
        raise NotImplementedError(f"Subclass {type(self)} must implement this function")

    def solve_state1(self, state1: BlockVec, options: Optional[dict[str, Any]]):
        raise NotImplementedError(f"Subclass {type(self)} must implement this function")


## Solid type models


def properties_bvec_from_forms(form, defaults=None):
    """Property BlockVector in coefficient insertion order (``transient.py:187-218``)."""
    defaults = {} if defaults is None else defaults
    labels = [key.split('/')[-1] for key in form.keys() if key.split('/', 1)[0] == 'prop']
    vecs = []
    for label in labels:
        vec = np.array(form['prop/' + label].vector(), dtype=np.float64, copy=True)
        if label in defaults:
            vec[:] = defaults[label]
        vecs.append(vec)
    return BlockVector(vecs, labels=[labels])


class FenicsModel(BaseTransientModel):
    """Discretised governing equations of the solid (``transient.py:221-513``)."""

    FORM_KEYS = ('u', 'v', 'a')
    STATE0_KEYS = ('state/u0', 'state/v0', 'state/a0')
    STATE1_KEYS = ('state/u1', 'state/v1', 'state/a1')
    CONTROL_KEYS = ('control/p1',)
    _CONTACT = False

    def __init__(self, residual: slr.FenicsResidual):
        self._residual = residual
        form = residual.form
        mesh = residual.mesh()
        d = mesh.topology().dim()
        N = d * mesh.num_vertices()
        # modify_newmark_time_discretization (equations/form.py:1067-1113): add u0, v0, a0, dt
        from ..residuals.base import Coefficient, FunctionSpace
        for key in ('state/u0', 'state/v0', 'state/a0'):
            if key not in form:
                form.coefficients[key] = Coefficient(FunctionSpace(mesh, 'CG', 1, d))
        if 'time/dt' not in form:
            form.coefficients['time/dt'] = Coefficient(FunctionSpace(mesh, 'R', 0, 1),
                                                       constant=True, default=0.0)

        self.state0 = BlockVector(
            [form['state/u0'].vector(), form['state/v0'].vector(), form['state/a0'].vector()],
            labels=[('u', 'v', 'a')])
        self.state1 = BlockVector(
            [form['state/u1'].vector(), form['state/v1'].vector(), form['state/a1'].vector()],
            labels=[('u', 'v', 'a')])
        self.control = BlockVector([form['control/p1'].vector()], labels=[('p',)])
        self.prop = properties_bvec_from_forms(form)
        assert self.state0['u'].size == N

        # reference configuration of the mesh: 'prop/umesh' displaces it (transient.py:347-360)
        self._ref_coords = mesh.coordinates().copy()
        self._build_tables()
        self._engine: Optional[Engine] = None
        self._engine_provider = None
        self._engine_attached = False
        self._geometry_listeners = []   # called after the mesh coordinates changed
        self._member = 0
        # Host -> device synchronisation policy.  The reference mutates host vectors freely, so
        # by default every device call re-uploads state, control and properties.  With
        # ``trust_setters = True`` only the groups changed through set_ini_state /
        # set_fin_state / set_control / set_prop / dt since the last device call are uploaded
        # (in-place edits of the BlockVectors then need ``mark_dirty()``), and large results
        # (F, J values, the constant Jacobian blocks) are returned as views of cached
        # page-locked buffers that the next call of the same method overwrites.
        self.trust_setters = False
        self._dirty = {'prop': True, 'state0': True, 'state1': True, 'control': True}
        self.set_prop(self.prop)

    # --- engine management -----------------------------------------------------------
    @property
    def engine(self) -> Engine:
        if self._engine is None:
            if self._engine_provider is not None:
                self._engine = self._engine_provider()
            else:
                self._engine = Engine(
                    self._tables, contact=self._CONTACT,
                    membrane=self.residual.form.terms.get('membrane', False),
                    damping=self.residual.form.terms.get('damping', 'kelvin_voigt'))
        return self._engine

    def _attach_engine(self, engine: Engine, member: int = 0):
        self._engine = engine
        self._engine_attached = True
        self._member = member

    def _build_tables(self):
        residual = self._residual
        mesh = residual.mesh()
        fids, pf_cell, pf_opp = residual.pressure_facets()
        self._tables = _tables.build_tables(mesh.coordinates(), mesh.cells(), pf_cell, pf_opp,
                                            residual.fixed_dofs())

    def _apply_mesh_displacement(self):
        """``transient.py:347-360``: mesh coordinates = reference coordinates + umesh (the DOFs of
        this package are vertex-major interleaved, so VERT_TO_VDOF is the identity).  The device
        tables hold the geometry, so a changed shape drops the engine; the next device call
        builds a new one and uploads state, control and properties again."""
        mesh = self._residual.mesh()
        new = self._ref_coords + np.asarray(
            self._residual.form['prop/umesh'].vector()).reshape(self._ref_coords.shape)
        if np.array_equal(new, mesh.coordinates()):
            return
        if self._engine_attached:
            raise NotImplementedError(
                "'umesh' cannot move the mesh of an engine shared by an ensemble")
        mesh.coordinates()[:] = new
        self._build_tables()
        self._retire_live_jacobian()
        self._engine = None
        self._grid = None
        self.mark_dirty()
        for notify in self._geometry_listeners:
            notify()

    @property
    def assembly_tables(self) -> dict:
        return self._tables

    @property
    def residual(self) -> slr.FenicsResidual:
        return self._residual

    @property
    def XREF(self) -> np.ndarray:
        """Reference nodal coordinates, interleaved (``transient.py:276-287``)."""
        return self.residual.mesh().coordinates().reshape(-1).copy()

    # --- parameter setting -------------------------------------------------------------
    @property
    def dt(self):
        return self.residual.form['time/dt'].vector()[0]

    @dt.setter
    def dt(self, value):
        self.residual.form['time/dt'].vector()[:] = value

    def mark_dirty(self, *groups):
        for g in (groups or self._dirty.keys()):
            self._dirty[g] = True

    def set_ini_state(self, state):
        self.state0[:] = state
        self._dirty['state0'] = True

    def set_fin_state(self, state):
        self.state1[:] = state
        self._dirty['state1'] = True

    def set_control(self, p1):
        self.control[:] = p1
        self._dirty['control'] = True

    def set_prop(self, prop):
        self._dirty['prop'] = True
        for key, value in prop.sub_items():
            self.residual.form['prop/' + key].vector()[:] = np.ravel(value) \
                if np.size(value) > 1 else np.ravel(value)[0]
        if prop is not self.prop:
            self.prop[:] = prop
        if 'prop/umesh' in self.residual.form:
            self._apply_mesh_displacement()

    # --- host -> device ----------------------------------------------------------------
    def _scalar_block(self, ymid: float = 0.0) -> np.ndarray:
        scal = np.zeros(SCAL_COUNT)
        p = self.prop
        d = self.residual.mesh().topology().dim()
        scal[SCAL['nu']] = p['nu'][0]
        scal[SCAL['ycontact']] = p['ycontact'][0]
        scal[SCAL['kcontact']] = p['kcontact'][0]
        scal[SCAL['ncontact']:SCAL['ncontact'] + d] = p['ncontact']
        scal[SCAL['ymid']] = ymid
        if 'rayleigh_m' in p:
            scal[SCAL['rayleigh_m']] = p['rayleigh_m'][0]
            scal[SCAL['rayleigh_k']] = p['rayleigh_k'][0]
        return scal

    def _push_prop(self, ymid: float = 0.0):
        e, m, p = self.engine, self._member, self.prop
        for name in ('rho', 'eta', 'emod'):
            e.upload(name, p[name] if name in p else 0.0, m)
        if self.residual.form.terms.get('membrane', False):
            for name in ('emod_membrane', 'nu_membrane', 'th_membrane'):
                e.upload(name, p[name], m)
        e.upload('scal', self._scalar_block(ymid), m)

    def _push_state(self):
        e, m = self.engine, self._member
        for name, vec in zip(('u0', 'v0', 'a0'), self.state0.vecs):
            e.upload(name, vec, m)
        for name, vec in zip(('u1', 'v1', 'a1'), self.state1.vecs):
            e.upload(name, vec, m)
        e.upload('p1', self.control['p'], m)

    def _push_all(self):
        e, m = self.engine, self._member
        every = not self.trust_setters
        if every or self._dirty['prop']:
            self._push_prop(getattr(self, '_ymid', 0.0))
        lock = self.trust_setters  # the model's own state vectors are page-locked in place
        if every or self._dirty['state0']:
            for name, vec in zip(('u0', 'v0', 'a0'), self.state0.vecs):
                e.upload(name, vec, m, persistent=lock)
        if every or self._dirty['state1']:
            for name, vec in zip(('u1', 'v1', 'a1'), self.state1.vecs):
                e.upload(name, vec, m, persistent=lock)
        if every or self._dirty['control']:
            e.upload('p1', self.control['p'], m)
        for g in self._dirty:
            self._dirty[g] = False

    # --- residual and sensitivities -----------------------------------------------------
    def assem_res(self):
        """``transient.py:363-382``: F_u by device assembly; F_v, F_a nodal; BCs on F_u."""
        self._push_all()
        self.engine.assemble(self._member, res=True, jac=False, dt=self.dt)
        res_u = self.engine.download('F', self._member, pinned=self.trust_setters)
        res_v, res_a = self.engine.newmark_residual(self.dt, self._member,
                                                    pinned=self.trust_setters)
        return BlockVector([res_u, res_v, res_a], labels=(self.FORM_KEYS,))

    def csr_pattern(self):
        if not hasattr(self, '_pattern'):
            self._pattern = self.engine.csr_pattern()
        return self._pattern

    # --- the engine has ONE J array: value semantics for the matrices handed out -------------
    lazy_jacobian = True   # assem_dres_dstate1 returns dF_u/du1 as a device-resident DeviceCSR

    def _live_jacobian(self):
        ref = getattr(self, '_live_jac_ref', None)
        return ref() if ref is not None else None

    def _retire_live_jacobian(self):
        """Before anything overwrites the engine's J: a DeviceCSR that is still referenced and
        still on the device keeps its values in a private device tensor (copy-on-write)."""
        live = self._live_jacobian()
        if live is not None:
            live._detach()
        self._live_jac_ref = None

    def _grid_solver(self, static: bool = False):
        """The whole-GPU solver for meshes that do not fit one CTA (None for small meshes, and
        for engines shared by an ensemble)."""
        from .. import gridsolve
        e = self.engine
        if e.N < gridsolve.grid_threshold(static) or e.n_members != 1 or self._member != 0:
            return None
        gs = getattr(self, '_grid', None)
        if gs is None or gs.e is not e:
            gs = self._grid = gridsolve.GridSolver(e)
        return gs

    def _assem_jac_uu(self, is_static: bool = False):
        self._push_all()
        self._retire_live_jacobian()
        self.engine.assemble(self._member, res=False, jac=True, dt=self.dt, is_static=is_static)
        if self.lazy_jacobian:
            import weakref
            from ..devmat import DeviceCSR
            A = DeviceCSR(self, self._member, pinned=self.trust_setters)
            self._live_jac_ref = weakref.ref(A)
            return A
        vals = self.engine.download('J', self._member, pinned=self.trust_setters)
        rowptr, colidx = self.csr_pattern()
        N = self.state0['u'].size
        return sp.csr_matrix((vals, colidx, rowptr), shape=(N, N))

    def assem_dres_dstate1(self):
        """``transient.py:384-406``: block Jacobian labelled (FORM_KEYS, STATE1_KEYS)."""
        N = self.state0['u'].size
        dt = self.dt
        # the constant blocks depend on (N, dt) only: built once per time step size
        const = getattr(self, '_jac_const_blocks', None)
        if const is None or const[0] != (N, dt) or not self.trust_setters:
            eye = sp.identity(N, format='csr')
            zero = sp.csr_matrix((N, N))
            const = ((N, dt), eye, zero, -newmark.newmark_v_du1(dt) * eye,
                     -newmark.newmark_a_du1(dt) * eye)
            self._jac_const_blocks = const
        _, eye, zero, dv_du, da_du = const
        mats = [
            self._assem_jac_uu(), zero, zero,
            dv_du, eye, zero,
            da_du, zero, eye,
        ]
        return BlockMatrix(mats, (3, 3), (self.FORM_KEYS, self.STATE1_KEYS))

    def assem_dres_dstate0(self):
        """``transient.py:408-421``: block matrix labelled (FORM_KEYS, STATE0_KEYS); as in the
        reference no Dirichlet condition is applied.  F_u depends on state0 only through
        v_nmk and a_nmk, so with C = dF_u/dv1 and M = dF_u/da1
        ``dF_u/dx0 = C dv_nmk/dx0 + M da_nmk/dx0`` -- one device assembly per block with the
        corresponding weights of (K, C, M, K_p).  The v, a rows are the nodal Newmark
        relations."""
        N = self.state0['u'].size
        dt = self.dt
        self._push_all()
        self._retire_live_jacobian()
        rowptr, colidx = self.csr_pattern()
        eye = sp.identity(N, format='csr')
        cv_c = {'u': newmark.newmark_v_du0(dt), 'v': newmark.newmark_v_dv0(dt),
                'a': newmark.newmark_v_da0(dt)}
        ca_c = {'u': newmark.newmark_a_du0(dt), 'v': newmark.newmark_a_dv0(dt),
                'a': newmark.newmark_a_da0(dt)}
        mats = {}
        for key in ('u', 'v', 'a'):
            self.engine.assemble_mix((0.0, cv_c[key], ca_c[key], 0.0), dt, self._member)
            vals = self.engine.download('J', self._member)
            mats['u', key] = sp.csr_matrix((vals, colidx, rowptr), shape=(N, N))
            mats['v', key] = -cv_c[key] * eye        # F_v = v1 - v_nmk(u1, u0, v0, a0)
            mats['a', key] = -ca_c[key] * eye
        flat = [mats[r, c] for r in self.FORM_KEYS for c in ('u', 'v', 'a')]
        return BlockMatrix(flat, (3, 3), (self.FORM_KEYS, self.STATE0_KEYS))

    def assem_dres_dcontrol(self):
        """``transient.py:423-435``: (FORM_KEYS, CONTROL_KEYS); only dF_u/dp1 is non-zero (the
        follower pressure is linear in the nodal pressure)."""
        N = self.state0['u'].size
        nn = self.control['p'].size
        d = N // nn
        self._push_all()
        rows, cols, blocks = self.engine.pressure_control_blocks(self._member)
        ii = (d * rows[:, None] + np.arange(d)[None, :]).ravel()
        jj = np.repeat(cols, d)
        dfu_dp = sp.coo_matrix((blocks.ravel(), (ii, jj)), shape=(N, nn)).tocsr()
        zero = sp.csr_matrix((N, nn))
        return BlockMatrix([dfu_dp, zero, zero], (3, 1), (self.FORM_KEYS, self.CONTROL_KEYS))

    def assem_dres_dprops(self):
        raise NotImplementedError("Not implemented yet!")  # transient.py:437-438

    # --- solvers -------------------------------------------------------------------------
    def solve_state1(self, state1, options=None):
        """``transient.py:441-468``: device Newton solve for u1, then Newmark v1, a1."""
        if options is None:
            options = DEFAULT_NEWTON_SOLVER_PRM
        self.set_fin_state(state1)
        self._push_all()
        self._retire_live_jacobian()
        e, m = self.engine, self._member
        grid = self._grid_solver()
        if grid is not None:
            # a mesh that does not fit one CTA: whole-GPU assembly + ILU(0)-GMRES (gridsolve.py)
            ginfo = grid.solve_state1(self.dt, dict(options))
            x = state1.copy()
            x['u'][:] = e.download('u1', m)
            x['v'][:] = e.download('v1', m)
            x['a'][:] = e.download('a1', m)
            return x, ginfo
        e.solve_state1(self.dt, m, 1, options)
        x = state1.copy()
        x['u'][:] = e.download('u1', m)
        x['v'][:] = e.download('v1', m)
        x['a'][:] = e.download('a1', m)
        info = e.download('info', m)
        solve_info = {'num_iter': int(info[0]), 'abs_err': float(info[1]),
                      'rel_err': float(info[2]), 'gmres_iters': int(info[3]),
                      'gmres_resid': float(info[4])}
        return x, solve_info

    def solve_dres_dstate1(self, dres_dstate1, x, b):
        """``transient.py:470-491``: one J_uu solve on the device + nodal v/a rows."""
        import torch
        from ..devmat import DeviceCSR
        e, m = self.engine, self._member
        bu, bv_, ba = b.sub_blocks
        # the solve runs on the engine's resident J: make sure it holds the values of the
        # matrix that was passed in (it does, without any copy, when the matrix came from the
        # latest assem_dres_dstate1 call)
        A = dres_dstate1.sub['u', 'state/u1']
        if not (isinstance(A, DeviceCSR) and A._is_live()):
            self._retire_live_jacobian()
            if isinstance(A, DeviceCSR) and A.on_device:
                e.view('J', m).copy_(A._values_tensor())
            else:
                e.upload('J', np.ascontiguousarray(A.tocsr().data), m)
        b_t = torch.as_tensor(np.ascontiguousarray(bu), device=e.device)
        x_t = torch.empty_like(b_t)
        grid = self._grid_solver()
        if grid is not None:
            grid.linear_solve(b_t, x_t)
        else:
            e.linear_solve(b_t, x_t, m)
        xu = x_t.cpu().numpy()
        x['u'][:] = xu
        x['v'][:] = bv_ - dres_dstate1.sub['v', 'state/u1'] @ xu
        x['a'][:] = ba - dres_dstate1.sub['a', 'state/u1'] @ xu
        return x


class NodalContactModel(FenicsModel):
    """Solid with nodal cubic-penalty contact tractions (``transient.py:516-583``)."""

    _CONTACT = True

    def set_fin_state(self, state):
        super().set_fin_state(state)
        self.residual.form['control/tcontact'].vector()[:] = self._contact_traction(
            self.state1.sub['u'])

    def _contact_traction(self, u):
        """Nodal contact traction (host mirror of ``control/tcontact``; the kernels evaluate
        the same expression on the fly, ``csrc/elem.cuh`` ``contact_pressure``)."""
        form = self.residual.form
        ndim = self.residual.mesh().topology().dim()
        ycontact = form['prop/ycontact'].values()[0]
        ncontact = form['prop/ncontact'].values()
        kcontact = form['prop/kcontact'].values()[0]
        gap = np.dot((self.XREF + u).reshape(-1, ndim), ncontact) - ycontact
        with np.errstate(invalid='ignore'):
            pgap = np.where(gap == -np.inf, 0.0, (gap + np.abs(gap)) / 2)
        return (-(kcontact * pgap**3)[:, None] * ncontact).reshape(-1).copy()


## Fluid type models


class JaxModel(BaseTransientModel):
    """1D Bernoulli fluid (``transient.py:590-672``); evaluated by the device fluid kernel."""

    def __init__(self, residual: flr.JaxResidual):
        self._residual = residual
        state, control, prop = residual.res_args
        self.state0 = BlockVector(list(state.values()), labels=[list(state.keys())])
        self.state1 = self.state0.copy()
        self.control = BlockVector(list(control.values()), labels=[list(control.keys())])
        self.prop = BlockVector(list(prop.values()), labels=[list(prop.keys())])
        self._dt = 0.0
        self._engine: Optional[Engine] = None
        self._engine_provider = None
        self._member = 0

    @property
    def residual(self) -> flr.JaxResidual:
        return self._residual

    @property
    def fluid(self):
        return self

    @property
    def engine(self) -> Engine:
        if self._engine is None and self._engine_provider is not None:
            self._engine = self._engine_provider()
        if self._engine is None:
            # a fluid on its own: the engine still needs a (dummy one-cell) solid mesh
            coords = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])
            cells = np.array([[0, 1, 2]])
            tb = _tables.build_tables(coords, cells, [], [], [])
            r = self._residual
            self._engine = Engine(tb, s=r.mesh(), fluid_kind=r.kind, idx_sep=r.idx_sep)
        return self._engine

    def _attach_engine(self, engine: Engine, member: int = 0):
        self._engine = engine
        self._member = member

    @property
    def dt(self):
        return self._dt

    @dt.setter
    def dt(self, value):
        self._dt = value

    def set_ini_state(self, state):
        self.state0[:] = state

    def set_fin_state(self, state):
        self.state1[:] = state

    def set_control(self, control):
        self.control[:] = control

    def set_prop(self, prop):
        self.prop[:] = prop

    def _fprop_block(self) -> np.ndarray:
        n_fluid = self.state0['q'].size
        fp = np.zeros((n_fluid, FPROP_COUNT))
        fp[:, FPROP['rho_air']] = 1.0
        fp[:, FPROP['r_sep']] = 1.0
        fp[:, FPROP['zeta_min']] = 1.0
        fp[:, FPROP['zeta_sep']] = 1.0
        for key, col in FPROP.items():
            if key in self.prop:
                fp[:, col] = self.prop[key]
        return fp

    def _push(self):
        e, m = self.engine, self._member
        e.upload('area', self.control['area'], m)
        e.upload('psub', self.control['psub'], m)
        e.upload('psup', self.control['psup'], m)
        e.upload('fprop', self._fprop_block(), m)

    def _bernoulli_qp(self) -> BlockVector:
        self._push()
        e, m = self.engine, self._member
        e.fluid_solve(m, 1)
        return BlockVector([e.download('q1', m), e.download('pf1', m)],
                           labels=self.state1.labels)

    def assem_res(self):
        """state1 - (q, p)(control, prop)  (``fluid.py:286-294``)."""
        return self.state1 - self._bernoulli_qp()

    def solve_state1(self, state1, options=None):
        """``transient.py:667-672``: state1 - res(state1) = Bernoulli (q, p)."""
        info = {}
        return self._bernoulli_qp(), info


## Coupled models


class BaseTransientFSIModel(BaseTransientModel):
    """Coupled solid + 1D fluid (``transient.py:678-817``)."""

    def __init__(self, solid: FenicsModel, fluid: JaxModel, solid_fsi_dofs, fluid_fsi_dofs):
        self.solid = solid
        self.fluid = fluid

        self.state0 = bv.concatenate([solid.state0, fluid.state0])
        self.state1 = bv.concatenate([solid.state1, fluid.state1])
        # the control is just the subglottal and supraglottal pressures (transient.py:714-715)
        self.control = fluid.control[1:]
        _self_properties = BlockVector((np.array([1.0]),), labels=(('ymid',),))
        self.prop = bv.concatenate([solid.prop, fluid.prop, _self_properties])

        (self._fsimap, self._solid_area, self._dflarea_dslu, self._dslp_dflp, _) = \
            fsi.make_coupling_stuff(solid, fluid, solid_fsi_dofs, fluid_fsi_dofs)

        # one shared device engine for solid + fluid + coupling, created on first use
        self._engine: Optional[Engine] = None
        self._fsi_dofs = (np.asarray(solid_fsi_dofs), np.asarray(fluid_fsi_dofs))
        solid._engine, fluid._engine = None, None
        solid._engine_provider = lambda: self.engine
        fluid._engine_provider = lambda: self.engine
        solid._geometry_listeners.append(self._drop_engine)

    def _drop_engine(self):
        """The solid's mesh moved ('prop/umesh'): the shared engine holds the old geometry."""
        self._engine = None
        self.solid._engine = None
        self.fluid._engine = None
        if hasattr(self.fluid, 'mark_dirty'):
            self.fluid.mark_dirty()

    @property
    def engine(self) -> Engine:
        if self._engine is None:
            solid, r = self.solid, self.fluid.residual
            self._engine = Engine(
                solid.assembly_tables, s=r.mesh(), fsi_solid=self._fsi_dofs[0],
                fsi_fluid=self._fsi_dofs[1], fluid_kind=r.kind, idx_sep=r.idx_sep,
                contact=solid._CONTACT,
                membrane=solid.residual.form.terms.get('membrane', False),
                damping=solid.residual.form.terms.get('damping', 'kelvin_voigt'))
        return self._engine

    @property
    def fsimap(self):
        return self._fsimap

    def _set_ini_solid_state(self, uva0):
        raise NotImplementedError("Subclasses must implement this method")

    def _set_fin_solid_state(self, uva1):
        raise NotImplementedError("Subclasses must implement this method")

    def _set_ini_fluid_state(self, qp0):
        raise NotImplementedError("Subclasses must implement this method")

    def _set_fin_fluid_state(self, qp1):
        raise NotImplementedError("Subclasses must implement this method")

    @property
    def dt(self):
        return self.solid.dt

    @dt.setter
    def dt(self, value):
        self.solid.dt = value
        self.fluid.dt = value

    def set_ini_state(self, state):
        sl_state, fl_state = bv.chunk(state, (self.solid.state0.size, self.fluid.state0.size))
        self._set_ini_solid_state(sl_state)
        self._set_ini_fluid_state(fl_state)

    def set_fin_state(self, state):
        sl_state, fl_state = bv.chunk(state, (self.solid.state1.size, self.fluid.state1.size))
        self._set_fin_solid_state(sl_state)
        self._set_fin_fluid_state(fl_state)

    def set_control(self, control):
        self.control[:] = control
        for key, value in control.sub_items():
            self.fluid.control[key][:] = value

    def set_prop(self, prop):
        self.prop[:] = prop
        chunk_sizes = [model.prop.size for model in (self.solid, self.fluid)] + [1]
        prop_chunks = bv.chunk(self.prop, chunk_sizes)[:-1]
        for set_prop, sub in zip((self.solid.set_prop, self.fluid.set_prop), prop_chunks):
            set_prop(sub)
        self.solid._ymid = float(self.prop['ymid'][0])


class ExplicitFSIModel(BaseTransientFSIModel):
    """Explicit (staggered) coupling (``transient.py:821-920``): the solid at step n+1 is
    loaded with the fluid pressure of step n; the fluid sees the solid geometry of step n+1."""

    def _set_ini_solid_state(self, uva0):
        self.solid.set_ini_state(uva0)

    def _set_fin_solid_state(self, uva1):
        self.solid.set_fin_state(uva1)
        ndim = self.solid.residual.mesh().topology().dim()
        self._solid_area[:] = 2 * (
            self.prop['ymid'][0] - (self.solid.XREF + self.solid.state1.sub['u'])[1::ndim]
        )
        fl_control = self.fluid.control.copy()
        self.fsimap.map_solid_to_fluid(self._solid_area, fl_control.sub['area'][:])
        self.fluid.set_control(fl_control)

    def _set_ini_fluid_state(self, qp0):
        sl_control = self.solid.control.copy()
        sl_control['p'] = 0
        self.fluid.set_ini_state(qp0)
        self.fsimap.map_fluid_to_solid(qp0[1], sl_control.sub['p'])
        self.solid.set_control(sl_control)

    def _set_fin_fluid_state(self, qp1):
        self.fluid.set_fin_state(qp1)

    def assem_res(self):
        res_sl = self.solid.assem_res()
        res_fl = self.fluid.assem_res()
        return bv.concatenate((res_sl, res_fl))

    # --- linearisation (transient.py:873-896, 922-937).  The reference marks these methods
    # incomplete: they call fluid methods that do not exist (JaxModel.assem_dres_dstate0,
    # solve_dqp1_du1_solid) and build dstate1 from the solid's dstate0.  What follows is the
    # linearisation of the residual this class actually assembles:
    #   F_solid(uva1; uva0, p_solid = map(p0)),   F_fluid = (q1, p1) - Bernoulli(area(u1)).
    def _fluid_dqp_darea(self):
        """d(q, p)/d(area) of every fluid channel at the current fluid control, stacked as
        (n_fluid x ns_total) and (ns_total x ns_total) CSR blocks."""
        from ..equations import bernoulli_lin
        fl, r = self.fluid, self.fluid.residual
        s = np.atleast_2d(r.mesh())
        n_fluid, ns = s.shape
        area = np.asarray(fl.control['area']).reshape(n_fluid, ns)
        get = lambda key, dflt: np.broadcast_to(fl.prop[key], (n_fluid,)) if key in fl.prop \
            else np.full(n_fluid, dflt)
        rho, r_sep, lb = get('rho_air', 1.0), get('r_sep', 1.0), get('area_lb', 0.0)
        psub = np.broadcast_to(fl.control['psub'], (n_fluid,))
        psup = np.broadcast_to(fl.control['psup'], (n_fluid,))
        dq = sp.lil_matrix((n_fluid, n_fluid * ns))
        dps = []
        for k in range(n_fluid):
            dqk, dpk = bernoulli_lin.dqp_darea(r.kind, s[k], area[k], psub[k], psup[k], rho[k],
                                               r_sep[k], lb[k], r.idx_sep)
            dq[k, k * ns:(k + 1) * ns] = dqk
            dps.append(sp.csr_matrix(dpk))
        return dq.tocsr(), sp.block_diag(dps, format='csr')

    def assem_dres_dstate1(self):
        """Block Jacobian over (u, v, a, q, p) x (u1, v1, a1, q1, p1)."""
        sl = self.solid.assem_dres_dstate1()
        N = self.solid.state0['u'].size
        nq, npf = self.fluid.state0['q'].size, self.fluid.state0['p'].size
        dq_da, dp_da = self._fluid_dqp_darea()
        dfq_du = -(dq_da @ self._dflarea_dslu)        # F_q = q1 - q(area(u1))
        dfp_du = -(dp_da @ self._dflarea_dslu)
        Z = lambda m, n: sp.csr_matrix((m, n))
        rows = []
        for i in range(3):
            rows.append([sl.mats[3 * i + j] for j in range(3)] + [Z(N, nq), Z(N, npf)])
        rows.append([dfq_du, Z(nq, N), Z(nq, N), sp.identity(nq, format='csr'), Z(nq, npf)])
        rows.append([dfp_du, Z(npf, N), Z(npf, N), Z(npf, nq), sp.identity(npf, format='csr')])
        keys0 = tuple(self.state1.keys())
        return BlockMatrix([m for row in rows for m in row], (5, 5),
                           (keys0, tuple(f'state/{k}1' for k in keys0)))

    def assem_dres_dstate0(self):
        """Block Jacobian over (u, v, a, q, p) x (u0, v0, a0, q0, p0): the solid depends on p0
        through the pressure map (explicit coupling); the fluid residual does not depend on
        state0."""
        sl = self.solid.assem_dres_dstate0()
        dsl_dp = self.solid.assem_dres_dcontrol()
        N = self.solid.state0['u'].size
        nq, npf = self.fluid.state0['q'].size, self.fluid.state0['p'].size
        Z = lambda m, n: sp.csr_matrix((m, n))
        rows = []
        for i in range(3):
            rows.append([sl.mats[3 * i + j] for j in range(3)] +
                        [Z(N, nq), (dsl_dp.mats[i] @ self._dslp_dflp).tocsr()])
        rows.append([Z(nq, N)] * 3 + [Z(nq, nq), Z(nq, npf)])
        rows.append([Z(npf, N)] * 3 + [Z(npf, nq), Z(npf, npf)])
        keys0 = tuple(self.state1.keys())
        return BlockMatrix([m for row in rows for m in row], (5, 5),
                           (keys0, tuple(f'state/{k}0' for k in keys0)))

    def solve_dres_dstate1(self, b, dres_dstate1=None):
        """Solve dF/dstate1 x = b (``transient.py:922-937``): the solid block on the device, then
        the fluid rows by substitution (their diagonal blocks are identities)."""
        if dres_dstate1 is None:
            dres_dstate1 = self.assem_dres_dstate1()
        x = self.state0.copy()
        sl_labels = (self.solid.FORM_KEYS, self.solid.STATE1_KEYS)
        dsl = BlockMatrix(dres_dstate1.mats[0:3] + dres_dstate1.mats[5:8] + dres_dstate1.mats[10:13],
                          (3, 3), sl_labels)
        xs = self.solid.solve_dres_dstate1(dsl, self.solid.state0.copy(), b[:3])
        for key in ('u', 'v', 'a'):
            x[key][:] = xs[key]
        xu = np.asarray(xs['u'])
        x['q'][:] = b['q'] - dres_dstate1.sub['q', 'state/u1'] @ xu
        x['p'][:] = b['p'] - dres_dstate1.sub['p', 'state/u1'] @ xu
        return x

    def solve_state1(self, ini_state, options=None):
        """``transient.py:899-920``: solid Newton solve, area update, fluid solve."""
        self.set_fin_state(ini_state)
        uva1, solid_info = self.solid.solve_state1(ini_state[:3], options)
        self._set_fin_solid_state(uva1)
        qp1s, _ = self.fluid.solve_state1(ini_state[3:], options)
        step_info = solid_info
        return bv.concatenate([uva1, qp1s], labels=self.state1.labels), step_info

    # --- device-resident time loop used by forward.integrate -----------------------------
    def push_to_device(self):
        """Upload properties, state0 and fluid properties (once per ``integrate``)."""
        self.solid._ymid = float(self.prop['ymid'][0])
        self.solid._push_prop(self.solid._ymid)
        e = self.engine
        e.upload('fprop', self.fluid._fprop_block(), 0)
        # entries of the fluid area that no solid DOF maps to keep their host value
        # (default 1.0, residuals/fluid.py:298)
        e.upload('area', self.fluid.control['area'], 0)
        names = ('u0', 'v0', 'a0', 'q0', 'p0')
        for name, vec in zip(names, self.state0.vecs):
            e.upload(name, vec, 0)

    def device_integrate(self, dts, controls, options=None, store_states: bool = True):
        """
        Run ``len(dts)`` explicit-coupling steps on the device from the uploaded state0.

        controls : list of control BlockVectors (psub, psup)
        Returns (states, infos): host arrays (nsteps+1, state_size) and (nsteps+1, 4); with
        ``store_states=False`` states is (1, state_size): the final state only.
        """
        n_fluid = self.engine.n_fluid
        # (steps usually share one control object: expand each distinct object once)
        rows = {}
        for c in controls:
            if id(c) not in rows:
                rows[id(c)] = np.array([np.broadcast_to(c['psub'], (n_fluid,)),
                                        np.broadcast_to(c['psup'], (n_fluid,))])
        ctl = np.array([rows[id(c)] for c in controls])
        self.solid._retire_live_jacobian()
        hs, hi = self.engine.integrate(dts, ctl, options, store_states=store_states,
                                       store_info=True)
        if store_states:
            return hs[0].cpu().numpy(), hi[0].cpu().numpy()
        # no history asked for: only the final state (resident in state0 after the last swap)
        # comes back, as a one-row "history"
        eng = self.engine
        fin = np.concatenate([eng.download(k) for k in ('u0', 'v0', 'a0', 'q0', 'p0')])
        return fin[None, :], hi[0].cpu().numpy()

    def state_from_row(self, row: np.ndarray) -> BlockVector:
        state = self.state0.copy()
        state[:] = row
        return state


class ImplicitFSIModel(BaseTransientFSIModel):
    """
    Implicit (fixed-point) coupling (``transient.py:964-1033``): the solid at step n+1 is
    loaded with the fluid pressure of step n+1; solid and fluid are solved in turn until the
    coupled state stops changing.

    The reference's ``iterative_solve`` comes from the un-vendored ``nonlineq`` package and its
    ``assem_res`` call is stale (``transient.py:1001``), so the stopping rule is restated
    (documented in ``oracle/model.py`` ``ImplicitCoupledOracle``): iterate x <- G(x) with
    G = (solid solve with p(x), area update, fluid solve); stop when ||x_{k+1} - x_k||_2 <=
    absolute_tolerance, or <= relative_tolerance * ||x_1 - x_0||_2, or after
    ``maximum_iterations``.  Each solid/fluid solve runs on the device.
    """

    def _set_ini_fluid_state(self, qp0):
        self.fluid.set_ini_state(qp0)

    def _set_fin_fluid_state(self, qp1):
        sl_control = self.solid.control.copy()
        sl_control['p'] = 0
        self.fluid.set_fin_state(qp1)
        self.fsimap.map_fluid_to_solid(qp1[1], sl_control.sub['p'])
        self.solid.set_control(sl_control)

    def _set_ini_solid_state(self, uva0):
        self.solid.set_ini_state(uva0)

    _set_fin_solid_state = ExplicitFSIModel._set_fin_solid_state

    def assem_res(self):
        return bv.concatenate((self.solid.assem_res(), self.fluid.assem_res()))

    def solve_state1(self, ini_state, options=None):
        prm = dict(FIXEDPOINT_SOLVER_PRM)
        prm['maximum_iterations'] = 50
        newton_options = None
        if options:
            newton_options = options
            for key in ('fixedpoint_absolute_tolerance', 'fixedpoint_relative_tolerance',
                        'fixedpoint_maximum_iterations'):
                if key in options:
                    prm[key.replace('fixedpoint_', '')] = options[key]
            newton_options = {k: v for k, v in options.items() if not k.startswith('fixedpoint_')}
        x = ini_state.copy()
        self.set_fin_state(x)
        k, err0, abs_err, rel_err = 0, None, np.inf, np.inf
        while True:
            self._set_fin_fluid_state(x[3:])          # p(n+1) iterate onto the solid
            uva1, solid_info = self.solid.solve_state1(x[:3], newton_options)
            self._set_fin_solid_state(uva1)
            qp1, _ = self.fluid.solve_state1(x[3:], newton_options)
            x_new = bv.concatenate([uva1, qp1], labels=self.state1.labels)
            abs_err = (x_new - x).norm()
            if err0 is None:
                err0 = abs_err
            rel_err = abs_err / err0 if err0 > 0 else 0.0
            x = x_new
            k += 1
            if abs_err <= prm['absolute_tolerance'] or rel_err <= prm['relative_tolerance'] \
                    or k >= prm['maximum_iterations']:
                break
        self.set_fin_state(x)
        return x, {'num_iter': k, 'abs_err': abs_err, 'rel_err': rel_err}
