"""
ORACLE (test infrastructure, NOT product code) -- explicit-coupling time step,
Newton loop and time integration of the Kelvin-Voigt solid + Bernoulli fluid.

PARITY UNPINNED (see ``oracle/fem.py``).  ``nonlineq.newton_solve`` and PETSc LU are
un-vendored third-party code; their semantics are restated from the call sites:
  forward.py:139-186, 247-268          integrate_steps / integrate_step
  models/transient.py:441-491          FenicsModel.solve_state1 / solve_dres_dstate1
  models/transient.py:833-862, 899-920 ExplicitFSIModel staggering
  models/transient.py:516-583          NodalContactModel
  models/fsi.py:66-70                  FSIMap gather/scatter
  solverconst.py:1-6                   Newton defaults
  static.py:68-168                     static_solid_configuration
  meshutils.py:295-334                 sort_vertices_by_nearest_neighbours
  load.py:283-293                      arclength of the 1D fluid mesh

Newton semantics chosen here (not determinable from the reference):
iterate k = 0, 1, ...: r_k = F_u(u_k) with BC rows zeroed; abs_err = ||r_k||_2,
rel_err = abs_err / ||r_0||_2 (0 if ||r_0|| = 0); stop when abs_err <= atol or
rel_err <= rtol or k == max_iter; otherwise u_{k+1} = u_k - J(u_k)^{-1} r_k.
v1, a1 follow from the Newmark relations (App. C, Q2).  LU = ``scipy.sparse.linalg.splu``.
"""

from __future__ import annotations

import numpy as np
import scipy.sparse.linalg as spla

from . import fem, fluid as ofl

DEFAULT_NEWTON_SOLVER_PRM = {  # solverconst.py:1-6
    'absolute_tolerance': 1e-8,
    'relative_tolerance': 1e-10,
    'maximum_iterations': 50,
}


def sort_vertices_by_nearest_neighbours(x, origin=None):
    """Greedy nearest-neighbour ordering starting from the point closest to ``origin``."""
    x = np.asarray(x, dtype=float)
    origin = np.zeros(x.shape[-1]) if origin is None else origin
    idx = [int(np.argmin(np.linalg.norm(x - origin, axis=-1)))]
    while len(idx) < x.shape[0]:
        dist = np.sum((x - x[idx[-1]]) ** 2, axis=-1) ** 0.5
        dist[idx] = np.nan
        idx.append(int(np.nanargmin(dist)))
    return np.array(idx)


def interface_from_edges(coords, edges):
    """1D fluid mesh ``s`` and ordered vertex ids from a set of surface edges."""
    verts = np.unique(np.asarray(edges).reshape(-1))
    xs = coords[verts]
    order = sort_vertices_by_nearest_neighbours(xs)
    xs = xs[order]
    dx = xs[1:] - xs[:-1]
    s = np.concatenate([[0.0], np.cumsum(np.sqrt(dx[:, 0] ** 2 + dx[:, 1] ** 2))])
    return s, verts[order]


class SolidOracle:
    """Newmark Kelvin-Voigt solid: residual, Jacobian and Newton solve (FenicsModel)."""

    def __init__(self, prob: fem.SolidProblem, contact: bool = False, membrane: bool = False):
        self.prob = prob
        self.contact = contact
        self.membrane = membrane

    def _contact_args(self, prop):
        return dict(ncontact=np.asarray(prop['ncontact'], dtype=float),
                    ycontact=float(np.ravel(prop['ycontact'])[0]),
                    kcontact=float(np.ravel(prop['kcontact'])[0]))

    def _membrane_args(self, prop):
        if not self.membrane:
            return None
        return {k: prop[k] for k in ('emod_membrane', 'nu_membrane', 'th_membrane')}

    def tcontact(self, u1, prop):
        if not self.contact:
            return None
        c = self._contact_args(prop)
        return fem.contact_traction(self.prob.coords, u1, c['ncontact'], c['ycontact'], c['kcontact'])

    def res(self, u1, state0, dt, prop, p1):
        u0, v0, a0 = state0
        return fem.assemble_res_u(self.prob, u1, u0, v0, a0, dt, prop, p1,
                                  tcontact=self.tcontact(u1, prop),
                                  membrane=self._membrane_args(prop))

    def jac(self, u1, dt, prop, p1):
        return fem.assemble_jac_uu(self.prob, u1, dt, prop, p1,
                                   contact=self._contact_args(prop) if self.contact else None,
                                   membrane=self._membrane_args(prop))

    def solve_state1(self, state0, dt, prop, p1, options=None, u_guess=None):
        opts = dict(DEFAULT_NEWTON_SOLVER_PRM)
        if options:
            opts.update(options)
        u0, v0, a0 = state0
        u = np.array(u0 if u_guess is None else u_guess, dtype=float)
        k = 0
        r0 = None
        while True:
            r = self.res(u, state0, dt, prop, p1)
            abs_err = float(np.linalg.norm(r))
            if r0 is None:
                r0 = abs_err
            rel_err = abs_err / r0 if r0 > 0 else 0.0
            if abs_err <= opts['absolute_tolerance'] or rel_err <= opts['relative_tolerance'] \
                    or k >= opts['maximum_iterations']:
                break
            J = self.jac(u, dt, prop, p1)
            du = spla.splu(J.tocsc()).solve(r)
            u = u - du
            k += 1
        v1 = fem.newmark_v(u, u0, v0, a0, dt)
        a1 = fem.newmark_a(u, u0, v0, a0, dt)
        return (u, v1, a1), {'num_iter': k, 'abs_err': abs_err, 'rel_err': rel_err}


class CoupledOracle:
    """ExplicitFSIModel: p(n) -> solid, solve solid, area(n+1) -> Bernoulli."""

    def __init__(self, solid: SolidOracle, s, solid_dofs, fluid_dofs=None,
                 fluid_kind='area_ratio', fluid_kwargs=None):
        self.solid = solid
        self.s = np.asarray(s, dtype=float)
        self.solid_dofs = np.asarray(solid_dofs, dtype=np.int64).reshape(-1)
        ns = self.s.shape[-1]
        if fluid_dofs is None:
            shape = self.s.shape[:-1]
            fluid_dofs = (np.ones(shape + (1,), dtype=int) * np.arange(ns)).reshape(-1) \
                if shape else np.arange(ns)
        self.fluid_dofs = np.asarray(fluid_dofs, dtype=np.int64).reshape(-1)
        self.fluid_kind = fluid_kind
        self.fluid_kwargs = fluid_kwargs or {}

    def solid_pressure(self, p_fluid):
        p1 = np.zeros(self.solid.prob.nn)
        p1[self.solid_dofs] = np.ravel(p_fluid)[self.fluid_dofs]  # fsi.py:66-67
        return p1

    def fluid_area(self, u1, ymid):
        prob = self.solid.prob
        d = prob.d
        area_solid = 2 * (ymid - (prob.coords.reshape(-1) + u1)[1::d])  # transient.py:841-844
        area = np.ones(self.s.size)  # fluid control default (fluid.py:298)
        area[self.fluid_dofs] = area_solid[self.solid_dofs]  # fsi.py:69-70
        return area.reshape(self.s.shape)

    def fluid_qp(self, area, control, prop):
        shape = self.s.shape[:-1] + (1,)
        psub = np.reshape(control['psub'], shape)
        psup = np.reshape(control['psup'], shape)
        rho = np.reshape(prop['rho_air'], shape)
        if self.fluid_kind == 'area_ratio':
            q, p = ofl.bernoulli_area_ratio_sep(
                self.s, area, psub, psup, rho,
                np.reshape(prop['r_sep'], shape), np.reshape(prop['area_lb'], shape))
        elif self.fluid_kind == 'fixed':
            q, p = ofl.bernoulli_fixed_sep(self.s, area, psub, psup, rho,
                                           self.fluid_kwargs.get('idx_sep', 0))
        elif self.fluid_kind == 'smooth_min':
            q, p = ofl.bernoulli_smooth_min_sep(
                self.s, area, psub, psup, rho,
                np.reshape(prop['zeta_min'], shape), np.reshape(prop['zeta_min'], shape))
        else:
            raise ValueError(self.fluid_kind)
        return np.reshape(q, -1), np.reshape(p, -1)

    def step(self, state0, control, prop, dt, options=None):
        u0, v0, a0, q0, p0 = state0
        p1 = self.solid_pressure(p0)
        (u1, v1, a1), info = self.solid.solve_state1((u0, v0, a0), dt, prop, p1, options)
        area = self.fluid_area(u1, float(np.ravel(prop['ymid'])[0]))
        q1, pf1 = self.fluid_qp(area, control, prop)
        info = dict(info)
        info['area'] = area
        return (u1, v1, a1, q1, pf1), info

    def integrate(self, ini_state, controls, prop, times, options=None):
        """Mirror of forward.integrate_steps; returns the state history (row 0 = initial)."""
        times = np.asarray(times, dtype=float)
        if len(times) < 1:
            raise ValueError("There must be at least 1 time integration point.")
        if times[-1] <= times[0]:
            raise ValueError("The final time point must be greater or equal to the initial one.")
        state = tuple(np.array(x, dtype=float) for x in ini_state)
        hist = [state]
        infos = [{'num_iter': 0, 'abs_err': 0, 'rel_err': 0}]
        for n in range(len(times) - 1):
            control = controls[min(n, len(controls) - 1)]
            dt = times[n + 1] - times[n]
            state, info = self.step(state, control, prop, dt, options)
            hist.append(state)
            infos.append(info)
        return hist, infos


def static_solid_configuration(solid: SolidOracle, prop, p1, options=None, u_guess=None):
    """
    Static equilibrium F_u(u; u0 == u1, v0 = a0 = 0) = 0 (static.py:68-168).

    With u0 == u1 and v0 = a0 = 0 the Newmark velocity and acceleration vanish for
    any dt, so the residual is K(u) + pressure + contact and the Jacobian drops the
    mass and damping blocks.
    """
    opts = dict(DEFAULT_NEWTON_SOLVER_PRM)
    if options:
        opts.update(options)
    prob = solid.prob
    N = prob.N
    zero = np.zeros(N)
    u = np.zeros(N) if u_guess is None else np.array(u_guess, dtype=float)
    static_prop = dict(prop)
    static_prop['rho'] = np.zeros(prob.ne)
    static_prop['eta'] = np.zeros(prob.ne)
    static_prop.pop('rayleigh_m', None)
    static_prop.pop('rayleigh_k', None)
    k = 0
    r0 = None
    while True:
        r = solid.res(u, (u, zero, zero), 1.0, static_prop, p1)
        abs_err = float(np.linalg.norm(r))
        if r0 is None:
            r0 = abs_err
        rel_err = abs_err / r0 if r0 > 0 else 0.0
        if abs_err <= opts['absolute_tolerance'] or rel_err <= opts['relative_tolerance'] \
                or k >= opts['maximum_iterations']:
            break
        J = solid.jac(u, 1.0, static_prop, p1)
        u = u - spla.splu(J.tocsc()).solve(r)
        k += 1
    return u, {'num_iter': k, 'abs_err': abs_err, 'rel_err': rel_err}


class ImplicitCoupledOracle(CoupledOracle):
    """
    ImplicitFSIModel (models/transient.py:964-1033): fixed-point iteration between the solid
    (loaded with the CURRENT iterate of the fluid pressure) and the fluid.  ``nonlineq
    .iterative_solve`` is un-vendored; the stopping rule chosen here: iterate x <- G(x); stop
    when ||x_{k+1} - x_k||_2 <= abs_tol, or <= rel_tol * ||x_1 - x_0||_2, or after max_iter
    (solverconst.py:14: abs 1e-8, rel 1e-11).
    """

    def step(self, state0, control, prop, dt, options=None, abs_tol=1e-8, rel_tol=1e-11,
             max_iter=50):
        u0, v0, a0, q0, p0 = state0
        x = [np.array(v, dtype=float) for v in state0]
        k, err0 = 0, None
        while True:
            p1 = self.solid_pressure(x[4])
            (u1, v1, a1), info = self.solid.solve_state1((u0, v0, a0), dt, prop, p1, options,
                                                         u_guess=x[0])
            area = self.fluid_area(u1, float(np.ravel(prop['ymid'])[0]))
            q1, pf1 = self.fluid_qp(area, control, prop)
            x_new = [u1, v1, a1, q1, pf1]
            abs_err = float(np.sqrt(sum(np.sum((a - b) ** 2) for a, b in zip(x_new, x))))
            if err0 is None:
                err0 = abs_err
            rel_err = abs_err / err0 if err0 > 0 else 0.0
            x = x_new
            k += 1
            if abs_err <= abs_tol or rel_err <= rel_tol or k >= max_iter:
                break
        return tuple(x), {'num_iter': k, 'abs_err': abs_err, 'rel_err': rel_err, 'area': area}
