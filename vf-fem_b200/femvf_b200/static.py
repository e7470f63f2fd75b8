"""
Static problems: mirror of the solid part of ``/root/reference/src/femvf/static.py``.

``static_solid_configuration`` (``static.py:68-168``) solves F_u(u; u0 == u1, v0 = a0 = 0) = 0,
i.e. K(u) + follower pressure + contact = 0, with the device Newton/GMRES loop in static
mode (no inertia or damping blocks; ``csrc/node_assembly.cuh`` ``is_static``).
"""

from __future__ import annotations

from typing import Any

from . import blockvec as bv
from .models import transient

Info = dict


def static_solid_configuration(model: transient.FenicsModel, control: bv.BlockVector,
                               prop: bv.BlockVector, state=None, solver: str = 'manual',
                               options=None):
    """Return the static state for a solid model (``static.py:68-168``)."""
    if not isinstance(model, transient.BaseTransientModel):
        raise TypeError(f"Unknown `model` type {type(model)}")
    if solver not in ('manual', 'automatic'):
        raise ValueError(f"Unknown `solver`: '{solver}'")

    state_n = model.state0.copy()
    if state is None:
        state_n[:] = 0.0
    else:
        state_n[:] = state

    model.set_control(control)
    model.set_prop(prop)

    zero_state = model.state1.copy()
    zero_state[:] = 0
    model.set_ini_state(zero_state)
    guess = zero_state.copy()
    guess['u'][:] = state_n['u']
    model.set_fin_state(guess)

    model._push_all()
    model._retire_live_jacobian()
    e, m = model.engine, model._member
    grid = model._grid_solver(static=True)
    if grid is not None:
        from .solverconst import DEFAULT_NEWTON_SOLVER_PRM
        ginfo = grid.solve_state1(1.0, dict(options or DEFAULT_NEWTON_SOLVER_PRM), is_static=True)
        state_n['u'][:] = e.download('u1', m)
        return state_n, {k: ginfo[k] for k in ('num_iter', 'abs_err', 'rel_err')}
    e.solve_state1(1.0, m, 1, options, is_static=True)
    state_n['u'][:] = e.download('u1', m)
    raw = e.download('info', m)
    info = {'num_iter': int(raw[0]), 'abs_err': float(raw[1]), 'rel_err': float(raw[2])}
    return state_n, info
