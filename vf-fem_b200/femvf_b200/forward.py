"""
Integrate models in time: mirror of ``/root/reference/src/femvf/forward.py``.

``integrate`` / ``integrate_steps`` / ``integrate_step`` / ``append_step_result`` keep the
reference signatures and semantics (``forward.py:22-186, 247-284``).  For the explicitly
coupled device model the step loop runs inside one persistent CUDA kernel per chunk of
``NCHUNK`` steps (``ExplicitFSIModel.device_integrate``) instead of one Python iteration per
step; the states written to the ``StateFile`` and the returned final state / info are the
same as those of the per-step loop, which remains available for any other model.

Uses CGS (cm-g-s) units unless otherwise stated
"""

from __future__ import annotations

from typing import Any, Optional

import numpy as np

from . import blockvec as bv
from . import statefile as sf
from .models.transient import BaseTransientModel, ExplicitFSIModel

Options = dict
Info = dict


def integrate(
    model: BaseTransientModel,
    f: Optional[sf.StateFile],
    ini_state: bv.BlockVector,
    controls: list,
    prop: bv.BlockVector,
    times,
    idx_meas: Optional[np.ndarray] = None,
    newton_solver_prm: Optional[dict] = None,
    write: bool = True,
    use_tqdm: bool = False,
):
    """Integrate the model over a set of time instances (``forward.py:22-102``)."""
    if idx_meas is None:
        idx_meas = np.array([])

    if len(times) < 1:
        raise ValueError("There must be at least 1 time integration point.")
    if times[-1] <= times[0]:
        raise ValueError(
            "The final time point must be greater or equal to the initial one."
            f"The input initial/final times were {times[0]}/{times[-1]}"
        )

    if write:
        f.init_layout()
        append_step_result(f, ini_state, controls[0], times[0],
                           {'num_iter': 0, 'abs_err': 0, 'rel_err': 0})
        f.append_prop(prop)
        if 0 in idx_meas:
            f.append_meas_index(0)

    fin_state, step_info = integrate_steps(
        model, f, ini_state, controls, prop, times, idx_meas=idx_meas,
        newton_solver_prm=newton_solver_prm, write=write, use_tqdm=use_tqdm,
    )
    return fin_state, step_info


def integrate_extend(model, f: sf.StateFile, controls, times, idx_meas=None,
                     newton_solver_prm=None, write: bool = True):
    """Continue a stored simulation from its last state (``forward.py:105-136``)."""
    prop = f.get_prop()
    _controls = controls[1:] if len(controls) > 1 else controls
    N = f.size
    ini_state = f.get_state(N - 1)
    ini_time = f.get_time(N - 1)
    times = np.asarray(times, dtype=float) + ini_time
    return integrate_steps(model, f, ini_state, _controls, prop, times, idx_meas=idx_meas,
                           newton_solver_prm=newton_solver_prm, write=write)


def integrate_steps(
    model: BaseTransientModel,
    f: Optional[sf.StateFile],
    ini_state: bv.BlockVector,
    controls: list,
    prop: bv.BlockVector,
    times,
    idx_meas: Optional[np.ndarray] = None,
    newton_solver_prm: Optional[dict] = None,
    write: bool = True,
    use_tqdm: bool = False,
):
    """See ``integrate`` (``forward.py:139-186``)."""
    if idx_meas is None:
        idx_meas = np.array([])

    state0 = ini_state
    model.set_prop(prop)
    step_info = {}
    times = np.asarray(times, dtype=float)

    # the in-kernel time loop runs one CTA per simulation: meshes that do not fit one CTA take the
    # per-step path below, whose solid solve is the whole-GPU Newton (gridsolve.py)
    if isinstance(model, ExplicitFSIModel) and model.solid._grid_solver() is None:
        return _integrate_steps_device(model, f, state0, controls, times, idx_meas,
                                       newton_solver_prm, write)

    for n, (time0, time1) in enumerate(zip(times[:-1], times[1:])):
        control1 = controls[min(n, len(controls) - 1)]
        dt = time1 - time0
        state1, step_info = integrate_step(model, state0, control1, prop, dt,
                                           options=newton_solver_prm)
        if write:
            append_step_result(f, state1, control1, time1, step_info)
            if n in idx_meas:
                f.append_meas_index(n)
        state0 = state1
    return state0, step_info


def _integrate_steps_device(model: ExplicitFSIModel, f, state0, controls, times, idx_meas,
                            newton_solver_prm, write):
    """Device-resident step loop; state history comes back once per chunk."""
    nsteps = len(times) - 1
    step_info = {}
    if nsteps <= 0:
        return state0, step_info
    model.set_ini_state(state0)
    model.push_to_device()
    nchunk = f.NCHUNK if (write and f is not None) else max(nsteps, 1)
    dts_all = np.diff(times)
    n0 = 0
    fin_state = state0
    while n0 < nsteps:
        n1 = min(n0 + nchunk, nsteps)
        # control index min(n, len-1) per step (forward.py:170), relative to this chunk
        chunk_controls = [controls[min(n, len(controls) - 1)] for n in range(n0, n1)]
        states, infos = model.device_integrate(dts_all[n0:n1], chunk_controls, newton_solver_prm,
                                               store_states=bool(write))
        for k in range(n1 - n0):
            n = n0 + k
            info = {'num_iter': int(infos[k + 1, 0]), 'abs_err': float(infos[k + 1, 1]),
                    'rel_err': float(infos[k + 1, 2])}
            if write:
                state1 = model.state_from_row(states[k + 1])
                append_step_result(f, state1, chunk_controls[k], times[n + 1], info)
                if n in idx_meas:
                    f.append_meas_index(n)
            step_info = info
        fin_state = model.state_from_row(states[-1])
        n0 = n1
    # leave the host mirrors of the model consistent with the final device state
    model.set_control(controls[min(nsteps - 1, len(controls) - 1)])
    model.set_ini_state(fin_state)
    model.set_fin_state(fin_state)
    return fin_state, step_info


def integrate_step(model: BaseTransientModel, ini_state, control, prop, dt: float,
                   set_prop: bool = False, options: Options = None):
    """Integrate a model over a single time step (``forward.py:247-268``)."""
    model.dt = dt
    model.set_ini_state(ini_state)
    model.set_control(control)
    if set_prop:
        model.set_prop(prop)
    fin_state, step_info = model.solve_state1(ini_state, options=options)
    return fin_state, step_info


def append_step_result(f: sf.StateFile, state, control, time: float, step_info: Info):
    """Append the result of an integration step to a statefile (``forward.py:271-284``)."""
    f.append_state(state)
    f.append_control(control)
    f.append_time(time)
    f.append_solver_info(step_info)
