"""
State history container: mirror of ``/root/reference/src/femvf/statefile.py``.

Same dataset layout (``statefile.py:163-270``): ``/time``, ``/meas_indices``,
``/mesh/solid/{coordinates,connectivity,dim}``, ``/dofmap/CG1``, ``/state/<name>``,
``/control/<name>``, ``/properties/<name>``, ``/solver_info/{num_iter,rel_err,abs_err}``; row 0
is the initial state.  Backed by h5py when available, otherwise by ``h5lite`` (one ``.npz``
with the same keys).
"""

from __future__ import annotations

from typing import Any, Union

import numpy as np

from . import blockvec as bv

try:  # pragma: no cover - h5py is absent in the build image
    import h5py as _h5
    _File, _Group = _h5.File, _h5.Group
    HAVE_H5PY = True
except ImportError:
    from . import h5lite as _h5
    _File, _Group = _h5.File, _h5.Group
    HAVE_H5PY = False


class StateFile:
    """History of states of a transient model simulation (``statefile.py:21-105``)."""

    def __init__(self, model, fname: Union[str, Any], mode: str = 'r', NCHUNK: int = 100,
                 **kwargs):
        self.model = model
        if isinstance(fname, str):
            self.file = _File(fname, mode=mode, **kwargs)
        elif isinstance(fname, _Group):
            self.file = fname
        else:
            raise TypeError(f"`fname` must be `str` or `h5py.Group` not {type(fname)}")
        self.NCHUNK = NCHUNK
        self.init_layout()

    def __enter__(self):
        return self

    def __exit__(self, type, value, traceback):
        self.file.close()

    def keys(self):
        return self.file.keys()

    def __getitem__(self, name):
        return self.file[name]

    def __setitem__(self, name, value):
        self.file[name] = value

    def __len__(self):
        return self.size

    def close(self):
        self.file.close()

    @property
    def size(self):
        """Number of states in the file = number of stored time indices."""
        if 'time' in self.file:
            return self.file['time'].shape[0]
        return 0

    @property
    def variable_controls(self):
        return self.num_controls > 1

    @property
    def num_controls(self):
        num = 1
        control_group = self.file['control']
        for key in self.model.control.keys():
            num = max(control_group[key].shape[0], num)
        return num

    ## layout (statefile.py:163-270)
    def init_layout(self):
        self.file.require_dataset('time', (self.size,), maxshape=(None,),
                                  chunks=(self.NCHUNK,), dtype=np.float64, exact=False)
        if 'meas_indices' not in self.file:
            self.file.create_dataset('meas_indices', (0,), maxshape=(None,),
                                     chunks=(self.NCHUNK,), dtype=np.intp)
        self.init_mesh()
        self.init_state()
        self.init_control()
        self.init_prop()
        self.init_solver_info()

    def init_mesh(self):
        solid = self.model.solid
        coords = solid.residual.mesh().coordinates()
        cells = solid.residual.mesh().cells()
        self.file.require_dataset('mesh/solid/coordinates', coords.shape, data=coords,
                                  dtype=np.float64)
        self.file.require_dataset('mesh/solid/connectivity', cells.shape, data=cells,
                                  dtype=np.intp)
        self.file.require_dataset('mesh/solid/dim', (),
                                  data=solid.residual.mesh().topology().dim(), dtype=np.intp)
        dofmap = solid.residual.form['state/u0'].function_space().dofmap()
        dofmap_array = np.array([dofmap.cell_dofs(idx) for idx in range(cells.shape[0])])
        self.file.require_dataset('dofmap/CG1', dofmap_array.shape, data=dofmap_array,
                                  dtype=np.intp)

    def _init_series(self, group_name, bvec):
        group = self.file.require_group(group_name)
        for name, ndof in zip(bvec.labels[0], bvec.bshape[0]):
            group.require_dataset(name, (self.size, ndof), maxshape=(None, ndof),
                                  chunks=(self.NCHUNK, ndof), dtype=np.float64)

    def init_state(self):
        self._init_series('state', self.model.state0)

    def init_control(self):
        self._init_series('control', self.model.control)

    def init_prop(self):
        group = self.file.require_group('properties')
        bvec = self.model.prop
        for name, ndof in zip(bvec.labels[0], bvec.bshape[0]):
            group.require_dataset(name, (ndof,), dtype=np.float64)

    def init_solver_info(self):
        group = self.file.require_group('solver_info')
        for key in ['num_iter', 'rel_err', 'abs_err']:
            group.require_dataset(key, (self.size,), dtype=np.float64, maxshape=(None,),
                                  chunks=(self.NCHUNK,))

    ## appending (statefile.py:273-339)
    def append_state(self, state: bv.BlockVector):
        group = self.file['state']
        for name, value in state.items():
            dset = group[name]
            dset.resize(dset.shape[0] + 1, axis=0)
            dset[-1, :] = value

    def append_control(self, control: bv.BlockVector):
        group = self.file['control']
        for name, value in control.items():
            dset = group[name]
            dset.resize(dset.shape[0] + 1, axis=0)
            dset[-1] = value

    def append_prop(self, properties: bv.BlockVector):
        group = self.file['properties']
        for name, value in properties.items():
            group[name][:] = value

    def append_time(self, time: float):
        dset = self.file['time']
        dset.resize(dset.shape[0] + 1, axis=0)
        dset[-1] = time

    def append_meas_index(self, index: int):
        dset = self.file['meas_indices']
        dset.resize(dset.shape[0] + 1, axis=0)
        dset[-1] = index

    def append_solver_info(self, solver_info: dict):
        group = self.file['solver_info']
        for key, dset in group.items():
            dset.resize(dset.shape[0] + 1, axis=0)
            dset[-1] = solver_info[key] if key in solver_info else np.nan

    ## reading (statefile.py:342-422)
    def get_time(self, n: int) -> float:
        return self.file['time'][n]

    def get_times(self) -> np.ndarray:
        return self.file['time'][:]

    def get_meas_indices(self) -> np.ndarray:
        return self.file['meas_indices'][:]

    def get_state(self, n: int) -> bv.BlockVector:
        state = self.model.state0.copy()
        for key, vec in state.items():
            vec[:] = self.file[f'state/{key}'][n]
        return state

    def get_control(self, n: int) -> bv.BlockVector:
        control = self.model.control.copy()
        num_controls = self.file[f'control/{control.keys()[0]}'].shape[0]
        n = min(n, num_controls - 1)
        for key, vec in control.items():
            vec[:] = self.file[f'control/{key}'][n]
        return control

    def get_prop(self) -> bv.BlockVector:
        properties = self.model.prop.copy()
        for name, vec in zip(properties.keys(), properties.blocks):
            vec[:] = self.file[f'properties/{name}'][:]
        return properties

    def get_solver_info(self, n) -> dict:
        group = self.file['solver_info']
        return {key: group[key][n] for key in group.keys()}

    def set_state(self, n: int, state: bv.BlockVector):
        for label, value in zip(state.keys(), state.vecs):
            self.file[f'state/{label}'][n] = value
