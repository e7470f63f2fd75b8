"""
State measures and their time series (interface of ``/root/reference/src/femvf/postprocess/
base.py``).

* a *state measure* maps one ``(state, control, prop)`` to a number or array
  (``base.py:18-64``): calling it first pushes the non-``None`` arguments through the model's
  setters (prop, control, then the state as both final and initial state), then evaluates;
* a *state-history measure* maps a ``StateFile`` to a result (``base.py:92-113``);
  ``TimeSeries`` evaluates a state measure at every stored time index (``base.py:138-161``).

``TimeSeries`` has a device path: a measure that defines ``assem_series(f, ns)`` is evaluated
for all requested time indices in one batched kernel launch (the glottal-width measures in
``postprocess/solid.py`` do, through ``vf_glottal_width_series``); anything else takes the
reference's per-state loop.
"""

from __future__ import annotations

from typing import Iterable, Optional

import numpy as np


class BaseStateMeasure:
    """Post-process one ``(state, control, prop)`` (``base.py:18-64``)."""

    def __init__(self, model, **kwargs):
        self._model = model

    def __call__(self, state=None, control=None, prop=None):
        model = self.model
        if prop is not None:
            model.set_prop(prop)
        if control is not None:
            model.set_control(control)
        if state is not None:
            model.set_fin_state(state)
            model.set_ini_state(state)
        return self.assem(state, control, prop)

    @property
    def model(self):
        return self._model

    def assem(self, state, control, prop):
        raise NotImplementedError("Method must be implemented by subclasses")


class BaseDerivedStateMeasure(BaseStateMeasure):
    """A measure computed from another state measure (``base.py:67-89``)."""

    def __init__(self, func: BaseStateMeasure):
        self._func = func
        super().__init__(func.model)

    @property
    def func(self):
        return self._func


class BaseStateHistoryMeasure:
    """Post-process a state history held in a ``StateFile`` (``base.py:92-113``)."""

    def __init__(self, model, **kwargs):
        self._model = model

    def __call__(self, f, **kwargs):
        return self.assem(f, **kwargs)

    @property
    def model(self):
        return self._model

    def assem(self, f, **kwargs):
        raise NotImplementedError("Method must be implemented by subclasses")


class BaseDerivedStateHistoryMeasure(BaseStateHistoryMeasure):
    """A history measure built on a state measure (``base.py:116-135``)."""

    def __init__(self, func: BaseStateMeasure):
        super().__init__(func.model)
        self._func = func

    @property
    def func(self):
        return self._func


class TimeSeries(BaseDerivedStateHistoryMeasure):
    """Time series of a state measure over a ``StateFile`` (``base.py:138-161``)."""

    def __call__(self, f, ns: Optional[Iterable] = None):
        return self.assem(f, ns=ns)

    def assem(self, f, ns: Optional[Iterable] = None):
        ns = list(range(f.size)) if ns is None else list(ns)
        prop = f.get_prop()
        self.func.model.set_prop(prop)
        batched = getattr(self.func, 'assem_series', None)
        if batched is not None:
            out = batched(f, ns)
            if out is not None:
                return np.asarray(out)
        return np.array([self.func(f.get_state(ii), f.get_control(ii), prop=None) for ii in ns])


class TimeSeriesStats(BaseDerivedStateHistoryMeasure):
    """Statistics of the time series of a state measure (``base.py:164-200``); calling it
    returns the mean."""

    def __init__(self, func: BaseStateMeasure):
        super().__init__(func)
        self._ts = TimeSeries(func)

    @property
    def ts(self):
        return self._ts

    def assem(self, f, ns: Optional[Iterable] = None):
        return self.mean(f, ns=ns)

    def _reduce(self, op, f, ns):
        return op(self.ts(f, ns=ns), axis=0)

    def max(self, f, ns: Optional[Iterable] = None):
        return self._reduce(np.max, f, ns)

    def min(self, f, ns: Optional[Iterable] = None):
        return self._reduce(np.min, f, ns)

    def mean(self, f, ns: Optional[Iterable] = None):
        return self._reduce(np.mean, f, ns)

    def std(self, f, ns: Optional[Iterable] = None):
        return self._reduce(np.std, f, ns)
