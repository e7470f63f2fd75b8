"""
Predefined 1D Bernoulli fluid residuals: mirror of
``/root/reference/src/femvf/residuals/fluid.py``.

The reference returns a JAX closure ``res(state, control, prop)``; here each residual
carries the prototype ``(state, control, prop)`` dictionaries (same keys, sizes and
defaults, ``fluid.py:109-128, 204-220, 296-311``) and a ``kind`` tag selecting the device
implementation in ``csrc/fluid.cuh``.  ``res`` evaluates the same closure through the
device kernel (see ``models.transient.JaxModel``); there is no CPU implementation.
"""

from __future__ import annotations

import numpy as np

from .base import JaxResidual

FLUID_AREA_RATIO_SEP = 0
FLUID_FIXED_SEP = 1
FLUID_SMOOTH_MIN_SEP = 2


class PredefinedFluidResidual(JaxResidual):
    """Predefined ``JaxResidual`` (``fluid.py:39-55``)."""

    KIND = -1

    def __init__(self, mesh: np.ndarray, *args, **kwargs):
        mesh = np.asarray(mesh, dtype=np.float64)
        res_args = self._make_residual(mesh, *args, **kwargs)
        super().__init__(None, res_args)
        self._mesh = mesh

    def mesh(self) -> np.ndarray:
        return self._mesh

    @property
    def kind(self) -> int:
        return self.KIND

    @property
    def idx_sep(self) -> int:
        return getattr(self, '_idx_sep', 0)

    def _make_residual(self, mesh, *args, **kwargs):
        raise NotImplementedError("Subclasses must implement this method")

    @staticmethod
    def _sizes(s):
        n_fluid = int(np.prod(s.shape[:-1]))
        return n_fluid, s.size


class BernoulliAreaRatioSep(PredefinedFluidResidual):
    """Separation where the area reaches ``r_sep * min(area)`` (``fluid.py:223-311``)."""

    KIND = FLUID_AREA_RATIO_SEP

    def _make_residual(self, mesh):
        n_fluid, n_total = self._sizes(mesh)
        state = {'q': np.ones(n_fluid), 'p': np.ones(n_total)}
        control = {'area': np.ones(n_total), 'psub': np.ones(n_fluid), 'psup': np.ones(n_fluid)}
        prop = {'rho_air': np.ones(n_fluid), 'r_sep': np.ones(n_fluid),
                'area_lb': np.zeros(n_fluid)}
        return state, control, prop


class BernoulliFixedSep(PredefinedFluidResidual):
    """Separation at a fixed mesh index (``fluid.py:58-128``)."""

    KIND = FLUID_FIXED_SEP

    def _make_residual(self, mesh, idx_sep=0):
        self._idx_sep = int(idx_sep)
        n_fluid, n_total = self._sizes(mesh)
        state = {'q': np.ones(n_fluid), 'p': np.ones(n_total)}
        control = {'area': np.ones(n_total), 'psub': np.ones(n_fluid), 'psup': np.ones(n_fluid)}
        prop = {'rho_air': np.ones(n_fluid)}
        return state, control, prop


class BernoulliSmoothMinSep(PredefinedFluidResidual):
    """Smooth-minimum separation (``fluid.py:131-220``)."""

    KIND = FLUID_SMOOTH_MIN_SEP

    def _make_residual(self, mesh):
        n_fluid, n_total = self._sizes(mesh)
        state = {'q': np.ones(n_fluid), 'p': np.ones(n_total)}
        control = {'area': np.ones(n_total), 'psub': np.ones(n_fluid), 'psup': np.ones(n_fluid)}
        prop = {'rho_air': np.ones(n_fluid), 'zeta_sep': np.ones(n_fluid),
                'zeta_min': np.ones(n_fluid)}
        return state, control, prop
