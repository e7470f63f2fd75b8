"""
Glottal-width state measures (``/root/reference/src/femvf/postprocess/solid.py:487-560``).

The glottal width of the 1D flow model is the minimum of the fluid area vector, which the
coupled model derives from the solid surface as ``2 (ymid - y)`` through the FSI map
(``models/transient.py:836-848``).  The measures below read it exactly as the reference does
(after ``set_fin_state``); for a whole ``StateFile`` the series is evaluated on the device in
one launch (``assem_series`` -> ``vf_glottal_width_series``).
"""

from __future__ import annotations

import numpy as np

from .base import BaseStateMeasure


class MeanGlottalWidth(BaseStateMeasure):
    """Minimum of the fluid area over all fluid DOFs (``solid.py:487-501``)."""

    def __init__(self, model):
        super().__init__(model)
        self.XREF = np.array(self.model.solid.XREF[:])

    def assem(self, state, control, prop):
        return np.min(self.model.fluid.control['area'])

    def assem_series(self, f, ns):
        """Device evaluation for the time indices ``ns`` of a StateFile; ``None`` when the
        model has no device engine with an FSI map (the caller then loops on the host)."""
        model = self.model
        engine = getattr(model, 'engine', None)
        if engine is None or not getattr(engine, 'n_fluid', 0) or not hasattr(model, 'push_to_device'):
            return None
        if len(ns) == 0:
            return np.zeros(0)
        model.push_to_device()      # ymid and the base fluid area of the current properties
        u_hist = np.stack([np.asarray(f.get_state(ii).sub['u']) for ii in ns])
        return engine.glottal_width_series(u_hist)


class MidpointGlottalWidth(BaseStateMeasure):
    """Minimum area of the middle fluid channel(s) (``solid.py:504-528``)."""

    def __init__(self, model):
        super().__init__(model)
        self.XREF = np.array(self.model.solid.XREF[:])

    def assem(self, state, control, prop):
        fluid = self.model.fluid
        shape_fluid = np.shape(fluid.residual.mesh())[:-1]
        area = np.asarray(fluid.control['area']).reshape(*shape_fluid, -1)
        if area.ndim == 1:          # a single channel stored without the leading axis
            area = area[None, :]
        assert area.ndim == 2
        n = area.shape[0]
        if n % 2 == 1:
            # (the reference's odd branch builds a float index, solid.py:519-520, and cannot
            # run; the middle channel is what it describes)
            idxs_mid = [(n - 1) // 2]
        else:
            idxs_mid = [n // 2 - 1, n // 2]
        mins = [np.min(area[ii, :]) for ii in idxs_mid]
        return sum(mins) / len(mins)


class MinGlottalWidthFromSolid(BaseStateMeasure):
    """``min 2 (ymid - y)`` over ALL solid vertices (``solid.py:531-549``)."""

    def __init__(self, model):
        super().__init__(model)
        self.XREF = np.array(self.model.solid.XREF[:])

    def assem(self, state, control, prop):
        xcur = self.XREF.reshape(-1) + self.model.state1.sub['u'][:]
        ndim = self.model.solid.residual.mesh().topology().dim()
        widths = 2 * (self.model.prop['ymid'] - xcur[1::ndim])
        return np.min(widths)
