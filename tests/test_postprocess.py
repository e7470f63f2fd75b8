"""Post-processing measures: host semantics on CPU, device series on GPU."""

import numpy as np
import pytest

from femvf_b200 import meshgen, statefile as sf
from femvf_b200.load import load_fsi_model
from femvf_b200.postprocess import base as ppbase, solid as ppsolid
from femvf_b200.residuals import solid as slr, fluid as flr


def _model():
    return load_fsi_model(meshgen.m5_cb_mesh(0.05), slr.KelvinVoigt, flr.BernoulliAreaRatioSep,
                          {'dirichlet_bcs': {'state/u1': [(np.zeros(2), 'facet', 'fixed')]}}, {})


def _prop(model):
    prop = model.prop.copy()
    prop['emod'][:] = 5e4; prop['rho'][:] = 1.0; prop['eta'][:] = 3.0; prop['nu'][:] = 0.45
    prop['ymid'][:] = 1.0
    return prop


def test_glottal_width_measures_host():
    model = _model()
    prop = _prop(model)
    rng = np.random.default_rng(0)
    state = model.state0.copy(); state[:] = 0
    state['u'][:] = rng.uniform(-1e-2, 1e-2, state['u'].size)
    control = model.control.copy(); control['psub'][:] = 8e3; control['psup'][:] = 0
    gw = ppsolid.MeanGlottalWidth(model)(state, control, prop)
    # definition: min over the fluid area vector, mapped entries are 2 (ymid - y)
    ndim = 2
    y = (model.solid.XREF + np.asarray(state['u']))[1::ndim]
    area = np.array(model.fluid.control['area'], copy=True)
    assert gw == area.min()
    gw_solid = ppsolid.MinGlottalWidthFromSolid(model)(state, control, prop)
    assert gw_solid == pytest.approx(np.min(2 * (1.0 - y)), rel=0, abs=0)
    assert gw_solid <= gw + 1e-15      # the surface is a subset of all vertices
    mid = ppsolid.MidpointGlottalWidth(model)(state, control, prop)
    assert mid == gw                   # one fluid channel in 2D


def test_time_series_host_loop(tmp_path):
    model = _model()
    prop = _prop(model)
    rng = np.random.default_rng(1)
    path = str(tmp_path / 'out.h5')
    control = model.control.copy(); control['psub'][:] = 8e3; control['psup'][:] = 0
    states = []
    with sf.StateFile(model, path, mode='w') as f:
        f.append_prop(prop)
        f.append_control(control)
        for n in range(4):
            s = model.state0.copy(); s[:] = 0
            s['u'][:] = rng.uniform(-1e-2, 1e-2, s['u'].size)
            f.append_state(s); f.append_time(1e-4 * n)
            states.append(s)
    with sf.StateFile(model, path, mode='r') as f:
        func = ppsolid.MinGlottalWidthFromSolid(model)        # no batched path: host loop
        series = ppbase.TimeSeries(func)(f)
        assert series.shape == (4,)
        for n, s in enumerate(states):
            assert series[n] == func(s, control, prop)
        sub = ppbase.TimeSeries(func)(f, ns=[3, 1])
        assert np.array_equal(sub, series[[3, 1]])
        stats = ppbase.TimeSeriesStats(func)
        assert stats(f) == pytest.approx(series.mean())
        assert stats.max(f) == series.max() and stats.min(f) == series.min()


@pytest.mark.gpu
def test_glottal_width_series_device_matches_host(tmp_path):
    """TimeSeries(MeanGlottalWidth) over a StateFile of a real forward run: the batched device
    kernel must reproduce the reference's per-state host evaluation bit for bit (it is a min
    of the same 2 (ymid - y) expressions)."""
    from femvf_b200 import forward
    model = _model()
    prop = _prop(model)
    state0 = model.state0.copy(); state0[:] = 0
    control = model.control.copy(); control['psub'][:] = 8e3; control['psup'][:] = 0
    path = str(tmp_path / 'run.h5')
    with sf.StateFile(model, path, mode='w') as f:
        forward.integrate(model, f, state0, [control], prop, 1e-4 * np.arange(40))
    with sf.StateFile(model, path, mode='r') as f:
        func = ppsolid.MeanGlottalWidth(model)
        dev = ppbase.TimeSeries(func)(f)
        host = np.array([func(f.get_state(i), f.get_control(i), None) for i in range(f.size)])
        assert dev.shape == host.shape == (40,)
        assert np.array_equal(dev, host)
        assert host.min() < host.max()       # the folds moved
        assert np.array_equal(ppbase.TimeSeries(func)(f, ns=[5, 7]), host[[5, 7]])
