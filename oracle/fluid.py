"""
ORACLE (test infrastructure, NOT product code) -- 1D quasi-steady Bernoulli fluid.

PARITY UNPINNED (see ``oracle/fem.py``): the reference evaluates these closures
through ``jax.jit`` (``/root/reference/src/femvf/models/transient.py:597,662``);
JAX is not installable here.  The restatement follows the closures line by line
in NumPy:
  residuals/fluid.py:17-34    bernoulliq_from_psub_psep, bernoullip_from_q_psep
  residuals/fluid.py:229-311  _BernoulliAreaRatioSep
  residuals/fluid.py:64-128   _BernoulliFixedSep
  residuals/fluid.py:137-220  _BernoulliSmoothMinSep
  equations/smoothapproximation.py:10-30
Shapes: ``s`` and ``area`` are ``(..., ns)``; ``psub, psup, rho`` ... are ``(..., 1)``.
"""

from __future__ import annotations

import numpy as np


def bernoulliq_from_psub_psep(psub, psep, area_sub, area_sep, rho):
    flow_sign = np.sign(psub - psep)
    with np.errstate(divide='ignore'):
        inv_sub = np.where(np.isinf(area_sub), 0.0, np.asarray(area_sub, dtype=float) ** -2.0)
    q = flow_sign * (2 / rho * np.abs(psub - psep) / (area_sep**-2.0 - inv_sub)) ** 0.5
    return q


def bernoullip_from_q_psep(qsub, psep, area_sep, area, rho):
    return psep + 1 / 2 * rho * qsub**2 * (area_sep**-2.0 - area**-2.0)


def bernoulli_area_ratio_sep(s, area, psub, psup, rho, r_sep, area_lb):
    """``_BernoulliAreaRatioSep.bernoulli_qp`` (residuals/fluid.py:252-294)."""
    s = np.asarray(s, dtype=float)
    area = np.maximum(area, area_lb)
    amin = np.min(area, axis=-1, keepdims=True)
    idx_min = np.argmax(area == amin, axis=-1, keepdims=True)
    smin = np.take_along_axis(np.broadcast_to(s, area.shape), idx_min, axis=-1)
    asep = r_sep * amin
    _area = np.where(s >= smin, area, np.nan)
    idx_sep = np.nanargmin(np.abs(_area - asep), axis=-1, keepdims=True)
    ssep = np.take_along_axis(np.broadcast_to(s, area.shape), idx_sep, axis=-1)
    f_sep = np.array(s < ssep, dtype=np.float64)
    q = bernoulliq_from_psub_psep(psub, psup, np.inf, asep, rho)
    p = bernoullip_from_q_psep(q, psup, asep, area, rho)
    p = f_sep * p + (1 - f_sep) * psup
    return q, p


def bernoulli_fixed_sep(s, area, psub, psup, rho, idx_sep):
    """``_BernoulliFixedSep.bernoulli_qp`` (residuals/fluid.py:94-107)."""
    f = np.ones(np.shape(s))
    f[..., idx_sep + 1:] = 0.0
    area_sep = area[..., idx_sep:idx_sep + 1]
    q = bernoulliq_from_psub_psep(psub, psup, np.inf, area_sep, rho)
    p = bernoullip_from_q_psep(q, psup, area_sep, area, rho)
    p = f * p + (1 - f) * psup
    return q, p


def _trapezoid(y, x):
    return np.sum(0.5 * (y[..., 1:] + y[..., :-1]) * np.diff(x, axis=-1), axis=-1)


def _softmax(x):
    m = np.max(x, axis=-1, keepdims=True)
    e = np.exp(x - m)
    return e / np.sum(e, axis=-1, keepdims=True)


def bernoulli_smooth_min_sep(s, area, psub, psup, rho, zeta_min, zeta_sep):
    """``_BernoulliSmoothMinSep.bernoulli_qp`` (residuals/fluid.py:171-191).
    Note ``reshape_args`` sets ``zeta_sep := zeta_min`` (fluid.py:153-154)."""
    s = np.broadcast_to(np.asarray(s, dtype=float), np.shape(area))
    wmin = _softmax(-area / zeta_min)
    amin = (_trapezoid(area * wmin, s) / _trapezoid(wmin, s))[..., None]
    smin = (_trapezoid(s * wmin, s) / _trapezoid(wmin, s))[..., None]
    asep, ssep = amin, smin
    q = bernoulliq_from_psub_psep(psub, psup, np.inf, asep, rho)
    p = bernoullip_from_q_psep(q, psup, asep, area, rho)
    x = -(s - ssep) / zeta_sep
    with np.errstate(over='ignore'):
        f_sep = np.where(x >= 0, 1.0 / (1.0 + np.exp(-x)), np.exp(x) / (1.0 + np.exp(x)))
    p = f_sep * p
    return q, p
