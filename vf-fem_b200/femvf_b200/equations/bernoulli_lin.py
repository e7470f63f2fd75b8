"""
Linearisation of the 1D Bernoulli fluid, d(q, p) / d(area), for the coupled-model Jacobians
(``ExplicitFSIModel.assem_dres_dstate1``).

The reference obtains these by ``jax.jvp`` of ``_BernoulliAreaRatioSep.bernoulli_qp`` /
``_BernoulliFixedSep.bernoulli_qp`` (``/root/reference/src/femvf/residuals/fluid.py:94-107,
252-284``, ``models/transient.py:597-600``).  The closed forms below are the derivatives of those
expressions with the separation / minimum indices frozen (as automatic differentiation of
``argmin`` does).  They are tiny (ns x ns, ns ~ 50) setup-level matrices evaluated on the host
from the model's own control vector; the forward evaluation q, p stays in ``csrc/fluid.cuh``.

With dp = psub - psup and A_sep the separation area,
    q   = sign(dp) sqrt(2 |dp| / rho) A_sep
    p_i = psup + f_i (|dp| - rho q^2 / (2 a_i^2))          (1/2 rho q^2 / A_sep^2 = |dp|)
so  dq/da_j   = sign(dp) sqrt(2 |dp| / rho) dA_sep/da_j
    dp_i/da_j = f_i (-rho q dq/da_j / a_i^2 + rho q^2 / a_i^3 delta_ij da_i/darea_i).
"""

from __future__ import annotations

import numpy as np

from ..residuals.fluid import FLUID_AREA_RATIO_SEP, FLUID_FIXED_SEP


def dqp_darea(kind: int, s, area, psub, psup, rho, r_sep=1.0, area_lb=0.0, idx_sep=0):
    """Returns (dq_darea (ns,), dp_darea (ns, ns)) for one channel."""
    s = np.asarray(s, dtype=float)
    area = np.asarray(area, dtype=float)
    ns = area.size
    dp = float(psub) - float(psup)
    sign = np.sign(dp)
    if kind == FLUID_AREA_RATIO_SEP:
        a = np.maximum(area, area_lb)
        da = (area >= area_lb).astype(float)     # np.maximum passes the first argument on ties
        imin = int(np.argmax(a == a.min()))
        asep = r_sep * a[imin]
        masked = np.where(s >= s[imin], np.abs(a - asep), np.inf)
        isep = int(np.argmin(masked))
        f = (s < s[isep]).astype(float)
        dasep = np.zeros(ns)
        dasep[imin] = r_sep * da[imin]
    elif kind == FLUID_FIXED_SEP:
        a = area
        da = np.ones(ns)
        asep = a[idx_sep]
        f = np.ones(ns)
        f[idx_sep + 1:] = 0.0
        dasep = np.zeros(ns)
        dasep[idx_sep] = 1.0
    else:
        raise NotImplementedError("linearisation is available for the area-ratio and fixed "
                                  "separation models")
    c = sign * np.sqrt(2.0 * abs(dp) / rho)
    q = c * asep
    dq = c * dasep
    dP = -(rho * q) * np.outer(f / a**2, dq) + np.diag(f * rho * q**2 / a**3 * da)
    return dq, dP
