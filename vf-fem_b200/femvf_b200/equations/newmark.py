"""
Newmark-beta update formulas and their partial derivatives (host-side helpers).

Mirror of ``/root/reference/src/femvf/equations/newmark.py:8-128``; the device path
evaluates the same expressions in ``csrc/elem.cuh`` (``newmark_v``, ``newmark_a``).
"""


def newmark_v(u, u0, v0, a0, dt, gamma=1 / 2, beta=1 / 4):
    """Newmark velocity update (``newmark.py:8-29``)."""
    return (
        gamma / beta / dt * (u - u0)
        - (gamma / beta - 1.0) * v0
        - dt * (gamma / 2.0 / beta - 1.0) * a0
    )


def newmark_v_du1(dt, gamma=1 / 2, beta=1 / 4):
    return gamma / beta / dt


def newmark_v_du0(dt, gamma=1 / 2, beta=1 / 4):
    return -gamma / beta / dt


def newmark_v_dv0(dt, gamma=1 / 2, beta=1 / 4):
    return -(gamma / beta - 1.0)


def newmark_v_da0(dt, gamma=1 / 2, beta=1 / 4):
    return -dt * (gamma / 2.0 / beta - 1.0)


def newmark_v_dt(u, u0, v0, a0, dt, gamma=1 / 2, beta=1 / 4):
    return -gamma / beta / dt**2 * (u - u0) - (gamma / 2.0 / beta - 1.0) * a0


def newmark_a(u, u0, v0, a0, dt, gamma=1 / 2, beta=1 / 4):
    """Newmark acceleration update (``newmark.py:57-73``)."""
    return 1 / beta / dt**2 * (u - u0 - dt * v0) - (1 / 2 / beta - 1) * a0


def newmark_a_du1(dt, gamma=1 / 2, beta=1 / 4):
    return 1.0 / beta / dt**2


def newmark_a_du0(dt, gamma=1 / 2, beta=1 / 4):
    return -1.0 / beta / dt**2


def newmark_a_dv0(dt, gamma=1 / 2, beta=1 / 4):
    return -1.0 / beta / dt


def newmark_a_da0(dt, gamma=1 / 2, beta=1 / 4):
    return -(1 / 2 / beta - 1)


def newmark_a_dt(u, u0, v0, a0, dt, gamma=1 / 2, beta=1 / 4):
    return -2 / beta / dt**3 * (u - u0 - dt * v0) + 1 / beta / dt**2 * (-v0)


def newmark_error_estimate(a1, a0, dt, beta=1 / 4):
    """Truncation error estimate of Zienkiewicz & Xie (``newmark.py:101-128``)."""
    return 0.5 * dt**2 * (2 * beta - 1 / 3) * (a1 - a0)
