"""CPU tests of the oracle's Bernoulli fluid against closed-form answers and fixtures."""

import os

import numpy as np
import pytest

from oracle import fluid as ofl

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def fixture():
    """tests/residuals/test_fluid.py:12-45 of the reference."""
    s = np.linspace(0, 1, 11)
    area = np.abs(s - 0.5)
    area[area < 0.1] = 0.1
    return s, area, np.array([100.0]), np.array([0.0]), np.array([1.0])


def test_area_ratio_closed_form():
    s, area, psub, psup, rho = fixture()
    q, p = ofl.bernoulli_area_ratio_sep(s, area, psub, psup, rho, np.array([1.0]), np.array([0.0]))
    amin = 0.1
    # q = sqrt(2 dp / rho) * A_sep ; separation at the first minimum (s = 0.4)
    assert np.isclose(q[0], np.sqrt(2 * 100.0) * amin, rtol=1e-14)
    expect = np.where(s < 0.4 - 1e-12, 0.5 * q[0]**2 * (amin**-2 - area**-2), 0.0)
    assert np.allclose(p, expect, rtol=1e-13, atol=1e-12)
    # Bernoulli: p + q^2/(2 A^2) is constant upstream of separation
    up = s < 0.4 - 1e-12
    assert np.allclose(p[up] + 0.5 * q[0]**2 / area[up]**2, 0.5 * q[0]**2 / amin**2, rtol=1e-13)


def test_area_ratio_separation_downstream_and_lower_bound():
    s, area, psub, psup, rho = fixture()
    # exact ties in the area (linspace round-off would break them): the first index wins
    area = np.array([0.5, 0.4, 0.3, 0.2, 0.1, 0.1, 0.1, 0.2, 0.3, 0.4, 0.5])
    q, p = ofl.bernoulli_area_ratio_sep(s, area, psub, psup, rho, np.array([1.2]), np.array([0.0]))
    # A_sep = 1.2 * 0.1: closest downstream area to 0.12 is 0.1 (s=0.4..0.6, first index wins)
    assert np.isclose(q[0], np.sqrt(200.0) * 0.12)
    assert np.all(p[s >= 0.4 - 1e-12] == 0.0)
    # a lower bound on the area replaces smaller areas
    q2, _ = ofl.bernoulli_area_ratio_sep(s, area, psub, psup, rho, np.array([1.0]), np.array([0.2]))
    assert np.isclose(q2[0], np.sqrt(200.0) * 0.2)
    # reversed pressure drop reverses the flow
    q3, _ = ofl.bernoulli_area_ratio_sep(s, area, psup, psub, rho, np.array([1.0]), np.array([0.0]))
    assert np.isclose(q3[0], -q[0] / 1.2)


def test_fixed_sep_and_batched_shapes():
    s, area, psub, psup, rho = fixture()
    q, p = ofl.bernoulli_fixed_sep(s, area, psub, psup, rho, 5)
    assert np.isclose(q[0], np.sqrt(200.0) * area[5])
    assert np.all(p[6:] == 0.0) and np.isclose(p[5], 0.0, atol=1e-12)
    # (nz, ns) batches evaluate each plane independently
    S = np.stack([s, s]); A = np.stack([area, 2 * area])
    col = lambda v: np.full((2, 1), v)
    qb, pb = ofl.bernoulli_area_ratio_sep(S, A, col(100.0), col(0.0), col(1.0), col(1.0), col(0.0))
    assert qb.shape == (2, 1) and pb.shape == (2, 11)
    assert np.isclose(qb[1, 0], 2 * qb[0, 0])


def test_smooth_min_tends_to_hard_min():
    s, area, psub, psup, rho = fixture()
    q, p = ofl.bernoulli_smooth_min_sep(s, area, psub, psup, rho, np.array([1e-4]), np.array([1e-4]))
    assert np.isclose(q[0], np.sqrt(200.0) * 0.1, rtol=1e-6)


def test_golden_bernoulli_fixture():
    from golden.make_golden import bernoulli_case
    z = np.load(os.path.join(GOLDEN, 'bernoulli.npz'))
    case = bernoulli_case()
    for key in z.files:
        assert np.allclose(np.ravel(case[key]), np.ravel(z[key]), rtol=1e-14, atol=1e-14), key


def test_golden_forward_fixture():
    from golden.make_golden import forward_case
    z = np.load(os.path.join(GOLDEN, 'forward_m5.npz'))
    case = forward_case()
    for key in ('u_last', 'q', 'p_last'):
        scale = np.max(np.abs(z[key]))
        assert np.max(np.abs(case[key] - z[key])) <= 1e-9 * scale, key


@pytest.mark.parametrize('kind', ['area_ratio', 'fixed'])
def test_bernoulli_linearisation_matches_finite_differences(kind):
    """femvf_b200.equations.bernoulli_lin (closed-form d(q, p)/d(area) used by the coupled-model
    Jacobians) against central differences of the oracle's Bernoulli functions."""
    from femvf_b200.equations import bernoulli_lin
    from femvf_b200.residuals.fluid import FLUID_AREA_RATIO_SEP, FLUID_FIXED_SEP
    rng = np.random.default_rng(2)
    s = np.linspace(0, 1, 21)
    area = 0.2 + 0.1 * np.cos(2 * np.pi * s) ** 2 + 0.02 * rng.uniform(size=s.size)
    area[3] = 0.05            # one entry below the lower bound (clipped: zero derivative)
    psub, psup, rho, r_sep, lb, idx = 800.0, 10.0, 1.2e-3, 1.2, 0.08, 12

    def qp(a):
        if kind == 'area_ratio':
            q, p = ofl.bernoulli_area_ratio_sep(s, a, psub, psup, rho, r_sep, lb)
        else:
            q, p = ofl.bernoulli_fixed_sep(s, a, psub, psup, rho, idx)
        return float(np.ravel(q)[0]), np.ravel(p)
    k = FLUID_AREA_RATIO_SEP if kind == 'area_ratio' else FLUID_FIXED_SEP
    dq, dP = bernoulli_lin.dqp_darea(k, s, area, psub, psup, rho, r_sep, lb, idx)
    h = 1e-7
    for j in range(s.size):
        ap, am = area.copy(), area.copy()
        ap[j] += h; am[j] -= h
        (qp_, pp), (qm, pm) = qp(ap), qp(am)
        assert abs((qp_ - qm) / (2 * h) - dq[j]) <= 1e-6 * max(abs(dq).max(), 1e-30)
        assert np.max(np.abs((pp - pm) / (2 * h) - dP[:, j])) <= 1e-5 * np.abs(dP).max()
