"""Pins of the P2 oracle (oracle/fem_p2.py): exact sympy integration of the weak forms on one
element, Taylor remainder of residual vs Jacobian (including the follower pressure), rigid
motion / mass / patch identities and the CSR pattern."""

import numpy as np
import pytest

from femvf_b200 import mesh as M
from oracle import fem, fem_p2


def _square(n=4):
    mt = M.fixture_mesh_tuple(M.unit_square_mesh(n, n))
    mesh = mt[0]
    return mesh.coordinates(), mesh.cells(), mt


def _problem(n=4, seed=0):
    from helpers import oracle_problem
    from femvf_b200.residuals import solid as slr
    coords, cells, mt = _square(n)
    res = slr.KelvinVoigt(*mt)
    p1 = oracle_problem(res)
    fixed_vertices = np.unique(p1.fixed_dofs // 2)
    # fixed facets: pairs of fixed vertices that form a boundary edge of the 'fixed' subdomain
    fids = res.facet_ids('fixed') if hasattr(res, 'facet_ids') else None
    prob = fem_p2.SolidProblemP2(coords, cells, p1.pfacets, p1.pfacet_cells, [])
    if fids is None:
        # edges whose two vertices are both fixed and that lie on the mesh boundary
        e = prob.edges
        both = np.isin(e[:, 0], fixed_vertices) & np.isin(e[:, 1], fixed_vertices)
        count = np.zeros(len(e), dtype=int)
        np.add.at(count, prob.cells6[:, 3:].ravel() - prob.nv, 1)
        edges_fixed = e[both & (count == 1)]
    else:
        edges_fixed = res.mesh().facets[fids]
    fixed_nodes = prob.closure_nodes(edges_fixed)
    prob = fem_p2.SolidProblemP2(coords, cells, p1.pfacets, p1.pfacet_cells, fixed_nodes)
    rng = np.random.default_rng(seed)
    prop = dict(rho=rng.uniform(0.9, 1.1, prob.ne), eta=rng.uniform(1, 5, prob.ne),
                emod=rng.uniform(2.5e4, 1e5, prob.ne), nu=0.45)
    return prob, prop, rng


def test_element_matrices_against_sympy():
    import sympy as sy
    x, y = sy.symbols('x y')
    X = np.array([[0.1, 0.2], [1.3, 0.1], [0.4, 1.1]])
    prob = fem_p2.SolidProblemP2(X, np.array([[0, 1, 2]]), np.zeros((0, 2), int), [], [])
    lam_, mu_, eta_, rho_ = 3.0, 1.25, 0.7, 1.3
    nu = lam_ / (2 * (lam_ + mu_))
    emod = mu_ * 2 * (1 + nu)
    K, C, Mm = fem_p2.element_matrices(prob, dict(emod=emod, nu=nu, eta=eta_, rho=rho_))
    # barycentric coordinates as polynomials in (x, y)
    A = sy.Matrix([[1, *map(sy.Rational, map(str, X[0]))], [1, *map(sy.Rational, map(str, X[1]))],
                   [1, *map(sy.Rational, map(str, X[2]))]])
    coef = A.inv()
    L = [coef[0, k] + coef[1, k] * x + coef[2, k] * y for k in range(3)]
    phi = [L[0] * (2 * L[0] - 1), L[1] * (2 * L[1] - 1), L[2] * (2 * L[2] - 1),
           4 * L[1] * L[2], 4 * L[0] * L[2], 4 * L[0] * L[1]]

    def integrate(f):
        # map to the reference triangle: x = X0 + s (X1 - X0) + t (X2 - X0)
        s, t = sy.symbols('s t')
        X0, X1, X2 = [sy.Matrix([sy.Rational(str(v)) for v in X[k]]) for k in range(3)]
        P = X0 + s * (X1 - X0) + t * (X2 - X0)
        det = (X1 - X0)[0] * (X2 - X0)[1] - (X1 - X0)[1] * (X2 - X0)[0]
        g = f.subs({x: P[0], y: P[1]}, simultaneous=True)
        return sy.integrate(sy.integrate(g, (t, 0, 1 - s)), (s, 0, 1)) * det

    def vec(a, i):
        v = [0, 0]; v[i] = phi[a]
        return sy.Matrix(v)

    def grad(v):
        return sy.Matrix([[sy.diff(v[0], x), sy.diff(v[0], y)], [sy.diff(v[1], x), sy.diff(v[1], y)]])
    rng = np.random.default_rng(0)
    pairs = [(0, 0, 0, 0), (0, 1, 3, 0), (4, 0, 5, 1), (2, 1, 2, 1), (3, 0, 3, 0), (5, 1, 1, 0),
             (1, 0, 4, 1), (3, 1, 4, 0)]
    for (a, i, b, j) in pairs:
        u, w = vec(b, j), vec(a, i)
        eu, ew = (grad(u) + grad(u).T) / 2, (grad(w) + grad(w).T) / 2
        sig = 2 * sy.Rational(str(mu_)) * eu + sy.Rational(str(lam_)) * eu.trace() * sy.eye(2)
        k_ex = float(integrate(sum(sig[p, q] * ew[p, q] for p in range(2) for q in range(2))))
        c_ex = float(integrate(sy.Rational(str(eta_)) *
                               sum(eu[p, q] * ew[p, q] for p in range(2) for q in range(2))))
        m_ex = float(integrate(sy.Rational(str(rho_)) * (u.T * w)[0]))
        assert abs(K[0, a, i, b, j] - k_ex) <= 1e-12 * max(abs(k_ex), 1.0)
        assert abs(C[0, a, i, b, j] - c_ex) <= 1e-12 * max(abs(c_ex), 1.0)
        assert abs(Mm[0, a, i, b, j] - m_ex) <= 1e-12 * max(abs(m_ex), 1.0)


def test_taylor_remainder_p2():
    prob, prop, rng = _problem()
    N = prob.N
    u1, u0 = rng.uniform(-1e-2, 1e-2, N), rng.uniform(-1e-2, 1e-2, N)
    v0, a0 = rng.uniform(-1, 1, N), rng.uniform(-1e3, 1e3, N)
    p1 = rng.uniform(0, 8e3, prob.nn)
    dt = 1e-4
    J = fem_p2.assemble_jac_uu(prob, u1, dt, prop, p1)
    R0 = fem_p2.assemble_res_u(prob, u1, u0, v0, a0, dt, prop, p1)
    du = rng.standard_normal(N)
    du[prob.fixed_dofs] = 0.0
    errs = []
    for h in (1e-2, 5e-3, 2.5e-3):
        R = fem_p2.assemble_res_u(prob, u1 + h * du, u0, v0, a0, dt, prop, p1)
        errs.append(np.linalg.norm(R - R0 - h * (J @ du)))
    # in 2D cof(F) is linear in grad u, so F_u is AFFINE in u1: the
    # remainder is round-off for every h, which pins J (pressure block included) against R
    for h, e in zip((1e-2, 5e-3, 2.5e-3), errs):
        assert e <= 1e-11 * np.linalg.norm(J @ du), (h, e)
    # and the pressure block is really there
    J0 = fem_p2.assemble_jac_uu(prob, u1, dt, prop, np.zeros(prob.nn))
    assert abs(J - J0).max() > 1.0


def test_rigid_motion_mass_and_patch_p2():
    prob, prop, rng = _problem()
    N, nn = prob.N, prob.nn
    K, C, Mm = fem_p2.element_matrices(prob, prop)
    # total mass
    area_mass = float((np.broadcast_to(prop['rho'], prob.ne) * prob.vol).sum())
    assert abs(Mm[:, :, 0, :, 0].sum() - area_mass) <= 1e-12 * area_mass
    # rigid translation: no elastic / viscous force
    c = np.array([0.3, -0.2])
    U = np.tile(c, 6)
    assert np.max(np.abs(np.einsum('eaibj,bj->eai', K, U.reshape(6, 2)))) <= 1e-9
    # affine displacement with uniform material: interior nodes in equilibrium (patch test)
    prop_u = dict(prop, emod=5e4, eta=0.0)
    A = np.array([[1e-3, 2e-3], [-1e-3, 5e-4]])
    u = (prob.coords6 @ A.T).ravel()
    zero = np.zeros(N)
    # static-like evaluation: u0 = u1, v0 = a0 = 0 -> v_nmk = a_nmk = 0
    R = fem_p2.assemble_res_u(prob, u, u, zero, zero, 1.0, prop_u, np.zeros(nn), apply_bc=False)
    boundary = np.zeros(nn, dtype=bool)
    x = prob.coords6
    boundary[(np.abs(x[:, 0]) < 1e-12) | (np.abs(x[:, 0] - 1) < 1e-12)
             | (np.abs(x[:, 1]) < 1e-12) | (np.abs(x[:, 1] - 1) < 1e-12)] = True
    Rn = R.reshape(nn, 2)
    assert np.max(np.abs(Rn[~boundary])) <= 1e-9 * np.max(np.abs(Rn))


def test_p2_pattern_and_dirichlet_rows():
    prob, prop, rng = _problem()
    u1 = rng.uniform(-1e-3, 1e-3, prob.N)
    J = fem_p2.assemble_jac_uu(prob, u1, 1e-4, prop, rng.uniform(0, 8e3, prob.nn))
    assert np.array_equal(J.indptr, prob.rowptr) and np.array_equal(J.indices, prob.colidx)
    for r in prob.fixed_dofs[:10]:
        row = J.getrow(r).toarray().ravel()
        assert row[r] == 1.0 and np.count_nonzero(row) == 1
    # vertex rows couple to ~19 nodes, mid-edge rows to 9 (interior)
    deg = np.diff(prob.rowptr)[::2] // 2
    assert deg[prob.nv:].max() == 9 and deg[:prob.nv].max() >= 13
