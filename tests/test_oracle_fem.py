"""
CPU tests pinning the oracle's solid arithmetic (no GPU, no reference runtime).

The reference holds no golden vectors for this path, so the oracle is
checked against (1) an independent sympy evaluation of the weak forms written as in the
UFL text, integrated exactly over one element, (2) Taylor-remainder tests, (3) analytic
known answers, and (4) the committed regression fixtures under tests/golden/.
"""

import os

import numpy as np
import pytest
import sympy as sym

from helpers import mesh_tuples, oracle_problem, random_solid_prop, random_state
from femvf_b200.residuals import solid as slr
from oracle import fem, model as om

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


# --- (1) sympy derivation of one P1 triangle from the weak form ---------------------------

def _sympy_triangle_residual(X, rho, eta, emod, nu, dt, u0, v0, a0, p, edge):
    """F_u(w = phi_a e_i) for one triangle, integrals done exactly by sympy.

    Weak form as in equations/form.py: inner(rho*a1, w) + inner(eta*eps(v1), eps(w))
    + inner(sigma(eps(u1)), eps(w)) over the cell, - inner(-p cof(F) N, w) over `edge`,
    with v1, a1 replaced by the Newmark expressions (form.py:1107-1111).
    """
    x, y = sym.symbols('x y')
    U = sym.Matrix(3, 2, sym.symbols('U0:6'))
    # barycentric coordinates from the vertex coordinates
    A = sym.Matrix([[1, X[a][0], X[a][1]] for a in range(3)])
    coef = A.inv()  # phi_a = coef[0,a] + coef[1,a] x + coef[2,a] y
    phi = [coef[0, a] + coef[1, a] * x + coef[2, a] * y for a in range(3)]

    def field(nodal):
        return sym.Matrix([sum(nodal[a, i] * phi[a] for a in range(3)) for i in range(2)])

    def grad(f):
        return sym.Matrix(2, 2, lambda i, j: sym.diff(f[i], (x, y)[j]))

    def eps(f):
        g = grad(f)
        return (g + g.T) / 2

    gamma, beta = sym.Rational(1, 2), sym.Rational(1, 4)
    u1 = field(U)
    U0 = sym.Matrix(3, 2, list(u0)); V0 = sym.Matrix(3, 2, list(v0)); A0 = sym.Matrix(3, 2, list(a0))
    u0f, v0f, a0f = field(U0), field(V0), field(A0)
    v1 = gamma / beta / dt * (u1 - u0f) - (gamma / beta - 1) * v0f - dt * (gamma / 2 / beta - 1) * a0f
    a1 = 1 / beta / dt**2 * (u1 - u0f - dt * v0f) - (1 / (2 * beta) - 1) * a0f
    lam = emod * nu / (1 + nu) / (1 - 2 * nu)
    mu = emod / 2 / (1 + nu)
    e_u = eps(u1)
    sigma = 2 * mu * e_u + lam * e_u.trace() * sym.eye(2)
    s_visc = eta * eps(v1)

    # integrate over the triangle through the affine map from the reference triangle
    r, s = sym.symbols('r s')
    xm = X[0][0] + (X[1][0] - X[0][0]) * r + (X[2][0] - X[0][0]) * s
    ym = X[0][1] + (X[1][1] - X[0][1]) * r + (X[2][1] - X[0][1]) * s
    detJ = (X[1][0] - X[0][0]) * (X[2][1] - X[0][1]) - (X[1][1] - X[0][1]) * (X[2][0] - X[0][0])

    def int_cell(expr):
        e = sym.expand(expr.subs({x: xm, y: ym}))
        return sym.integrate(sym.integrate(e, (s, 0, 1 - r)), (r, 0, 1)) * abs(detJ)

    i0, i1 = edge
    o = 3 - i0 - i1
    t = sym.symbols('t')
    ex, ey = X[i1][0] - X[i0][0], X[i1][1] - X[i0][1]
    L = sym.sqrt(ex**2 + ey**2)
    n = sym.Matrix([ey, -ex]) / L
    if (X[i0][0] - X[o][0]) * n[0] + (X[i0][1] - X[o][1]) * n[1] < 0:
        n = -n
    Fdef = grad(u1) + sym.eye(2)
    cof = Fdef.adjugate().T  # = det(F) F^{-T}, kept polynomial
    pfield = sum(p[a] * phi[a] for a in range(3))
    traction = -pfield * cof * n  # reference_traction, form.py:752

    def int_edge(expr):
        e = sym.expand(expr.subs({x: X[i0][0] + ex * t, y: X[i0][1] + ey * t}))
        return sym.integrate(e, (t, 0, 1)) * L

    F = []
    for a in range(3):
        for i in range(2):
            w = sym.Matrix([phi[a] if i == 0 else 0, phi[a] if i == 1 else 0])
            ew = eps(w)
            cell = int_cell((rho * a1.T * w)[0] + sum(s_visc[k, l] * ew[k, l] for k in range(2) for l in range(2))
                            + sum(sigma[k, l] * ew[k, l] for k in range(2) for l in range(2)))
            facet = int_edge((traction.T * w)[0])
            F.append(sym.expand(cell - facet))  # solid.py:186: "- SurfacePressureForm"
    return sym.Matrix(F), list(U)


def test_p1_triangle_against_sympy_weak_form():
    rng = np.random.default_rng(3)
    X = [[0.1, 0.05], [0.9, 0.2], [0.3, 0.8]]
    coords = np.array(X)
    cells = np.array([[0, 1, 2]])
    edge = (1, 2)
    prob = fem.SolidProblem(coords, cells, np.array([[1, 2]]), np.array([0]), np.zeros(0, int))
    rho, eta, emod, nu, dt = 1.3, 2.5, 4.0e4, 0.45, 1e-3
    u1 = rng.uniform(-1e-2, 1e-2, 6); u0 = rng.uniform(-1e-2, 1e-2, 6)
    v0 = rng.uniform(-1, 1, 6); a0 = rng.uniform(-10, 10, 6)
    p = rng.uniform(1e3, 5e3, 3)
    prop = dict(rho=np.array([rho]), eta=np.array([eta]), emod=np.array([emod]), nu=nu)

    Rat = lambda v: sym.Rational(repr(float(v)))  # exact rational arithmetic in sympy
    Fsym, Usym = _sympy_triangle_residual(
        [[Rat(c) for c in row] for row in X], Rat(rho), Rat(eta), Rat(emod), Rat(nu), Rat(dt),
        [Rat(v) for v in u0], [Rat(v) for v in v0], [Rat(v) for v in a0], [Rat(v) for v in p],
        edge)
    subs = {s: Rat(v) for s, v in zip(Usym, u1)}
    F_ref = np.array([float(f.subs(subs)) for f in Fsym])
    J_ref = np.array([[float(sym.diff(f, s).subs(subs)) for s in Usym] for f in Fsym])

    F = fem.assemble_res_u(prob, u1, u0, v0, a0, dt, prop, p)
    J = fem.assemble_jac_uu(prob, u1, dt, prop, p).toarray()
    assert np.max(np.abs(F - F_ref)) <= 1e-12 * np.max(np.abs(F_ref))
    assert np.max(np.abs(J - J_ref)) <= 1e-12 * np.max(np.abs(J_ref))


# --- (2) Taylor remainder: Jacobian consistent with the residual ------------------------------

@pytest.mark.parametrize('mesh_name', ['square5', 'cube332', 'm5'])
def test_taylor_remainder(mesh_name):
    rng = np.random.default_rng(0)
    res = slr.KelvinVoigtWEpithelium(*mesh_tuples()[mesh_name]())
    prob = oracle_problem(res)
    N = prob.N
    prop = random_solid_prop(prob, rng, membrane=True)
    so = om.SolidOracle(prob, contact=True, membrane=True)
    u1 = rng.uniform(-1e-2, 1e-2, N)
    zero = np.zeros(N)
    p1 = rng.uniform(0, 8e3, prob.nn)
    dt = 1e-2
    J = fem.assemble_jac_uu(prob, u1, dt, prop, p1, contact=so._contact_args(prop),
                            membrane=so._membrane_args(prop), apply_bc=False)
    du = rng.standard_normal(N)

    def f(u):
        return fem.assemble_res_u(prob, u, zero, zero, zero, dt, prop, p1,
                                  tcontact=so.tcontact(u, prop),
                                  membrane=so._membrane_args(prop), apply_bc=False)
    errs = []
    for eps in (1e-3, 1e-4, 1e-5):
        fd = (f(u1 + eps * du) - f(u1 - eps * du)) / (2 * eps)
        errs.append(np.linalg.norm(fd - J @ du) / np.linalg.norm(J @ du))
    rates = np.log10(np.array(errs[:-1]) / np.array(errs[1:]))
    assert np.all(rates > 1.9), (errs, rates)  # second order for a central difference
    assert errs[-1] < 1e-8


# --- (3) analytic known answers -----------------------------------------------------------------

@pytest.mark.parametrize('mesh_name', ['square5', 'cube332'])
def test_rigid_translation_and_mass(mesh_name):
    res = slr.KelvinVoigt(*mesh_tuples()[mesh_name]())
    prob = oracle_problem(res)
    d, nn, ne, N = prob.d, prob.nn, prob.ne, prob.N
    prop = dict(rho=np.full(ne, 1.7), eta=np.full(ne, 3.0), emod=np.full(ne, 5e4), nu=0.45)
    zero = np.zeros(N)
    # rigid translation at constant velocity: no elastic / viscous force, no inertia
    c = np.arange(1, d + 1, dtype=float)
    u0 = np.tile(0.1 * c, nn); v0 = np.tile(c, nn); dt = 1e-3
    u1 = u0 + dt * v0
    F = fem.assemble_res_u(prob, u1, u0, v0, zero, dt, prop, np.zeros(nn), apply_bc=False)
    assert np.max(np.abs(F)) < 1e-9
    # uniform acceleration g: residual sums to rho * |Omega| * g per component
    g = np.tile(c, nn)
    u1 = 0.25 * dt**2 * g  # a_nmk = 4/dt^2 * u1 with zero initial state
    F = fem.assemble_res_u(prob, u1, zero, zero, zero, dt,
                           dict(prop, eta=np.zeros(ne), emod=np.zeros(ne)), np.zeros(nn),
                           apply_bc=False)
    vol = prob.vol.sum()
    assert np.allclose(F.reshape(nn, d).sum(axis=0), 1.7 * vol * c, rtol=1e-12)


def test_patch_linear_field_equilibrium():
    """A linear displacement field has constant stress: interior nodes are in equilibrium."""
    res = slr.KelvinVoigt(*mesh_tuples()['square5']())
    prob = oracle_problem(res)
    nn, ne, N = prob.nn, prob.ne, prob.N
    Hm = np.array([[1e-3, 2e-3], [-5e-4, 3e-3]])
    u = (prob.coords @ Hm.T).reshape(-1)
    prop = dict(rho=np.zeros(ne), eta=np.zeros(ne), emod=np.full(ne, 5e4), nu=0.45)
    F = fem.assemble_res_u(prob, u, u, np.zeros(N), np.zeros(N), 1.0, prop, np.zeros(nn),
                           apply_bc=False).reshape(nn, 2)
    x = prob.coords
    interior = (x[:, 0] > 1e-9) & (x[:, 0] < 1 - 1e-9) & (x[:, 1] > 1e-9) & (x[:, 1] < 1 - 1e-9)
    assert np.max(np.abs(F[interior])) < 1e-9
    # total force on the boundary balances: sum of nodal forces vanishes
    assert np.max(np.abs(F.sum(axis=0))) < 1e-9


def test_uniform_pressure_total_force_2d():
    """Sum of follower-pressure nodal forces = p * sum(L N) over the loaded boundary (u = 0)."""
    res = slr.KelvinVoigt(*mesh_tuples()['m5']())
    prob = oracle_problem(res)
    nn, ne, N = prob.nn, prob.ne, prob.N
    prop = dict(rho=np.zeros(ne), eta=np.zeros(ne), emod=np.zeros(ne), nu=0.45)
    p0 = 1234.5
    F = fem.assemble_res_u(prob, np.zeros(N), np.zeros(N), np.zeros(N), np.zeros(N), 1.0, prop,
                           np.full(nn, p0), apply_bc=False).reshape(nn, 2)
    expect = p0 * (prob.pf_meas[:, None] * prob.pf_normal).sum(axis=0)
    assert np.allclose(F.sum(axis=0), expect, rtol=1e-12)
    # the loaded boundary closes with the fixed bottom edge (length 0.7895, normal -e_y)
    assert np.allclose(expect, [0.0, p0 * 0.7895], atol=1e-9 * p0)


def test_dirichlet_rows_and_pattern():
    res = slr.KelvinVoigt(*mesh_tuples()['square5']())
    prob = oracle_problem(res)
    prop = dict(rho=np.ones(prob.ne), eta=np.ones(prob.ne), emod=np.full(prob.ne, 5e4), nu=0.45)
    J = fem.assemble_jac_uu(prob, np.zeros(prob.N), 1e-4, prop, np.zeros(prob.nn))
    dense = J.toarray()
    for r in prob.fixed_dofs:
        row = dense[r].copy()
        assert row[r] == 1.0
        row[r] = 0.0
        assert not row.any()
    # columns are NOT eliminated (App. A.4)
    free = np.setdiff1d(np.arange(prob.N), prob.fixed_dofs)
    assert np.abs(dense[np.ix_(free, prob.fixed_dofs)]).max() > 0
    # canonical pattern: sorted columns, full d x d blocks, symmetric structure
    assert np.array_equal(J.indptr, prob.rowptr) and np.array_equal(J.indices, prob.colidx)
    S = (abs(J) > -1).astype(int)  # structural pattern incl. explicit zeros
    assert J.nnz == len(prob.colidx)
    for r in range(prob.N):
        cols = J.indices[J.indptr[r]:J.indptr[r + 1]]
        assert np.all(np.diff(cols) > 0)
        assert len(cols) % prob.d == 0


def test_static_newton_with_contact_converges():
    """config 2 (gentle variant): static equilibrium against the cubic contact penalty."""
    res = slr.KelvinVoigt(*mesh_tuples()['m5']())
    prob = oracle_problem(res)
    ymax = prob.coords[:, 1].max()
    prop = dict(rho=np.ones(prob.ne), eta=np.full(prob.ne, 5.0), emod=np.full(prob.ne, 1e5),
                nu=0.45, ncontact=np.array([0.0, 1.0]), ycontact=ymax - 0.01, kcontact=1e13)
    so = om.SolidOracle(prob, contact=True)
    u, info = om.static_solid_configuration(so, prop, np.zeros(prob.nn))
    assert info['abs_err'] <= 1e-8 or info['rel_err'] <= 1e-10
    gap = fem.contact_gap(prob.coords, u, prop['ncontact'], prop['ycontact'])
    # the surface has been pushed (almost) out of the contact plane
    assert gap.max() < 0.01 and gap.max() > 0
    assert u.reshape(-1, 2)[:, 1].min() < 0


# --- (4) regression fixtures -----------------------------------------------------------------

@pytest.mark.parametrize('name', ['square5_kv', 'cube332_kv', 'm5_epi_contact'])
def test_golden_assembly_fixture(name):
    from golden.make_golden import assembly_case
    z = np.load(os.path.join(GOLDEN, f'assembly_{name}.npz'))
    case = assembly_case(name)
    assert np.array_equal(case['rowptr'], z['rowptr'])
    assert np.array_equal(case['colidx'], z['colidx'])
    assert np.allclose(case['F'], z['F'], rtol=1e-13, atol=1e-13 * np.abs(z['F']).max())
    assert np.allclose(case['J'], z['J'], rtol=1e-13, atol=1e-13 * np.abs(z['J']).max())
