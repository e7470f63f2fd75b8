"""
ORACLE (test infrastructure, NOT product code) -- CPU restatement of the solid
residual / Jacobian assembly on the hot path of ``femvf.forward.integrate``.

PARITY UNPINNED: the reference's arithmetic for this path lives in DOLFIN/FFC
(form compilation + assembly) and PETSc, none of which is vendored under
``/root/reference`` or installable here, and the reference's own tests hold no
golden vectors (``/root/reference/tests/test_forward.py:193`` is ``assert True``).
This restatement is therefore checked against an independent sympy derivation
from the UFL text, Taylor-remainder tests and analytic known answers
(``tests/test_oracle_*.py``), not against outputs of the reference.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this package.

Follows (all paths under /root/reference/src/femvf/):
  equations/form.py:516-533   InertialForm            rho * a1 . w dx
  equations/form.py:540-572   IsotropicElasticForm    sigma(eps(u1)) : eps(w) dx
  equations/form.py:965-990   KelvinVoigtForm         eta * eps(v1) : eps(w) dx
  equations/form.py:918-956   RayleighDampingForm     (rayleigh_m rho v1 . w + rayleigh_k sigma(eps(v1)) : eps(w)) dx
  equations/form.py:733-756   SurfacePressureForm     -p cof(F) N . w ds   (subtracted)
  equations/form.py:759-794   ManualSurfaceContactTractionForm   tc . w ds (subtracted)
  equations/form.py:800-855   IsotropicMembraneForm
  equations/form.py:1067-1113 modify_newmark_time_discretization
  equations/uflcontinuum.py:9-26,73-88,172-186
  equations/newmark.py:8-73
  residuals/solid.py:168-189  KelvinVoigt, :218-240 KelvinVoigtWEpithelium
  residuals/base.py:47-65     Dirichlet BCs on facet subdomain 'fixed'
  models/transient.py:363-406 assem_res / assem_dres_dstate1
  models/transient.py:516-583 NodalContactModel

Conventions: vertex-major interleaved vector DOFs; DG0 properties per cell;
P1 simplices (triangles: plane strain; tetrahedra).  fp64 throughout.
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp

GAMMA = 0.5  # equations/form.py:1083
BETA = 0.25  # equations/form.py:1084


# --- Newmark (equations/newmark.py:8-73) ------------------------------------------

def newmark_v(u, u0, v0, a0, dt, gamma=GAMMA, beta=BETA):
    return gamma / beta / dt * (u - u0) - (gamma / beta - 1.0) * v0 \
        - dt * (gamma / 2.0 / beta - 1.0) * a0


def newmark_a(u, u0, v0, a0, dt, gamma=GAMMA, beta=BETA):
    return 1 / beta / dt**2 * (u - u0 - dt * v0) - (1 / 2 / beta - 1) * a0


def newmark_v_du1(dt, gamma=GAMMA, beta=BETA):
    return gamma / beta / dt


def newmark_a_du1(dt, gamma=GAMMA, beta=BETA):
    return 1.0 / beta / dt**2


# --- problem container -------------------------------------------------------------

class SolidProblem:
    """
    Plain-array description of a tagged P1 simplex mesh for the solid residual.

    Parameters
    ----------
    coords : (nn, d)   vertex coordinates
    cells : (ne, d+1)  positively oriented cells
    pfacets : (nfp, d) vertex ids of the exterior facets in the 'pressure' subdomain
    pfacet_cells : (nfp,) parent cell of each pressure facet
    fixed_dofs : (nfix,) vector DOFs constrained by the Dirichlet BC (u = 0)
    """

    def __init__(self, coords, cells, pfacets, pfacet_cells, fixed_dofs):
        self.coords = np.asarray(coords, dtype=np.float64)
        self.cells = np.asarray(cells, dtype=np.int64)
        self.pfacets = np.asarray(pfacets, dtype=np.int64).reshape(-1, self.coords.shape[1])
        self.pfacet_cells = np.asarray(pfacet_cells, dtype=np.int64)
        self.fixed_dofs = np.unique(np.asarray(fixed_dofs, dtype=np.int64))
        self.d = self.coords.shape[1]
        self.nn = self.coords.shape[0]
        self.ne = self.cells.shape[0]
        self.N = self.d * self.nn
        self.G, self.vol = p1_gradients(self.coords, self.cells)
        self.rowptr, self.colidx = build_pattern(self.nn, self.cells, self.d)
        self._facet_geometry()

    def _facet_geometry(self):
        d = self.d
        x = self.coords
        f = self.pfacets
        pc = self.cells[self.pfacet_cells]  # (nfp, d+1)
        # vertex of the parent cell opposite to the facet
        is_on = (pc[:, :, None] == f[:, None, :]).any(axis=2)
        opp_local = np.argmin(is_on, axis=1)
        opp = pc[np.arange(len(pc)), opp_local]
        if d == 2:
            t = x[f[:, 1]] - x[f[:, 0]]
            n = np.stack([t[:, 1], -t[:, 0]], axis=1)
            meas = np.linalg.norm(t, axis=1)
        else:
            n = np.cross(x[f[:, 1]] - x[f[:, 0]], x[f[:, 2]] - x[f[:, 0]])
            meas = 0.5 * np.linalg.norm(n, axis=1)
        n = n / np.linalg.norm(n, axis=1)[:, None]
        # outward: pointing away from the opposite vertex
        s = np.sign(((x[f[:, 0]] - x[opp]) * n).sum(axis=1))
        self.pf_normal = n * s[:, None]
        self.pf_meas = meas
        # local index (in the parent cell) of each facet vertex
        self.pf_local = np.argmax(pc[:, None, :] == f[:, :, None], axis=2)  # (nfp, d)


def p1_gradients(coords, cells):
    """Constant shape-function gradients G[e, a, :] and cell measures |K|."""
    x = coords[cells]  # (ne, d+1, d)
    d = coords.shape[1]
    e = x[:, 1:, :] - x[:, :1, :]  # rows: edge vectors  (ne, d, d)
    det = np.linalg.det(e)
    vol = det / (2.0 if d == 2 else 6.0)
    inv = np.linalg.inv(e)  # inv[:, :, k] = grad of barycentric coordinate k+1
    G = np.empty((cells.shape[0], d + 1, d))
    G[:, 1:, :] = np.swapaxes(inv, 1, 2)
    G[:, 0, :] = -G[:, 1:, :].sum(axis=1)
    return G, vol


def build_pattern(nn, cells, d):
    """
    Canonical CSR pattern (SURVEY.md section 7 'Hard parts'): full d x d block for
    every vertex pair sharing a cell, explicit zeros kept, columns ascending.
    """
    nen = cells.shape[1]
    ii = np.repeat(cells, nen, axis=1).ravel()
    jj = np.tile(cells, (1, nen)).ravel()
    g = sp.coo_matrix((np.ones(len(ii), dtype=np.int8), (ii, jj)), shape=(nn, nn)).tocsr()
    g.sort_indices()
    brptr, bcol = g.indptr.astype(np.int64), g.indices.astype(np.int64)
    deg = np.diff(brptr)
    rowptr = np.zeros(d * nn + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(np.repeat(deg * d, d))
    colidx = np.empty(rowptr[-1], dtype=np.int64)
    # row (i, a): columns d*bcol + b
    cols_block = (d * bcol[:, None] + np.arange(d)[None, :])  # (nnzb, d)
    # vectorised fill: for node i, its d rows each hold cols_block[brptr[i]:brptr[i+1]].ravel()
    node_of_blk = np.repeat(np.arange(nn), deg)
    k_in_row = np.arange(len(bcol)) - brptr[node_of_blk]
    for a in range(d):
        start = rowptr[d * node_of_blk + a] + k_in_row * d
        for b in range(d):
            colidx[start + b] = cols_block[:, b]
    return rowptr.astype(np.int32), colidx.astype(np.int32)


# --- kinematics (equations/uflcontinuum.py) ----------------------------------------

def lame(emod, nu):
    lam = emod * nu / (1 + nu) / (1 - 2 * nu)  # uflcontinuum.py:22
    mu = emod / 2 / (1 + nu)  # uflcontinuum.py:23
    return lam, mu


def grad_field(G, U):
    """grad u = sum_a U_a (x) G_a  ->  (ne, d, d) with [i, j] = du_i/dx_j."""
    return np.einsum('eai,eaj->eij', U, G)


def cofactor(F):
    """cof(F) = det(F) F^{-T}  (uflcontinuum.py:181-186), closed form for d = 2, 3."""
    d = F.shape[-1]
    C = np.empty_like(F)
    if d == 2:
        C[..., 0, 0] = F[..., 1, 1]
        C[..., 0, 1] = -F[..., 1, 0]
        C[..., 1, 0] = -F[..., 0, 1]
        C[..., 1, 1] = F[..., 0, 0]
    else:
        for i in range(3):
            C[..., i, :] = np.cross(F[..., (i + 1) % 3, :], F[..., (i + 2) % 3, :])
    return C


# --- contact penalty (equations/form.py:1173-1202, models/transient.py:538-583) ------

def contact_gap(coords, u, ncontact, ycontact):
    d = coords.shape[1]
    return (coords + u.reshape(-1, d)) @ ncontact - ycontact


def positive_gap(gap):
    with np.errstate(invalid='ignore'):
        pg = (gap + np.abs(gap)) / 2
    return np.where(gap == -np.inf, 0.0, pg)


def contact_traction(coords, u, ncontact, ycontact, kcontact):
    """tc_i = -k (g+)^3 n   (transient.py:538-552)."""
    g = contact_gap(coords, u, ncontact, ycontact)
    return (-(kcontact * positive_gap(g) ** 3)[:, None] * ncontact[None, :]).reshape(-1)


def contact_dtc_du(coords, u, ncontact, ycontact, kcontact):
    """Per-DOF diagonal d tc / d u used by the reference (transient.py:565-576)."""
    g = contact_gap(coords, u, ncontact, ycontact)
    dp = kcontact * 3 * positive_gap(g) ** 2 * np.sign(g)
    return (-dp[:, None] * ncontact[None, :]).reshape(-1)


# --- element-level residuals ---------------------------------------------------------

def _facet_mass_weights(d):
    """Exact P1 facet mass matrix / measure: (1 + delta_ab) / (d (d+1))."""
    nfv = d
    return (np.ones((nfv, nfv)) + np.eye(nfv)) / (d * (d + 1))


def assemble_res_u(prob: SolidProblem, u1, u0, v0, a0, dt, prop, p1, tcontact=None,
                   membrane=None, apply_bc=True):
    """
    F_u of the Newmark-substituted KelvinVoigt residual (models/transient.py:363-382;
    residuals/solid.py:182-188).

    prop : dict with 'rho', 'eta', 'emod' (per cell, DG0) and 'nu' (scalar)
    p1 : (nn,) nodal pressure;  tcontact : (N,) nodal contact traction or None
    membrane : None or dict(emod_membrane, nu_membrane, th_membrane) per cell
    """
    d, nn, ne = prob.d, prob.nn, prob.ne
    cells, G, vol = prob.cells, prob.G, prob.vol
    v1 = newmark_v(u1, u0, v0, a0, dt)
    a1 = newmark_a(u1, u0, v0, a0, dt)
    U = u1.reshape(nn, d)[cells]
    V = v1.reshape(nn, d)[cells]
    A = a1.reshape(nn, d)[cells]
    lam, mu = lame(np.broadcast_to(prop['emod'], (ne,)), prop['nu'])
    eta = np.broadcast_to(prop.get('eta', 0.0), (ne,))
    rho = np.broadcast_to(prop['rho'], (ne,))

    gu = grad_field(G, U)
    gv = grad_field(G, V)
    eps_u = 0.5 * (gu + np.swapaxes(gu, 1, 2))
    eps_v = 0.5 * (gv + np.swapaxes(gv, 1, 2))
    tr = np.trace(eps_u, axis1=1, axis2=2)
    eye = np.eye(d)
    sigma = 2 * mu[:, None, None] * eps_u + (lam * tr)[:, None, None] * eye
    nen = d + 1
    Mw = (np.ones((nen, nen)) + np.eye(nen)) / ((d + 1) * (d + 2))
    if 'rayleigh_k' in prop:
        # Rayleigh damping (form.py:932-956): rayleigh_k * sigma_iso(eps(v)) + rayleigh_m rho v
        rk, rm = float(np.ravel(prop['rayleigh_k'])[0]), float(np.ravel(prop['rayleigh_m'])[0])
        trv = np.trace(eps_v, axis1=1, axis2=2)
        sigma = sigma + rk * (2 * mu[:, None, None] * eps_v + (lam * trv)[:, None, None] * eye)
        body = A + rm * V
    else:
        sigma = sigma + eta[:, None, None] * eps_v  # form.py:984: eta * eps(v), no factor 2
        body = A
    Re = vol[:, None, None] * np.einsum('eij,eaj->eai', sigma, G)
    Re = Re + (rho * vol)[:, None, None] * np.einsum('ab,ebi->eai', Mw, body)

    F = np.zeros((nn, d))
    np.add.at(F, cells.ravel(), Re.reshape(-1, d))

    # pressure follower load: + int p (cof(F) N) . w ds  (form.py:752, solid.py:186)
    if len(prob.pfacets):
        f = prob.pfacets
        pcell = prob.pfacet_cells
        Fdef = eye + gu[pcell]
        c = np.einsum('fij,fj->fi', cofactor(Fdef), prob.pf_normal)
        Wf = _facet_mass_weights(d)
        pw = prob.pf_meas[:, None] * (p1[f] @ Wf.T)  # int p phi_a ds  (nfp, d)
        Rf = pw[:, :, None] * c[:, None, :]
        if tcontact is not None:
            T = tcontact.reshape(nn, d)[f]  # (nfp, nfv, d)
            Rf = Rf - prob.pf_meas[:, None, None] * np.einsum('ab,fbi->fai', Wf, T)
        np.add.at(F, f.ravel(), Rf.reshape(-1, d))

    F = F.reshape(-1)
    if membrane is not None and len(prob.pfacets):
        F = F + assemble_res_membrane(prob, u1, membrane)
    if apply_bc:
        F[prob.fixed_dofs] = 0.0  # transient.py:379-380
    return F


def _membrane_projector(prob):
    n = prob.pf_normal
    if prob.d == 2:
        n = np.concatenate([n, np.zeros((len(n), 1))], axis=1)
    P = np.eye(3)[None] - n[:, :, None] * n[:, None, :]
    return P


def _embed3(g):
    """Embed a (.., d, d) gradient in 3x3 with zero z row/column (uflcontinuum.py:82-88)."""
    d = g.shape[-1]
    if d == 3:
        return g
    out = np.zeros(g.shape[:-2] + (3, 3))
    out[..., :2, :2] = g
    return out


def _membrane_coeffs(membrane, pcell):
    em = np.asarray(membrane['emod_membrane'])[pcell]
    num = np.asarray(membrane['nu_membrane'])[pcell]
    th = np.asarray(membrane['th_membrane'])[pcell]
    mu = em / 2 / (1 + num)
    lam = em * num / (1 + num) / (1 - 2 * num)
    with np.errstate(divide='ignore', invalid='ignore'):
        lam_pp = np.where(em == 0, 0.0, 2 * mu * lam / (lam + 2 * mu))  # form.py:848-850
    return mu, lam_pp, th


def _membrane_facet_residual(prob, gu_parent, membrane):
    """IsotropicMembraneForm (form.py:812-855) on the pressure facets, P1: constant integrand."""
    d = prob.d
    pcell = prob.pfacet_cells
    mu, lam_pp, th = _membrane_coeffs(membrane, pcell)
    P = _membrane_projector(prob)
    eps = _embed3(0.5 * (gu_parent + np.swapaxes(gu_parent, 1, 2)))
    eps_pp = P @ eps @ P
    trpp = np.trace(eps_pp, axis1=1, axis2=2)
    S = 2 * mu[:, None, None] * eps_pp + (lam_pp * trpp)[:, None, None] * P
    # inner(S, P eps(w) P) = inner(P S P, eps(w)) ; w = phi_a e_i ; grad w = e_i (x) G_a
    PSP = (P @ S @ P)[:, :d, :d]
    Gp = prob.G[pcell]  # (nfp, nen, d) all parent-cell nodes carry a gradient
    Rall = (th * prob.pf_meas)[:, None, None] * np.einsum('fij,faj->fai', PSP, Gp)
    return Rall  # (nfp, nen, d) -- indexed by parent cell node, caller must scatter accordingly


# NOTE: the membrane term tests against ALL parent-cell nodes (grad w is non-zero for the
# node opposite to the facet as well), so it cannot share the facet-node scatter above.
def assemble_res_membrane(prob: SolidProblem, u1, membrane):
    d, nn = prob.d, prob.nn
    pcell = prob.pfacet_cells
    U = u1.reshape(nn, d)[prob.cells[pcell]]
    gu = grad_field(prob.G[pcell], U)
    R = _membrane_facet_residual(prob, gu, membrane)
    F = np.zeros((nn, d))
    np.add.at(F, prob.cells[pcell].ravel(), R.reshape(-1, d))
    return F.reshape(-1)


# --- Jacobian ------------------------------------------------------------------------

def cell_matrices(prob: SolidProblem, dt, prop):
    """Element matrices of d F_u / d u1 from the cell integrals: 4M/dt^2 + 2C/dt + K."""
    d, ne = prob.d, prob.ne
    G, vol = prob.G, prob.vol
    lam, mu = lame(np.broadcast_to(prop['emod'], (ne,)), prop['nu'])
    eta = np.broadcast_to(prop.get('eta', 0.0), (ne,))
    rho = np.broadcast_to(prop['rho'], (ne,))
    cv = newmark_v_du1(dt)
    ca = newmark_a_du1(dt)
    eye = np.eye(d)
    GG = np.einsum('eai,ebi->eab', G, G)  # G_a . G_b
    GaGb = np.einsum('eai,ebj->eabij', G, G)  # G_a (x) G_b
    GbGa = np.swapaxes(GaGb, 3, 4)  # [i,j] = G_b[i] G_a[j]
    nen = d + 1
    Mw = (np.ones((nen, nen)) + np.eye(nen)) / ((d + 1) * (d + 2))
    K = (lam[:, None, None, None, None] * GaGb
         + mu[:, None, None, None, None] * GbGa
         + (mu[:, None, None] * GG)[..., None, None] * eye)
    M = (rho[:, None, None] * Mw[None])[..., None, None] * eye
    if 'rayleigh_k' in prop:
        rk, rm = float(np.ravel(prop['rayleigh_k'])[0]), float(np.ravel(prop['rayleigh_m'])[0])
        C = rk * K + rm * M
    else:
        C = 0.5 * eta[:, None, None, None, None] * (GG[..., None, None] * eye + GbGa)
    Ke = vol[:, None, None, None, None] * (K + cv * C + ca * M)
    return Ke  # (ne, nen, nen, d, d)  block [a, b] = d R_a / d U_b


def assemble_jac_uu(prob: SolidProblem, u1, dt, prop, p1, contact=None, membrane=None,
                    apply_bc=True):
    """
    d F_u / d u1 in the canonical CSR pattern (models/transient.py:384-406; App. A.3).

    contact : None or dict(ncontact, ycontact, kcontact) -> NodalContactModel term
    Returns a scipy CSR matrix whose indptr/indices are exactly ``prob.rowptr/colidx``.
    """
    d, nn, N = prob.d, prob.nn, prob.N
    cells = prob.cells
    nen = d + 1
    Ke = cell_matrices(prob, dt, prop)  # (ne, a, b, i, j)
    rows = (d * cells[:, :, None, None, None] + np.arange(d)[None, None, None, :, None])
    cols = (d * cells[:, None, :, None, None] + np.arange(d)[None, None, None, None, :])
    rows = np.broadcast_to(rows, Ke.shape).ravel()
    cols = np.broadcast_to(cols, Ke.shape).ravel()
    vals = [Ke.ravel()]
    rws = [rows]
    cls = [cols]

    if len(prob.pfacets):
        f = prob.pfacets
        pcell = prob.pfacet_cells
        pc = cells[pcell]  # (nfp, nen)
        Gp = prob.G[pcell]
        Nn = prob.pf_normal
        Wf = _facet_mass_weights(d)
        pw = prob.pf_meas[:, None] * (p1[f] @ Wf.T)  # (nfp, nfv)
        # d c / d U_b, c = cof(F) N
        if d == 2:
            t = Gp[:, :, 1] * Nn[:, None, 0] - Gp[:, :, 0] * Nn[:, None, 1]  # (nfp, nen)
            dc = np.zeros((len(f), nen, 2, 2))
            dc[:, :, 0, 1] = t
            dc[:, :, 1, 0] = -t
        else:
            U = u1.reshape(nn, d)[pc]
            Fdef = np.eye(3) + grad_field(Gp, U)
            q = np.einsum('fij,fbj->fbi', Fdef, np.cross(Nn[:, None, :], Gp))  # F (N x G_b)
            dc = np.zeros((len(f), nen, 3, 3))
            # dc[i, m] = eps_imk q_k = -[q]_x
            dc[:, :, 0, 1] = q[:, :, 2]
            dc[:, :, 0, 2] = -q[:, :, 1]
            dc[:, :, 1, 0] = -q[:, :, 2]
            dc[:, :, 1, 2] = q[:, :, 0]
            dc[:, :, 2, 0] = q[:, :, 1]
            dc[:, :, 2, 1] = -q[:, :, 0]
        Kp = pw[:, :, None, None, None] * dc[:, None, :, :, :]  # (nfp, a_facet, b_cell, i, j)
        r = (d * f[:, :, None, None, None] + np.arange(d)[None, None, None, :, None])
        c = (d * pc[:, None, :, None, None] + np.arange(d)[None, None, None, None, :])
        vals.append(Kp.ravel())
        rws.append(np.broadcast_to(r, Kp.shape).ravel())
        cls.append(np.broadcast_to(c, Kp.shape).ravel())

        if contact is not None:
            # J_c[(a,i),(b,j)] = delta_ij * m_ab * 3k (g+_b)^2 sign(g_b) n_j   (transient.py:554-583)
            dtc = contact_dtc_du(prob.coords, u1, contact['ncontact'], contact['ycontact'],
                                 contact['kcontact']).reshape(nn, d)
            Mf = prob.pf_meas[:, None, None] * Wf[None]  # (nfp, a, b)
            Jc = -Mf[:, :, :, None] * dtc[f][:, None, :, :]  # (nfp, a, b, i) diagonal in (i, j=i)
            r = (d * f[:, :, None, None] + np.arange(d)[None, None, None, :])
            c = (d * f[:, None, :, None] + np.arange(d)[None, None, None, :])
            vals.append(Jc.ravel())
            rws.append(np.broadcast_to(r, Jc.shape).ravel())
            cls.append(np.broadcast_to(c, Jc.shape).ravel())

        if membrane is not None:
            Km = _membrane_facet_matrices(prob, membrane)  # (nfp, a, b, i, j) over parent-cell nodes
            r = (d * pc[:, :, None, None, None] + np.arange(d)[None, None, None, :, None])
            c = (d * pc[:, None, :, None, None] + np.arange(d)[None, None, None, None, :])
            vals.append(Km.ravel())
            rws.append(np.broadcast_to(r, Km.shape).ravel())
            cls.append(np.broadcast_to(c, Km.shape).ravel())

    vals = np.concatenate(vals)
    rws = np.concatenate(rws)
    cls = np.concatenate(cls)
    # explicit zeros for the whole canonical pattern so that the structure is fixed
    prow = np.repeat(np.arange(N), np.diff(prob.rowptr))
    J = sp.coo_matrix(
        (np.concatenate([vals, np.zeros(len(prob.colidx))]),
         (np.concatenate([rws, prow]), np.concatenate([cls, prob.colidx]))),
        shape=(N, N),
    ).tocsr()
    J.sort_indices()
    assert np.array_equal(J.indptr, prob.rowptr) and np.array_equal(J.indices, prob.colidx)
    if apply_bc:
        apply_dirichlet_matrix(J, prob.fixed_dofs)
    return J


def _membrane_facet_matrices(prob, membrane):
    d = prob.d
    pcell = prob.pfacet_cells
    mu, lam_pp, th = _membrane_coeffs(membrane, pcell)
    P = _membrane_projector(prob)[:, :, :]
    Gp = prob.G[pcell]
    nen = d + 1
    nfp = len(pcell)
    K = np.zeros((nfp, nen, nen, d, d))
    for b in range(nen):
        for j in range(d):
            g = np.zeros((nfp, d, d))
            g[:, j, :] = Gp[:, b, :]  # grad of phi_b e_j
            eps = _embed3(0.5 * (g + np.swapaxes(g, 1, 2)))
            eps_pp = P @ eps @ P
            trpp = np.trace(eps_pp, axis1=1, axis2=2)
            S = 2 * mu[:, None, None] * eps_pp + (lam_pp * trpp)[:, None, None] * P
            PSP = (P @ S @ P)[:, :d, :d]
            K[:, :, b, :, j] = (th * prob.pf_meas)[:, None, None] * np.einsum('fik,fak->fai', PSP, Gp)
    return K


def apply_dirichlet_matrix(J, fixed_dofs):
    """``DirichletBC.apply(A)``: zero the rows, unit diagonal; columns untouched (App. A.4)."""
    for r in fixed_dofs:
        lo, hi = J.indptr[r], J.indptr[r + 1]
        J.data[lo:hi] = 0.0
        k = lo + np.searchsorted(J.indices[lo:hi], r)
        J.data[k] = 1.0
    return J
