"""
TEST INFRASTRUCTURE ONLY (imported by tests/, smoke() and the CPU legs of bench.py).

P2 (6-node, straight-sided) triangle restatement of the Newmark Kelvin-Voigt solid residual and
its Jacobian -- the "P2 extension" of BASELINE.json's north_star / configs[2].  The reference
itself is P1 only (`/root/reference/src/femvf/equations/form.py:521-524, 545-550` build every
coefficient on `CG 1`), so there is no reference code to follow beyond the weak forms, which are
the ones of oracle/fem.py (form.py:516-533 inertia, 540-572 elastic, 965-990 Kelvin-Voigt,
733-756 follower pressure, 1067-1113 Newmark substitution; residuals/base.py:47-65 Dirichlet).
**Parity unpinned** against the reference; pinned by tests/test_oracle_p2.py (sympy exact
integration of the same weak forms on one element, Taylor remainder, rigid motion / mass /
patch identities, and P1-vs-P2 agreement on affine displacement fields).

Everything is evaluated by numerical quadrature (6-point degree-4 rule on the triangle, 3-point
Gauss rule on the pressure edges): an implementation independent of the closed-form reference
tensors the CUDA kernel uses (csrc/p2_assembly.cuh).

Local node order (DOLFIN / UFC): vertices 0, 1, 2; node 3 on edge (1, 2), node 4 on edge (0, 2),
node 5 on edge (0, 1).
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import fem

# Dunavant degree-4 rule (6 points), weights sum to 1 (fractions of the area)
_A1, _A2 = 0.445948490915965, 0.091576213509771
_W1, _W2 = 0.223381589678011, 0.109951743655322
TRI_QP = np.array([[1 - 2 * _A1, _A1, _A1], [_A1, 1 - 2 * _A1, _A1], [_A1, _A1, 1 - 2 * _A1],
                   [1 - 2 * _A2, _A2, _A2], [_A2, 1 - 2 * _A2, _A2], [_A2, _A2, 1 - 2 * _A2]])
TRI_QW = np.array([_W1, _W1, _W1, _W2, _W2, _W2])
# 3-point Gauss-Legendre on [0, 1]
_G = np.sqrt(3.0 / 5.0)
EDGE_QP = np.array([0.5 - 0.5 * _G, 0.5, 0.5 + 0.5 * _G])
EDGE_QW = np.array([5.0, 8.0, 5.0]) / 18.0

EDGE_OF_NODE = {3: (1, 2), 4: (0, 2), 5: (0, 1)}


def shape(L):
    """P2 shape functions at barycentric points L (..., 3) -> (..., 6)."""
    L0, L1, L2 = L[..., 0], L[..., 1], L[..., 2]
    return np.stack([L0 * (2 * L0 - 1), L1 * (2 * L1 - 1), L2 * (2 * L2 - 1),
                     4 * L1 * L2, 4 * L0 * L2, 4 * L0 * L1], axis=-1)


def dshape_dL(L):
    """d phi_a / d L_k at barycentric points: (..., 6, 3)."""
    L0, L1, L2 = L[..., 0], L[..., 1], L[..., 2]
    z = np.zeros_like(L0)
    return np.stack([
        np.stack([4 * L0 - 1, z, z], -1), np.stack([z, 4 * L1 - 1, z], -1),
        np.stack([z, z, 4 * L2 - 1], -1), np.stack([z, 4 * L2, 4 * L1], -1),
        np.stack([4 * L2, z, 4 * L0], -1), np.stack([4 * L1, 4 * L0, z], -1)], axis=-2)


def p2_mesh(coords, cells):
    """Mid-edge nodes appended after the vertices: returns (coords6 (nn2, 2), cells6 (ne, 6),
    edges (nedge, 2) with edge k -> node nn + k)."""
    coords = np.asarray(coords, dtype=np.float64)
    cells = np.asarray(cells, dtype=np.int64)
    nn = coords.shape[0]
    loc = np.array([EDGE_OF_NODE[3], EDGE_OF_NODE[4], EDGE_OF_NODE[5]])
    ev = np.sort(cells[:, loc], axis=2)                     # (ne, 3, 2)
    key = ev[..., 0] * nn + ev[..., 1]
    uniq, inv = np.unique(key.ravel(), return_inverse=True)
    edges = np.stack([uniq // nn, uniq % nn], axis=1)
    cells6 = np.concatenate([cells, nn + inv.reshape(-1, 3)], axis=1)
    coords6 = np.concatenate([coords, 0.5 * (coords[edges[:, 0]] + coords[edges[:, 1]])])
    return coords6, cells6, edges


class SolidProblemP2:
    """P2 counterpart of fem.SolidProblem (2D).  pfacets: (nfp, 2) VERTEX ids of the pressure
    edges; fixed_nodes: P2 node ids (vertices and mid-edge nodes) with u = 0."""

    def __init__(self, coords, cells, pfacets, pfacet_cells, fixed_nodes):
        self.p1 = fem.SolidProblem(coords, cells, pfacets, pfacet_cells, [])
        self.coords6, self.cells6, self.edges = p2_mesh(coords, cells)
        self.nn = self.coords6.shape[0]
        self.nv = self.p1.nn
        self.ne = self.cells6.shape[0]
        self.d = 2
        self.N = 2 * self.nn
        self.G, self.vol = self.p1.G, self.p1.vol
        fixed_nodes = np.unique(np.asarray(fixed_nodes, dtype=np.int64))
        self.fixed_dofs = (2 * fixed_nodes[:, None] + np.arange(2)[None, :]).ravel()
        self.rowptr, self.colidx = fem.build_pattern(self.nn, self.cells6, 2)
        # pressure edges: local nodes (va, vb, mid) in the parent cell
        pf_local = self.p1.pf_local                                  # (nfp, 2) local vertex ids
        mid = np.array([[-1, 5, 4], [5, -1, 3], [4, 3, -1]])[pf_local[:, 0], pf_local[:, 1]] \
            if len(pf_local) else np.zeros(0, dtype=np.int64)
        self.pf_local3 = np.concatenate([pf_local, mid[:, None]], axis=1) if len(pf_local) \
            else np.zeros((0, 3), dtype=np.int64)

    def closure_nodes(self, vertex_pairs):
        """P2 nodes (both vertices + the mid-edge node) of the given mesh edges."""
        vp = np.sort(np.asarray(vertex_pairs, dtype=np.int64).reshape(-1, 2), axis=1)
        key = self.edges[:, 0] * self.nv + self.edges[:, 1]
        idx = np.searchsorted(key, vp[:, 0] * self.nv + vp[:, 1])
        return np.unique(np.concatenate([vp.ravel(), self.nv + idx]))


def _qp_gradients(prob):
    """grad phi_a at the 6 quadrature points: (nq, ne, 6, 2)."""
    dN = dshape_dL(TRI_QP)                                          # (nq, 6, 3)
    return np.einsum('qak,eki->qeai', dN, prob.G)


def element_matrices(prob: SolidProblemP2, prop):
    """Per-cell K (elastic), C (Kelvin-Voigt viscous), M (mass): each (ne, 6, 2, 6, 2)."""
    ne = prob.ne
    lam, mu = fem.lame(np.broadcast_to(prop['emod'], (ne,)), prop['nu'])
    eta = np.broadcast_to(prop.get('eta', 0.0), (ne,))
    rho = np.broadcast_to(prop['rho'], (ne,))
    g = _qp_gradients(prob)
    eye = np.eye(2)
    K = np.zeros((ne, 6, 2, 6, 2))
    C = np.zeros_like(K)
    for q in range(len(TRI_QW)):
        gq = g[q]                                                   # (ne, 6, 2)
        w = TRI_QW[q] * prob.vol
        gg = np.einsum('eai,ebj->eaibj', gq, gq)                    # g_a,i g_b,j
        dot = np.einsum('eak,ebk->eab', gq, gq)
        # sigma(u):eps(w): lam (div u)(div w) + mu (grad u : grad w + grad u : grad w^T)
        #   block (a, b)[i, j] = lam g_a,i g_b,j + mu g_a,j g_b,i + mu (g_a . g_b) delta_ij
        base = np.einsum('eajbi->eaibj', gg)                        # g_a,j g_b,i
        dd = np.einsum('eab,ij->eaibj', dot, eye)
        K += w[:, None, None, None, None] * (lam[:, None, None, None, None] * gg
                                             + mu[:, None, None, None, None] * (base + dd))
        # eta eps(v) : eps(w) = eta/2 (grad v : grad w + grad v : grad w^T)
        C += (w * 0.5 * eta)[:, None, None, None, None] * (base + dd)
    N = shape(TRI_QP)                                               # (nq, 6)
    m = np.einsum('q,qa,qb->ab', TRI_QW, N, N)
    M = (rho * prob.vol)[:, None, None, None, None] * np.einsum('ab,ij->aibj', m, eye)[None]
    return K, C, M


def _scatter_matrix(prob, Ke, extra=None):
    """Global matrix on the canonical pattern; explicit zeros are kept (the pattern is appended
    with zero values and duplicates are summed by the COO -> CSR conversion, which does not
    prune)."""
    ne = prob.ne
    dofs = (2 * prob.cells6[:, :, None] + np.arange(2)[None, None, :]).reshape(ne, 12)
    rows = [np.repeat(dofs, 12, axis=1).ravel()]
    cols = [np.tile(dofs, (1, 12)).ravel()]
    vals = [Ke.reshape(ne, -1).ravel()]
    prow = np.repeat(np.arange(prob.N), np.diff(prob.rowptr))
    rows.append(prow); cols.append(np.asarray(prob.colidx, dtype=np.int64))
    vals.append(np.zeros(len(prow)))
    if extra is not None:
        vals.append(extra[0]); rows.append(extra[1]); cols.append(extra[2])
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(prob.N, prob.N)).tocsr()
    A.sort_indices()
    return A


def _pressure_terms(prob, u1, p1, want_jac):
    """Follower pressure + int p (cof(F) N) . w ds on the P2 pressure edges, 3-point Gauss.
    Returns (nodal residual (nn, 2), COO triplets of d/du1 or None)."""
    nn = prob.nn
    F = np.zeros((nn, 2))
    trip = ([], [], [])
    if not len(prob.pf_local3):
        return F, None
    pcell = prob.p1.pfacet_cells
    c6 = prob.cells6[pcell]                                         # (nfp, 6)
    U = u1.reshape(nn, 2)[c6]                                       # (nfp, 6, 2)
    P = p1[c6]                                                      # (nfp, 6)
    G = prob.G[pcell]
    nf = len(pcell)
    la, lb = prob.pf_local3[:, 0], prob.pf_local3[:, 1]
    ar = np.arange(nf)
    for t, w in zip(EDGE_QP, EDGE_QW):
        L = np.zeros((nf, 3))
        L[ar, la] = 1 - t
        L[ar, lb] = t
        N = shape(L)                                                # (nfp, 6)
        g = np.einsum('fak,fki->fai', dshape_dL(L), G)              # (nfp, 6, 2)
        gu = np.einsum('fai,faj->fij', U, g)                        # du_i/dx_j
        Fdef = np.eye(2) + gu
        cofn = np.einsum('fij,fj->fi', fem.cofactor(Fdef), prob.p1.pf_normal)
        pq = (P * N).sum(axis=1)
        wq = w * prob.p1.pf_meas
        R = (wq * pq)[:, None, None] * N[:, :, None] * cofn[:, None, :]   # (nfp, 6, 2)
        np.add.at(F, c6.ravel(), R.reshape(-1, 2))
        if want_jac:
            # cof(F) in 2D: [[F11, -F10], [-F01, F00]]; d cof(F)[i, j] / d gu[k, l]
            n = prob.p1.pf_normal
            # d(cof n)_0 = dF11 n0 - dF10 n1 ; d(cof n)_1 = -dF01 n0 + dF00 n1, dF_kl = dU_bk g_b,l
            D = np.zeros((nf, 2, 6, 2))                              # (f, i, b, k)
            D[:, 0, :, 1] = g[:, :, 1] * n[:, None, 0] - g[:, :, 0] * n[:, None, 1]
            D[:, 1, :, 0] = -g[:, :, 1] * n[:, None, 0] + g[:, :, 0] * n[:, None, 1]
            Kp = (wq * pq)[:, None, None, None, None] * N[:, :, None, None, None] * \
                D[:, None, :, :, :]                                  # (f, a, i, b, k)
            dofs = (2 * c6[:, :, None] + np.arange(2)[None, None, :])
            rows = np.broadcast_to(dofs[:, :, :, None, None], Kp.shape)
            cols = np.broadcast_to(dofs[:, None, None, :, :], Kp.shape)
            trip[0].append(Kp.ravel()); trip[1].append(rows.ravel()); trip[2].append(cols.ravel())
    if want_jac:
        return F, tuple(np.concatenate(t) for t in trip)
    return F, None


def assemble_res_u(prob: SolidProblemP2, u1, u0, v0, a0, dt, prop, p1, apply_bc=True):
    """F_u = M a_nmk + C v_nmk + K u1 + pressure, Dirichlet rows zeroed."""
    v1 = fem.newmark_v(u1, u0, v0, a0, dt)
    a1 = fem.newmark_a(u1, u0, v0, a0, dt)
    K, C, M = element_matrices(prob, prop)
    ne = prob.ne
    c6 = prob.cells6
    U = u1.reshape(-1, 2)[c6]; V = v1.reshape(-1, 2)[c6]; A = a1.reshape(-1, 2)[c6]
    Re = np.einsum('eaibj,ebj->eai', K, U) + np.einsum('eaibj,ebj->eai', C, V) + \
        np.einsum('eaibj,ebj->eai', M, A)
    F = np.zeros((prob.nn, 2))
    np.add.at(F, c6.ravel(), Re.reshape(-1, 2))
    Fp, _ = _pressure_terms(prob, u1, p1, False)
    F = (F + Fp).reshape(-1)
    if apply_bc:
        F[prob.fixed_dofs] = 0.0
    return F


def assemble_jac_uu(prob: SolidProblemP2, u1, dt, prop, p1, apply_bc=True):
    """J_uu = ca M + cv C + K + K_p on the canonical P2 pattern (explicit zeros kept)."""
    K, C, M = element_matrices(prob, prop)
    _, trip = _pressure_terms(prob, u1, p1, True)
    J = _scatter_matrix(prob, fem.newmark_a_du1(dt) * M + fem.newmark_v_du1(dt) * C + K, trip)
    if apply_bc:
        J = fem.apply_dirichlet_matrix(J, prob.fixed_dofs)
    return J
