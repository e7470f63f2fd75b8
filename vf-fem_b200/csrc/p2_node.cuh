// Per-node arithmetic of the P2 triangle kernels (csrc/p2.cu), host/device: the CUDA kernel calls
// p2_node_row from one thread per node, and the test-only CPU harness (tests/hostcheck) runs the
// same function in a loop against the oracle.  Weak forms and references: see p2.cu.
#pragma once

#include <cstddef>

#include "elem.cuh"

namespace vf {

struct alignas(16) P2Pair { double x, y; };  // nodal (x, y) pair: one 16-byte access
struct alignas(8) P2IPair { int x, y; };

// device tables of the P2 assembler (subset of vf_p2 the per-node code reads)
struct P2View {
  const double* xy;        // (nn, 2)
  const int* cells;        // (ne, 6)
  const int* brptr;
  const int* bcol;
  const int* n2e_ptr;
  const int* n2e;          // pairs: cell * 8 + local node
  const unsigned* n2e_slots;  // per pair: CSR slots of the cell's 6 nodes in this node's row
  const int* n2f_ptr;
  const int* n2f;          // pressure edge * 4 + local position
  const int* n2f_pair;
  const int* pf_cell;
  const int* pf_loc;       // (nfp, 3)
  const double* pf_geo;    // (nfp, 3)
  const unsigned char* fixed;
};

struct P2Args {
  const double *emod, *eta, *rho, *u1, *u0, *v0, *a0, *p1;
  // optional (nn, 6) packed nodal (u1x, u1y, v_nmk x, y, a_nmk x, y) written by a pre-pass
  // (p2_pack_state): two 32-byte sectors per gathered node instead of four, and the Newmark
  // update evaluated once per node instead of once per (node, cell, local node) visit
  const double* uva = nullptr;
  double* F;
  double* J;
  double nu, dt;
  int jac, res;
};

VF_HD void p2_shape_edge(double t, int la, int lb, double (&N)[6],
                                              double (&D)[6][3]) {
  double L[3] = {0.0, 0.0, 0.0};
  L[la] = 1.0 - t;
  L[lb] = t;
  N[0] = L[0] * (2 * L[0] - 1);
  N[1] = L[1] * (2 * L[1] - 1);
  N[2] = L[2] * (2 * L[2] - 1);
  N[3] = 4 * L[1] * L[2];
  N[4] = 4 * L[0] * L[2];
  N[5] = 4 * L[0] * L[1];
  for (int a = 0; a < 6; ++a)
    for (int k = 0; k < 3; ++k) D[a][k] = 0.0;
  D[0][0] = 4 * L[0] - 1;
  D[1][1] = 4 * L[1] - 1;
  D[2][2] = 4 * L[2] - 1;
  D[3][1] = 4 * L[2];
  D[3][2] = 4 * L[1];
  D[4][0] = 4 * L[2];
  D[4][2] = 4 * L[0];
  D[5][0] = 4 * L[1];
  D[5][1] = 4 * L[0];
}

// Packed nodal state of node n for P2Args::uva.
VF_HD void p2_pack_state(const P2Args& A, const NewmarkCoef& nc, int n, double* uva) {
  const double u1x = A.u1[2 * n], u1y = A.u1[2 * n + 1];
  const double u0x = A.u0[2 * n], u0y = A.u0[2 * n + 1];
  const double v0x = A.v0[2 * n], v0y = A.v0[2 * n + 1];
  const double a0x = A.a0[2 * n], a0y = A.a0[2 * n + 1];
  double* q = uva + 6 * (size_t)n;
  q[0] = u1x;
  q[1] = u1y;
  q[2] = newmark_v(nc, u1x, u0x, v0x, a0x);
  q[3] = newmark_v(nc, u1y, u0y, v0y, a0y);
  q[4] = newmark_a(nc, u1x, u0x, v0x, a0x);
  q[5] = newmark_a(nc, u1y, u0y, v0y, a0y);
}

// Block row and residual entries of node i.  CLS 0: i is a vertex node (local index a < 3 in all
// its cells), CLS 1: a mid-edge node (a >= 3).  W = kP2W (324 doubles), M = kP2M (36).  The row
// is written in the layout of the CSR array: [scalar row 0: deg x (c0, c1)][scalar row 1: ...],
// 4 * deg doubles at `row`.  dphi_a/dL_k vanishes unless k = a (vertex node) or k is a vertex of
// the edge (mid-edge node), so only the 1 / 2 / 4 structural non-zeros of W_ab.. are summed, in
// the (k, l) order of the full loop.
template <int CLS>
VF_HD void p2_node_row(const P2View& P, const P2Args& A, const double* W, const double* M, int i,
                       double* row, double& r0_out, double& r1_out) {
  const int b0 = P.brptr[i];
  const int deg = P.brptr[i + 1] - b0;
  double* row1 = row + 2 * deg;  // second scalar row of the block row
  if (A.jac)
    for (int s = 0; s < 4 * deg; ++s) row[s] = 0.0;
  double r0 = 0.0, r1 = 0.0;
  const NewmarkCoef nc = newmark_coef(A.dt);
  const double cv = nc.cv, ca = nc.ca;
  const LameFac lf = lame_fac(A.nu);
  const P2Pair* xy2 = reinterpret_cast<const P2Pair*>(P.xy);
  const P2Pair* u1v = reinterpret_cast<const P2Pair*>(A.u1);
  const P2Pair* u0v = reinterpret_cast<const P2Pair*>(A.u0);
  const P2Pair* v0v = reinterpret_cast<const P2Pair*>(A.v0);
  const P2Pair* a0v = reinterpret_cast<const P2Pair*>(A.a0);

  for (int t = P.n2e_ptr[i]; t < P.n2e_ptr[i + 1]; ++t) {
    const int ref = P.n2e[t];
    const int e = ref >> 3, a = ref & 7;
    const unsigned slots = P.n2e_slots[t];
    int nd[6];
    {
      const P2IPair* c2 = reinterpret_cast<const P2IPair*>(P.cells + 6 * (size_t)e);
      const P2IPair q0 = c2[0], q1 = c2[1], q2 = c2[2];
      nd[0] = q0.x; nd[1] = q0.y; nd[2] = q1.x; nd[3] = q1.y; nd[4] = q2.x; nd[5] = q2.y;
    }
    const P2Pair x0 = xy2[nd[0]], x1 = xy2[nd[1]], x2 = xy2[nd[2]];
    const double e1x = x1.x - x0.x, e1y = x1.y - x0.y;
    const double e2x = x2.x - x0.x, e2y = x2.y - x0.y;
    const double det = e1x * e2y - e1y * e2x, idet = 1.0 / det;
    double G[3][2];
    G[1][0] = e2y * idet;
    G[1][1] = -e2x * idet;
    G[2][0] = -e1y * idet;
    G[2][1] = e1x * idet;
    G[0][0] = -G[1][0] - G[2][0];
    G[0][1] = -G[1][1] - G[2][1];
    const double vol = 0.5 * det;
    const double emod = A.emod[e], eta = A.eta[e], rho = A.rho[e];
    const double lam = emod * lf.lam_fac, mu = emod * lf.mu_fac;
    // the L-derivatives of phi_a that do not vanish: k = a, or the vertices of edge node a
    // (3: 1,2   4: 0,2   5: 0,1), ascending like the full (k, l) loop of version 1
    constexpr int NK = CLS == 0 ? 1 : 2;
    int ks[2];
    if (CLS == 0) {
      ks[0] = ks[1] = a;
    } else {
      ks[0] = (a == 3) ? 1 : 0;
      ks[1] = (a == 5) ? 1 : 2;
    }
    double Gk[NK][2];
#pragma unroll
    for (int j = 0; j < NK; ++j) {
      Gk[j][0] = ks[j] == 0 ? G[0][0] : (ks[j] == 1 ? G[1][0] : G[2][0]);
      Gk[j][1] = ks[j] == 0 ? G[0][1] : (ks[j] == 1 ? G[1][1] : G[2][1]);
    }
    const double* Wa = W + a * 54;
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      constexpr int kEdgeLo[3] = {1, 0, 0}, kEdgeHi[3] = {2, 2, 1};
      const int nl = b < 3 ? 1 : 2;
      double T00 = 0.0, T01 = 0.0, T10 = 0.0, T11 = 0.0;
#pragma unroll
      for (int j = 0; j < NK; ++j)
#pragma unroll
        for (int li = 0; li < nl; ++li) {
          const int l = b < 3 ? b : (li == 0 ? kEdgeLo[b < 3 ? 0 : b - 3] : kEdgeHi[b < 3 ? 0 : b - 3]);
          const double w = Wa[(b * 3 + ks[j]) * 3 + l];
          T00 += w * Gk[j][0] * G[l][0];
          T01 += w * Gk[j][0] * G[l][1];
          T10 += w * Gk[j][1] * G[l][0];
          T11 += w * Gk[j][1] * G[l][1];
        }
      const double tr = T00 + T11;
      const double mab = rho * vol * M[a * 6 + b];
      const double S00 = T00 + tr, S01 = T10, S10 = T01, S11 = T11 + tr;
      const double K00 = vol * (lam * T00 + mu * S00), K01 = vol * (lam * T01 + mu * S01);
      const double K10 = vol * (lam * T10 + mu * S10), K11 = vol * (lam * T11 + mu * S11);
      const double ch = 0.5 * eta * vol;
      if (A.res) {
        const int n = nd[b];
        P2Pair u1;
        double vx, vy, ax, ay;
        if (A.uva) {
          const P2Pair* q = reinterpret_cast<const P2Pair*>(A.uva) + 3 * (size_t)n;
          const P2Pair v = q[1], acc = q[2];
          u1 = q[0];
          vx = v.x; vy = v.y; ax = acc.x; ay = acc.y;
        } else {
          const P2Pair u0 = u0v[n], v0 = v0v[n], a0 = a0v[n];
          u1 = u1v[n];
          vx = newmark_v(nc, u1.x, u0.x, v0.x, a0.x); vy = newmark_v(nc, u1.y, u0.y, v0.y, a0.y);
          ax = newmark_a(nc, u1.x, u0.x, v0.x, a0.x); ay = newmark_a(nc, u1.y, u0.y, v0.y, a0.y);
        }
        r0 += K00 * u1.x + K01 * u1.y + ch * (S00 * vx + S01 * vy) + mab * ax;
        r1 += K10 * u1.x + K11 * u1.y + ch * (S10 * vx + S11 * vy) + mab * ay;
      }
      if (A.jac) {
        const int s = (slots >> (5 * b)) & 31;
        const double cc = cv * ch, mm = ca * mab;
        row[2 * s + 0] += K00 + cc * S00 + mm;
        row[2 * s + 1] += K01 + cc * S01;
        row1[2 * s + 0] += K10 + cc * S10;
        row1[2 * s + 1] += K11 + cc * S11 + mm;
      }
    }
  }

  // follower pressure on the P2 edges: + int p (cof(F) N) . w ds, three Gauss points
  for (int t = P.n2f_ptr[i]; t < P.n2f_ptr[i + 1]; ++t) {
    const int ref = P.n2f[t];
    const int f = ref >> 2, pos = ref & 3;
    const int e = P.pf_cell[f];
    const int la = P.pf_loc[3 * f], lb = P.pf_loc[3 * f + 1], lm = P.pf_loc[3 * f + 2];
    const int a = pos == 0 ? la : (pos == 1 ? lb : lm);
    const double nx = P.pf_geo[3 * f], ny = P.pf_geo[3 * f + 1], len = P.pf_geo[3 * f + 2];
    const unsigned slots = P.n2e_slots[P.n2f_pair[t]];
    int nd[6];
    for (int b = 0; b < 6; ++b) nd[b] = P.cells[6 * e + b];
    double x[3][2];
    for (int k = 0; k < 3; ++k) {
      x[k][0] = P.xy[2 * nd[k]];
      x[k][1] = P.xy[2 * nd[k] + 1];
    }
    const double e1x = x[1][0] - x[0][0], e1y = x[1][1] - x[0][1];
    const double e2x = x[2][0] - x[0][0], e2y = x[2][1] - x[0][1];
    const double idet = 1.0 / (e1x * e2y - e1y * e2x);
    double G[3][2];
    G[1][0] = e2y * idet;
    G[1][1] = -e2x * idet;
    G[2][0] = -e1y * idet;
    G[2][1] = e1x * idet;
    G[0][0] = -G[1][0] - G[2][0];
    G[0][1] = -G[1][1] - G[2][1];
    const double gq = 0.7745966692414834;  // sqrt(3/5)
    const double tq[3] = {0.5 - 0.5 * gq, 0.5, 0.5 + 0.5 * gq};
    const double wq[3] = {5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0};
    for (int q = 0; q < 3; ++q) {
      double N[6], D[6][3];
      p2_shape_edge(tq[q], la, lb, N, D);
      double g[6][2];
      double gu00 = 0.0, gu01 = 0.0, gu10 = 0.0, gu11 = 0.0, pq = 0.0;
      for (int b = 0; b < 6; ++b) {
        g[b][0] = D[b][0] * G[0][0] + D[b][1] * G[1][0] + D[b][2] * G[2][0];
        g[b][1] = D[b][0] * G[0][1] + D[b][1] * G[1][1] + D[b][2] * G[2][1];
        const double ux = A.u1[2 * nd[b]], uy = A.u1[2 * nd[b] + 1];
        gu00 += ux * g[b][0];
        gu01 += ux * g[b][1];
        gu10 += uy * g[b][0];
        gu11 += uy * g[b][1];
        pq += A.p1[nd[b]] * N[b];
      }
      const double c0 = (1.0 + gu11) * nx - gu10 * ny;
      const double c1 = -gu01 * nx + (1.0 + gu00) * ny;
      const double w = wq[q] * len * pq * N[a];
      if (A.res) {
        r0 += w * c0;
        r1 += w * c1;
      }
      if (A.jac) {
        for (int b = 0; b < 6; ++b) {
          const int s = (slots >> (5 * b)) & 31;
          const double dd = g[b][1] * nx - g[b][0] * ny;
          row[2 * s + 1] += w * dd;
          row1[2 * s + 0] -= w * dd;
        }
      }
    }
  }

  // Dirichlet rows (residuals/base.py:47-65): zero row, unit diagonal, zero residual
  if (P.fixed[i]) {
    r0 = r1 = 0.0;
    if (A.jac) {
      for (int s = 0; s < deg; ++s) {
        const bool self = P.bcol[b0 + s] == i;
        row[2 * s + 0] = self ? 1.0 : 0.0;
        row[2 * s + 1] = 0.0;
        row1[2 * s + 0] = 0.0;
        row1[2 * s + 1] = self ? 1.0 : 0.0;
      }
    }
  }
  r0_out = r0;
  r1_out = r1;
}

}  // namespace vf
