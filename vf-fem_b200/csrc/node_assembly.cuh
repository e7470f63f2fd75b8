// Owner-computes ("sorted-segment") assembly of one node's block row of
// J_uu = d F_u / d u1 and of its residual entries F_u.
//
// Stands in for DOLFIN's assemble_cells / assemble_exterior_facets + MatSetValues +
// DirichletBC.apply as driven by the reference at
// /root/reference/src/femvf/models/transient.py:363-406 (assem_res, assem_dres_dstate1)
// and :516-583 (NodalContactModel).  One caller (a thread) owns node i: it visits the
// cells and pressure facets adjacent to i in a fixed order, so no atomics are needed and
// the result is bit-reproducible run to run.
//
// Layout of a block row (identical in shared memory tiles and in the global CSR array):
// the d rows of node i are stored one after another, row a at rowblk + a*ld with
// ld = d*deg(i); entry for neighbour slot k, component b at k*d + b.  With
// rowblk = J + d*d*brptr[i] this is exactly the canonical scalar CSR (columns ascending,
// full d x d blocks, explicit zeros kept).
#pragma once

#include "elem.cuh"

namespace vf {

// indices into the per-member scalar property block
enum ScalarProp {
  SC_NU = 0,
  SC_YCONTACT = 1,
  SC_KCONTACT = 2,
  SC_NCONTACT = 3,  // 3 entries
  SC_YMID = 6,
  SC_RAYLEIGH_M = 7,
  SC_RAYLEIGH_K = 8,
  SC_COUNT = 10
};

struct MeshView {
  int dim, nn, ne, nfp;
  const double* xyz;    // SoA coordinates: xyz[c*nn + node]
  const double* xy;     // 2D only: interleaved (x, y) per node for 16-byte gathers (else null)
  const int* cells;     // SoA connectivity: cells[a*ne + e]
  const int* brptr;     // node graph = block CSR pattern of J
  const int* bcol;
  const int* n2e_ptr;   // node -> adjacent cells, packed e*4 + local index
  const int* n2e;
  const int* n2f_ptr;   // node -> adjacent pressure facets, packed f*4 + local index in parent cell
  const int* n2f;
  const int* pf_cell;   // pressure facet -> parent cell
  const int* pf_opp;    // pressure facet -> local vertex of the parent cell opposite to it
  const unsigned char* bc;  // per-DOF Dirichlet flag
  // tetrahedra only, optional (tet_tables.h; null: generic gathers): one record per cell / node
  // and the precomputed CSR slots of every (node, cell) pair
  const int* cells4 = nullptr;           // (ne, 4), 16-byte aligned
  const double* xyz4 = nullptr;          // (nn, 4): x, y, z, 0, 32-byte aligned
  const unsigned* n2e_slots = nullptr;   // per n2e entry: byte c = slot of local node c
};

struct
#if defined(__CUDACC__)
    __align__(16)
#else
    alignas(16)
#endif
        I4 {
  int x, y, z, w;
};

struct PropView {
  const double* rho;   // DG0, per cell
  const double* eta;
  const double* emod;
  const double* scal;  // ScalarProp block
  const double* emod_m;  // membrane DG0 fields (may be null when membrane == 0)
  const double* nu_m;
  const double* th_m;
  int contact;   // NodalContactModel semantics on/off (App. C, Q3)
  int membrane;  // KelvinVoigtWEpithelium membrane term on/off
  int damping;   // DampingKind: Kelvin-Voigt or Rayleigh
};

VF_HD Damping prop_damping(const PropView& p) {
  Damping d;
  d.kind = p.damping;
  d.rm = p.scal[SC_RAYLEIGH_M];
  d.rk = p.scal[SC_RAYLEIGH_K];
  return d;
}

struct StateView {
  const double* u1;
  const double* u0;
  const double* v0;
  const double* a0;
  const double* p1;  // nodal pressure control (nn)
  double dt;
  int is_static;  // static.py:105-124: u0 == u1, v0 = a0 = 0 -> no inertia / damping
  JacMix mix;     // matrix weights of the Jacobian-like output (jac_mix_du1 for d F_u / d u1)
};

VF_HD int find_slot(const int* bcol_i, int deg, int node) {
  int k = 0;
  while (k < deg - 1 && bcol_i[k] < node) ++k;
  return k;
}

template <int D>
VF_HD void load_cell(const MeshView& m, int e, int (&nd)[D + 1], double (&x)[D + 1][D]) {
  for (int a = 0; a <= D; ++a) {
    nd[a] = m.cells[a * m.ne + e];
    for (int c = 0; c < D; ++c) x[a][c] = m.xyz[c * m.nn + nd[a]];
  }
}

template <int D>
VF_HD void gather_vec(const double* v, const int (&nd)[D + 1], double (&out)[D + 1][D]) {
  for (int a = 0; a <= D; ++a)
    for (int c = 0; c < D; ++c) out[a][c] = v[D * nd[a] + c];
}

template <int D>
VF_HD void add_block(double* rowblk, int ld, int k, const double (&blk)[D][D], double scale) {
  for (int a = 0; a < D; ++a)
    for (int b = 0; b < D; ++b) rowblk[a * ld + k * D + b] += scale * blk[a][b];
}

// Exterior-facet terms of node i (follower pressure, contact, membrane) added to its block
// row / residual, then the Dirichlet rows (App. A.4).  Shared by all assembly kernels.
template <int D, bool JAC, bool RES>
VF_HD void assemble_node_facets_bc(int i, const MeshView& m, const PropView& p,
                                   const StateView& s, double* rowblk, double (&res)[D]) {
  const int b0 = m.brptr[i];
  const int deg = m.brptr[i + 1] - b0;
  const int ld = D * deg;
  const int* bcol_i = m.bcol + b0;
  // ---- 'pressure' exterior facets: follower pressure, contact, membrane ------------
  for (int t = m.n2f_ptr[i]; t < m.n2f_ptr[i + 1]; ++t) {
    const int ref = m.n2f[t];
    const int f = ref >> 2, a = ref & 3;
    const int e = m.pf_cell[f], o = m.pf_opp[f];
    int nd[D + 1];
    double x[D + 1][D];
    load_cell<D>(m, e, nd, x);
    CellGeo<D> g;
    p1_geometry(x, g);
    double N[D], meas;
    facet_geometry<D>(g, o, N, meas);
    double U[D + 1][D], gu[D][D];
    gather_vec<D>(s.u1, nd, U);
    grad_u<D>(g, U, gu);
    const double mw = meas / double(D * (D + 1));  // facet mass weight: mw (1 + delta_ab)

    if (a != o) {
      // int p phi_a ds
      double pw = 0.0;
      for (int b = 0; b <= D; ++b)
        if (b != o) pw += (a == b ? 2.0 : 1.0) * s.p1[nd[b]];
      pw *= mw;
      if (RES) {
        double c[D];
        cof_normal(gu, N, c);
        for (int k = 0; k < D; ++k) res[k] += pw * c[k];
      }
      if (JAC) {
        for (int b = 0; b <= D; ++b) {
          double dc[D][D];
          dcof_normal(gu, N, g.G[b], dc);
          add_block<D>(rowblk, ld, find_slot(bcol_i, deg, nd[b]), dc, pw * s.mix.p);
        }
      }
      if (p.contact) {
        const double yc = p.scal[SC_YCONTACT], kc = p.scal[SC_KCONTACT];
        for (int b = 0; b <= D; ++b) {
          if (b == o) continue;
          double gap = -yc;
          for (int k = 0; k < D; ++k) gap += (x[b][k] + U[b][k]) * p.scal[SC_NCONTACT + k];
          const double mab = mw * (a == b ? 2.0 : 1.0);
          if (RES) {
            const double pc = contact_pressure(gap, kc);  // tc_b = pc * n
            for (int k = 0; k < D; ++k) res[k] -= mab * pc * p.scal[SC_NCONTACT + k];
          }
          if (JAC) {
            const double dp = contact_dpressure(gap, kc);
            const int slot = find_slot(bcol_i, deg, nd[b]);
            for (int k = 0; k < D; ++k)
              rowblk[k * ld + slot * D + k] -= s.mix.k * mab * dp * p.scal[SC_NCONTACT + k];
          }
        }
      }
    }
    if (p.membrane) {
      double mu_m, lam_pp;
      membrane_coef(p.emod_m[e], p.nu_m[e], mu_m, lam_pp);
      const double w = p.th_m[e] * meas;
      if (RES) {
        double S[D][D];
        membrane_stress<D>(gu, N, mu_m, lam_pp, S);
        for (int k = 0; k < D; ++k) {
          double sacc = 0.0;
          for (int j = 0; j < D; ++j) sacc += S[k][j] * g.G[a][j];
          res[k] += w * sacc;
        }
      }
      if (JAC) {
        for (int b = 0; b <= D; ++b) {
          double blk[D][D];
          for (int j = 0; j < D; ++j) {
            double gb[D][D];
            for (int r = 0; r < D; ++r)
              for (int c = 0; c < D; ++c) gb[r][c] = (r == j) ? g.G[b][c] : 0.0;
            double S[D][D];
            membrane_stress<D>(gb, N, mu_m, lam_pp, S);
            for (int k = 0; k < D; ++k) {
              double sacc = 0.0;
              for (int c = 0; c < D; ++c) sacc += S[k][c] * g.G[a][c];
              blk[k][j] = sacc;
            }
          }
          add_block<D>(rowblk, ld, find_slot(bcol_i, deg, nd[b]), blk, w * s.mix.k);
        }
      }
    }
  }

  // ---- Dirichlet rows: zero row, unit diagonal, zero residual (App. A.4) ------------
  const int self = find_slot(bcol_i, deg, i);
  for (int a = 0; a < D; ++a) {
    if (m.bc[D * i + a]) {
      if (JAC && s.mix.bc) {
        for (int t = 0; t < ld; ++t) rowblk[a * ld + t] = 0.0;
        rowblk[a * ld + self * D + a] = 1.0;
      }
      if (RES) res[a] = 0.0;
    }
  }
}

template <int D, bool JAC, bool RES>
VF_HD void assemble_node(int i, const MeshView& m, const PropView& p, const StateView& s,
                         double* rowblk, double (&res)[D]) {
  const int b0 = m.brptr[i];
  const int deg = m.brptr[i + 1] - b0;
  const int ld = D * deg;
  const int* bcol_i = m.bcol + b0;
  if (JAC)
    for (int t = 0; t < D * ld; ++t) rowblk[t] = 0.0;
  if (RES)
    for (int c = 0; c < D; ++c) res[c] = 0.0;

  const LameFac lf = lame_fac(p.scal[SC_NU]);
  const Damping dp = prop_damping(p);
  const NewmarkCoef nc = newmark_coef(s.dt);

  // ---- cell integrals --------------------------------------------------------
  for (int t = m.n2e_ptr[i]; t < m.n2e_ptr[i + 1]; ++t) {
    const int ref = m.n2e[t];
    const int e = ref >> 2, a = ref & 3;
    int nd[D + 1];
    double x[D + 1][D];
    load_cell<D>(m, e, nd, x);
    CellGeo<D> g;
    p1_geometry(x, g);
    const CellCoef cf = cell_coef<D>(p.emod[e], lf, p.eta[e], p.rho[e], g.vol, dp);
    if (JAC) {
      for (int c = 0; c <= D; ++c) {
        double blk[D][D];
        cell_block<D>(g, cf, s.mix, a, c, blk);
        add_block<D>(rowblk, ld, find_slot(bcol_i, deg, nd[c]), blk, 1.0);
      }
    }
    if (RES) {
      double U[D + 1][D], V[D + 1][D], A[D + 1][D];
      for (int b = 0; b <= D; ++b)
        for (int c = 0; c < D; ++c) {
          const int dof = D * nd[b] + c;
          const double u1 = s.u1[dof], u0 = s.u0[dof], v0 = s.v0[dof], a0 = s.a0[dof];
          U[b][c] = u1;
          V[b][c] = s.is_static ? 0.0 : newmark_v(nc, u1, u0, v0, a0);
          A[b][c] = s.is_static ? 0.0 : newmark_a(nc, u1, u0, v0, a0);
        }
      double r[D];
      cell_residual<D>(g, cf, a, U, V, A, r);
      for (int c = 0; c < D; ++c) res[c] += r[c];
    }
  }

  assemble_node_facets_bc<D, JAC, RES>(i, m, p, s, rowblk, res);
}

// assemble_node<3> reading the gather tables of tet_tables.h: the same cells in the same order with
// the same arithmetic (identical bits), but 1 + 4 record loads per cell instead of 4 + 12 scattered
// ones, and the CSR slots from the pair's packed word instead of four scans of the column list.
template <bool JAC, bool RES>
VF_HD void assemble_node_tet(int i, const MeshView& m, const PropView& p, const StateView& s,
                             double* rowblk, double (&res)[3]) {
  constexpr int D = 3;
  const int b0 = m.brptr[i];
  const int deg = m.brptr[i + 1] - b0;
  const int ld = D * deg;
  if (JAC)
    for (int t = 0; t < D * ld; ++t) rowblk[t] = 0.0;
  if (RES)
    for (int c = 0; c < D; ++c) res[c] = 0.0;

  const LameFac lf = lame_fac(p.scal[SC_NU]);
  const Damping dp = prop_damping(p);
  const NewmarkCoef nc = newmark_coef(s.dt);
  const I4* cells4 = reinterpret_cast<const I4*>(m.cells4);

  for (int t = m.n2e_ptr[i]; t < m.n2e_ptr[i + 1]; ++t) {
    const int ref = m.n2e[t];
    const int e = ref >> 2, a = ref & 3;
    const unsigned slots = m.n2e_slots[t];
    const I4 c4 = cells4[e];
    const int nd[D + 1] = {c4.x, c4.y, c4.z, c4.w};
    double x[D + 1][D];
    for (int b = 0; b <= D; ++b) {
      const D2* q = reinterpret_cast<const D2*>(m.xyz4 + 4 * (size_t)nd[b]);
      const D2 q0 = q[0], q1 = q[1];
      x[b][0] = q0.x;
      x[b][1] = q0.y;
      x[b][2] = q1.x;
    }
    CellGeo<D> g;
    p1_geometry(x, g);
    const CellCoef cf = cell_coef<D>(p.emod[e], lf, p.eta[e], p.rho[e], g.vol, dp);
    if (JAC) {
      for (int c = 0; c <= D; ++c) {
        double blk[D][D];
        cell_block<D>(g, cf, s.mix, a, c, blk);
        add_block<D>(rowblk, ld, (int)((slots >> (8 * c)) & 255u), blk, 1.0);
      }
    }
    if (RES) {
      double U[D + 1][D], V[D + 1][D], A[D + 1][D];
      for (int b = 0; b <= D; ++b)
        for (int c = 0; c < D; ++c) {
          const int dof = D * nd[b] + c;
          const double u1 = s.u1[dof], u0 = s.u0[dof], v0 = s.v0[dof], a0 = s.a0[dof];
          U[b][c] = u1;
          V[b][c] = s.is_static ? 0.0 : newmark_v(nc, u1, u0, v0, a0);
          A[b][c] = s.is_static ? 0.0 : newmark_a(nc, u1, u0, v0, a0);
        }
      double r[D];
      cell_residual<D>(g, cf, a, U, V, A, r);
      for (int c = 0; c < D; ++c) res[c] += r[c];
    }
  }

  assemble_node_facets_bc<D, JAC, RES>(i, m, p, s, rowblk, res);
}

// assemble_node, or its table-driven form when the mesh view carries the tetrahedral gather tables
template <int D, bool JAC, bool RES>
VF_HD void assemble_node_auto(int i, const MeshView& m, const PropView& p, const StateView& s,
                              double* rowblk, double (&res)[D]) {
  if constexpr (D == 3) {
    if (m.n2e_slots) {
      assemble_node_tet<JAC, RES>(i, m, p, s, rowblk, res);
      return;
    }
  }
  assemble_node<D, JAC, RES>(i, m, p, s, rowblk, res);
}

}  // namespace vf
