// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
// Runs vf::assemble_node (the arithmetic the CUDA kernels execute) in a plain CPU loop so
// that the element math can be checked against the oracle without a GPU.
#include <cmath>
#include <cstdlib>
using std::sqrt; using std::fabs; using std::isinf;
#include "../../vf-fem_b200/csrc/node_assembly.cuh"
#include "../../vf-fem_b200/csrc/fan_assembly.cuh"

template <int D>
static void run(const vf::MeshView& m, const vf::PropView& p, const vf::StateView& s,
                double* J, double* F) {
  for (int i = 0; i < m.nn; ++i) {
    double res[D];
    vf::assemble_node<D, true, true>(i, m, p, s, J + D * D * m.brptr[i], res);
    for (int c = 0; c < D; ++c) F[D * i + c] = res[c];
  }
}

extern "C" int hostcheck_assemble(
    int dim, int nn, int ne, int nfp, const double* xyz, const int* cells, const int* brptr,
    const int* bcol, const int* n2e_ptr, const int* n2e, const int* n2f_ptr, const int* n2f,
    const int* pf_cell, const int* pf_opp, const unsigned char* bc, const double* rho,
    const double* eta, const double* emod, const double* scal, const double* emod_m,
    const double* nu_m, const double* th_m, int contact, int membrane, int damping,
    const double* u1,
    const double* u0, const double* v0, const double* a0, const double* p1, double dt,
    double* J, double* F) {
  vf::MeshView m{dim, nn, ne, nfp, xyz, nullptr, cells, brptr, bcol, n2e_ptr, n2e,
                 n2f_ptr, n2f, pf_cell, pf_opp, bc};
  vf::PropView p{rho, eta, emod, scal, emod_m, nu_m, th_m, contact, membrane, damping};
  vf::StateView s{u1, u0, v0, a0, p1, dt, 0, vf::jac_mix_du1(vf::newmark_coef(dt), false)};
  if (dim == 2) run<2>(m, p, s, J, F);
  else if (dim == 3) run<3>(m, p, s, J, F);
  else return 1;
  return 0;
}

// CPU emulation of asm_tile2_kernel + facet_bc_kernel (two-phase tile assembly, triangles):
// same device functions (tri_record, tri_row_block, assemble_node_facets_bc), same order.
extern "C" int hostcheck_assemble_tile2(
    int nn, int ne, int nfp, const double* xyz, const int* cells, const int* brptr,
    const int* bcol, const int* n2e_ptr, const int* n2e, const int* n2f_ptr, const int* n2f,
    const int* pf_cell, const int* pf_opp, const unsigned char* bc, const double* rho,
    const double* eta, const double* emod, const double* scal, const double* emod_m,
    const double* nu_m, const double* th_m, int contact, int membrane, int damping,
    const double* u1,
    const double* u0, const double* v0, const double* a0, const double* p1, double dt,
    int ntiles, const int* tile_start, const int* te_ptr, const int* te_elem,
    const unsigned* pair_info, int max_tile_elems, double* J, double* F) {
  vf::MeshView m{2, nn, ne, nfp, xyz, nullptr, cells, brptr, bcol, n2e_ptr, n2e,
                 n2f_ptr, n2f, pf_cell, pf_opp, bc};
  vf::PropView p{rho, eta, emod, scal, emod_m, nu_m, th_m, contact, membrane, damping};
  vf::StateView s{u1, u0, v0, a0, p1, dt, 0, vf::jac_mix_du1(vf::newmark_coef(dt), false)};
  const vf::LameFac lf = vf::lame_fac(scal[vf::SC_NU]);
  const vf::NewmarkCoef nc = vf::newmark_coef(dt);
  // 16-byte aligned like the shared-memory buffer of the kernel
  double* recs = static_cast<double*>(aligned_alloc(16, sizeof(double) * (size_t)max_tile_elems * vf::kRec2D));
  for (int t = 0; t < ntiles; ++t) {
    const int i0 = tile_start[t], i1 = tile_start[t + 1];
    const int q0 = te_ptr[t], q1 = te_ptr[t + 1];
    if (q1 - q0 > max_tile_elems) return 2;
    for (int q = q0; q < q1; ++q) {
      const int e = te_elem[q];
      int nd[3];
      double x[3][2];
      for (int a = 0; a < 3; ++a) {
        nd[a] = cells[a * ne + e];
        x[a][0] = xyz[nd[a]];
        x[a][1] = xyz[nn + nd[a]];
      }
      vf::tri_record(x, nd, emod[e], lf, eta[e], rho[e], vf::prop_damping(p), nc, false, true, u1,
                     u0, v0, a0,
                     recs + (size_t)(q - q0) * vf::kRec2D);
    }
    for (int i = i0; i < i1; ++i) {
      const int b0 = brptr[i], deg = brptr[i + 1] - b0;
      double* row0 = J + 4 * (size_t)b0;
      double* row1 = row0 + 2 * deg;
      for (int k = 0; k < 4 * deg; ++k) row0[k] = 0.0;
      double r0 = 0.0, r1 = 0.0;
      for (int q = n2e_ptr[i]; q < n2e_ptr[i + 1]; ++q) {
        const unsigned info = pair_info[q];
        const double* rec = recs + (size_t)(info & 0xfffu) * vf::kRec2D;
        const int a = (info >> 12) & 3;
        vf::D2 w0[3], w1[3];
        vf::tri_row_fan(rec, a, 0, w0[0], w0[1], w0[2]);
        vf::tri_row_fan(rec, a, 1, w1[0], w1[1], w1[2]);
        for (int c = 0; c < 3; ++c) {   // slots are (self, next, prev)
          const int slot = (info >> (14 + 6 * c)) & 63;
          row0[2 * slot] += w0[c].x;
          row0[2 * slot + 1] += w0[c].y;
          row1[2 * slot] += w1[c].x;
          row1[2 * slot + 1] += w1[c].y;
        }
        r0 += rec[9 + 2 * a];
        r1 += rec[10 + 2 * a];
      }
      F[2 * i] = r0;
      F[2 * i + 1] = r1;
    }
  }
  free(recs);
  for (int i = 0; i < nn; ++i) {
    bool touch = n2f_ptr[i + 1] > n2f_ptr[i] || bc[2 * i] || bc[2 * i + 1];
    if (!touch) continue;
    double res[2] = {F[2 * i], F[2 * i + 1]};
    vf::assemble_node_facets_bc<2, true, true>(i, m, p, s, J + 4 * (size_t)brptr[i], res);
    F[2 * i] = res[0];
    F[2 * i + 1] = res[1];
  }
  return 0;
}

// CPU emulation of asm_fan_kernel + facet_bc_kernel (node-centric fan assembly, triangles): the
// same device function (fan_walk_node) driven by the product's fan tables, tile by tile, with
// the staged vertex list (own vertices, then the tile's halo), the tile-ordered property
// blocks and the tile's CSR slice rebuilt like the kernel does.
extern "C" int hostcheck_assemble_fan(
    int nn, int ne, int nfp, const double* xyz, const int* cells, const int* brptr,
    const int* bcol, const int* n2e_ptr, const int* n2e, const int* n2f_ptr, const int* n2f,
    const int* pf_cell, const int* pf_opp, const unsigned char* bc, const double* rho,
    const double* eta, const double* emod, const double* scal, const double* emod_m,
    const double* nu_m, const double* th_m, int contact, int membrane, int damping,
    const double* u1, const double* u0, const double* v0, const double* a0, const double* p1,
    double dt, int is_static, int tile_nodes, int ntiles, const int* desc, const unsigned* ring,
    const int* halo, const int* tcell, int max_verts, int max_rows, int max_cells,
    int max_blocks, double* J, double* F) {
  vf::MeshView m{2, nn, ne, nfp, xyz, nullptr, cells, brptr, bcol, n2e_ptr, n2e,
                 n2f_ptr, n2f, pf_cell, pf_opp, bc};
  vf::PropView p{rho, eta, emod, scal, emod_m, nu_m, th_m, contact, membrane, damping};
  const vf::NewmarkCoef nc = vf::newmark_coef(dt);
  const vf::JacMix mix = vf::jac_mix_du1(nc, is_static != 0);
  vf::StateView s{u1, is_static ? u1 : u0, v0, a0, p1, dt, is_static, mix};
  const vf::FanCoef fc = vf::fan_coef(vf::lame_fac(scal[vf::SC_NU]), vf::prop_damping(p), mix);
  const int TN = tile_nodes;
  vf::D2* sxy = new vf::D2[max_verts];
  vf::NodeUVA* suva = new vf::NodeUVA[max_verts];
  double* smat = new double[3 * (size_t)max_cells];
  double* sJ = static_cast<double*>(aligned_alloc(16, sizeof(double) * 4 * (size_t)max_blocks));
  int rc = 0;
  for (int t = 0; t < ntiles && rc == 0; ++t) {
    const int* d = desc + 12 * t;
    const int i0 = d[0], nT = d[1] & 0xffff, nH = (int)((unsigned)d[1] >> 16);
    const int h0 = d[2], ring0 = d[3], rows = d[4], tc0 = d[5], ncp = d[6], bbase = d[7];
    const int nblk = d[8];
    if (nT + nH > max_verts || rows > max_rows || ncp > max_cells || nblk > max_blocks ||
        bbase != brptr[i0] || nblk != brptr[i0 + nT] - bbase) {
      rc = 2;
      break;
    }
    for (int k = 0; k < nT + nH; ++k) {
      const int v = k < nT ? i0 + k : halo[h0 + k - nT];
      sxy[k] = vf::D2{xyz[v], xyz[nn + v]};
      suva[k] = vf::gather_node_uva(nc, is_static != 0, v, u1, s.u0, v0, a0);
    }
    for (int k = 0; k < ncp; ++k) {   // fan_pack_kernel + the bulk copy of the tile's block
      const int e = tcell[tc0 + k];
      smat[k] = emod[e];
      smat[ncp + k] = eta[e];
      smat[2 * ncp + k] = rho[e];
    }
    for (int k = 0; k < 4 * nblk; ++k) sJ[k] = std::nan("");
    const unsigned* ring_t = ring + (size_t)ring0;
    for (int k = 0; k < nT; ++k) {
      auto ringf = [&](int r) {
        if (r >= rows) std::abort();
        return ring_t[(size_t)r * TN + k];
      };
      auto vxy = [&](int sl) { return sxy[sl]; };
      auto vuva = [&](int sl, vf::D2& u, vf::D2& v, vf::D2& a) {
        u = suva[sl].u; v = suva[sl].v; a = suva[sl].a;
      };
      auto mat = [&](int c, double& em, double& et, double& rh) {
        if (c >= ncp) std::abort();
        em = smat[c]; et = smat[ncp + c]; rh = smat[2 * ncp + c];
      };
      double res[2];
      vf::fan_walk_node<true, true>(k, ringf, vxy, vuva, mat, fc, sJ, res);
      F[2 * (i0 + k)] = res[0];
      F[2 * (i0 + k) + 1] = res[1];
    }
    for (int k = 0; k < 4 * nblk; ++k) J[4 * (size_t)bbase + k] = sJ[k];   // the bulk store
  }
  delete[] sxy;
  delete[] suva;
  delete[] smat;
  free(sJ);
  if (rc) return rc;
  for (int i = 0; i < nn; ++i) {
    bool touch = n2f_ptr[i + 1] > n2f_ptr[i] || bc[2 * i] || bc[2 * i + 1];
    if (!touch) continue;
    double res[2] = {F[2 * i], F[2 * i + 1]};
    vf::assemble_node_facets_bc<2, true, true>(i, m, p, s, J + 4 * (size_t)brptr[i], res);
    F[2 * i] = res[0];
    F[2 * i + 1] = res[1];
  }
  return 0;
}

// CPU run of the second P2 triangle kernel (p2_assemble_warp_kernel, csrc/p2.cu): the same
// per-node function (vf::p2_node_row) in a loop over the class-sorted thread -> node map, with
// the kernel's staging (64 rows at an odd stride, then 4 * deg consecutive doubles copied to the
// CSR array per row).
#define __device__
#define __constant__
#include "../../vf-fem_b200/csrc/p2_tables.h"
#undef __device__
#undef __constant__
#include "../../vf-fem_b200/csrc/p2_node.cuh"

extern "C" int hostcheck_p2_assemble(
    int nn, const double* xy, const int* cells6, const int* brptr, const int* bcol,
    const int* n2e_ptr, const int* n2e, const unsigned* n2e_slots, const int* n2f_ptr,
    const int* n2f, const int* n2f_pair, const int* pf_cell, const int* pf_loc,
    const double* pf_geo, const unsigned char* fixed, const int* order, int n_class0,
    const double* emod, const double* eta, const double* rho, double nu, const double* u1,
    const double* u0, const double* v0, const double* a0, const double* p1, double dt, int flags,
    double* J, double* F) {
  vf::P2View V{xy, cells6, brptr, bcol, n2e_ptr, n2e, n2e_slots, n2f_ptr, n2f, n2f_pair,
               pf_cell, pf_loc, pf_geo, fixed};
  vf::P2Args A;
  A.emod = emod; A.eta = eta; A.rho = rho; A.u1 = u1; A.u0 = u0; A.v0 = v0; A.a0 = a0; A.p1 = p1;
  A.F = F; A.J = J; A.nu = nu; A.dt = dt;
  A.res = flags & 1;
  A.jac = (flags & 2) != 0;
  double* uva = nullptr;
  if (flags & 4) {  // the pre-pass of version 2: packed (u1, v_nmk, a_nmk) per node
    uva = static_cast<double*>(aligned_alloc(16, sizeof(double) * 6 * (size_t)nn));
    for (int n = 0; n < nn; ++n) vf::p2_pack_state(A, vf::newmark_coef(dt), n, uva);
    A.uva = uva;
  }
  const int first[2] = {0, n_class0}, count[2] = {n_class0, nn - n_class0};
  for (int c = 0; c < 2; ++c) {
    int max_deg = 1;
    for (int t = 0; t < count[c]; ++t) {
      const int i = order[first[c] + t];
      max_deg = std::max(max_deg, brptr[i + 1] - brptr[i]);
    }
    const int stride = (4 * max_deg) | 1, block = 64;
    double* s_rows = static_cast<double*>(aligned_alloc(16, sizeof(double) * ((size_t)block * stride + 2)));
    for (int blk = 0; blk * block < count[c]; ++blk) {
      int b0[64], deg[64];
      for (int tid = 0; tid < block; ++tid) {
        const int tix = blk * block + tid;
        b0[tid] = deg[tid] = 0;
        if (tix >= count[c]) continue;
        const int i = order[first[c] + tix];
        b0[tid] = brptr[i];
        deg[tid] = brptr[i + 1] - brptr[i];
        double r0, r1;
        if (c == 0) vf::p2_node_row<0>(V, A, &vf::kP2W[0][0][0][0], &vf::kP2M[0][0], i, s_rows + (size_t)tid * stride, r0, r1);
        else vf::p2_node_row<1>(V, A, &vf::kP2W[0][0][0][0], &vf::kP2M[0][0], i, s_rows + (size_t)tid * stride, r0, r1);
        if (A.res) {
          F[2 * i] = r0;
          F[2 * i + 1] = r1;
        }
      }
      if (A.jac)
        for (int tid = 0; tid < block; ++tid) {  // a warp copies its 32 rows one after another
          const double* rs = s_rows + (size_t)tid * stride;
          double* out = J + (size_t)4 * b0[tid];
          for (int j = 0; j < 4 * deg[tid]; ++j) out[j] = rs[j];
        }
    }
    free(s_rows);
  }
  free(uva);
  return 0;
}

// CPU run of the table-driven tetrahedral node assembly (vf::assemble_node_tet via
// assemble_node_auto) with the gather tables of csrc/tet_tables.h built like the engine does.
#include "../../vf-fem_b200/csrc/tet_tables.h"

extern "C" int hostcheck_assemble_tet_tables(
    int nn, int ne, int nfp, const double* xyz, const int* cells, const int* brptr,
    const int* bcol, const int* n2e_ptr, const int* n2e, const int* n2f_ptr, const int* n2f,
    const int* pf_cell, const int* pf_opp, const unsigned char* bc, const double* rho,
    const double* eta, const double* emod, const double* scal, const double* emod_m,
    const double* nu_m, const double* th_m, int contact, int membrane, int damping,
    const double* u1, const double* u0, const double* v0, const double* a0, const double* p1,
    double dt, double* J, double* F) {
  std::vector<int32_t> cells4;
  std::vector<double> xyz4;
  std::vector<uint32_t> slots;
  if (!vf::build_tet_gather_tables(nn, ne, xyz, cells, brptr, bcol, n2e_ptr, n2e, cells4, xyz4,
                                   slots))
    return 2;
  // 32-byte aligned copies, as on the device
  int* c4 = static_cast<int*>(aligned_alloc(32, (cells4.size() * sizeof(int) + 31) / 32 * 32));
  double* x4 = static_cast<double*>(aligned_alloc(32, xyz4.size() * sizeof(double)));
  std::copy(cells4.begin(), cells4.end(), c4);
  std::copy(xyz4.begin(), xyz4.end(), x4);
  vf::MeshView m{3, nn, ne, nfp, xyz, nullptr, cells, brptr, bcol, n2e_ptr, n2e,
                 n2f_ptr, n2f, pf_cell, pf_opp, bc};
  m.cells4 = c4;
  m.xyz4 = x4;
  m.n2e_slots = slots.data();
  vf::PropView p{rho, eta, emod, scal, emod_m, nu_m, th_m, contact, membrane, damping};
  vf::StateView s{u1, u0, v0, a0, p1, dt, 0, vf::jac_mix_du1(vf::newmark_coef(dt), false)};
  for (int i = 0; i < nn; ++i) {
    double res[3];
    vf::assemble_node_auto<3, true, true>(i, m, p, s, J + 9 * (size_t)brptr[i], res);
    for (int c = 0; c < 3; ++c) F[3 * i + c] = res[c];
  }
  free(c4);
  free(x4);
  return 0;
}
