// Node-centric fan assembly of P1 triangles (fp64): one caller owns vertex n, walks its adjacent
// cells counter-clockwise and produces the two scalar rows of n's block row of
// J_uu = d F_u / d u1 and the two residual entries F_u[n].
//
// Stands in for the same dfn.assemble calls as node_assembly.cuh
// (/root/reference/src/femvf/models/transient.py:363-406, models/assemblyutils.py:49-50,
// forms equations/form.py:516-533, 540-572, 965-990, 918-956, 1067-1113; closed forms in
// SURVEY.md App. A.3), cell integrals only -- the exterior-facet terms and Dirichlet rows are
// applied afterwards by assemble_node_facets_bc.
//
// Why a fan walk.  Cell j of vertex n is (n, p_j, p_{j+1}) in the cell's own counter-clockwise
// order, so (i) one NEW ring vertex is read per cell, (ii) the off-diagonal block (n, p_j) is the
// sum of the contributions of cells j-1 and j: it is completed in registers and stored once,
// straight to its place in the CSR array (no zero-fill, no read-modify-write, no atomics,
// bit-reproducible), (iii) nothing is exchanged between callers: no per-cell records, no second
// phase, no barrier between "cells" and "rows".  The price is that a cell's geometry is
// evaluated by each of its three vertices (about 20 of ~125 fp64 operations per visit).
//
// With unnormalised gradients g_a = det * G_a (det = 2 |K|) and h = 1 / (2 det):
//   |K| (lam G_a (x) G_c + mu G_c (x) G_a + mu (G_a . G_c) I) = h (lam g_a (x) g_c + ...)
//   |K| sigma(grad u) G_n = h sigma(H) g_n,   H = (U_p - U_n) (x) g_p + (U_q - U_n) (x) g_q
// (sum_a g_a = 0), so one reciprocal per visit is the only division.
//
// Everything is __host__ __device__: tests/hostcheck runs the same arithmetic on the CPU.
#pragma once

#include "elem.cuh"

namespace vf {

// Per-launch coefficients (material laws folded with the Jacobian mix and the damping model).
struct FanCoef {
  // Jacobian:  lam_c = emod jl,  mu_c = emod jm + eta jmv,  mass block = rho det jmass
  double jl, jm, jmv, jmass;
  // residual:  lam = emod rl, mu = emod rm, vlam = emod rvl, vmu = emod rvm_e + eta rvm_eta,
  //            lumped mass = rho det rmass (a_n + sum a), damping mass = rho det rvmass (...)
  double rl, rm, rvl, rvm_e, rvm_eta, rmass, rvmass;
};

VF_HD FanCoef fan_coef(const LameFac& lf, const Damping& dp, const JacMix& mix) {
  FanCoef c;
  const bool ray = dp.kind == DAMP_RAYLEIGH;
  // cell_coef / cell_block: lv = (k lam + c vlam) |K|, mv = (k mu + c vmu) |K|,
  // diagonal mass (m massv + c vmass)(1 + delta), massv = rho |K| / 12 = rho det / 24
  c.jl = mix.k * lf.lam_fac + (ray ? mix.c * dp.rk * lf.lam_fac : 0.0);
  c.jm = mix.k * lf.mu_fac + (ray ? mix.c * dp.rk * lf.mu_fac : 0.0);
  c.jmv = ray ? 0.0 : 0.5 * mix.c;  // Kelvin-Voigt: viscous stress eta eps(v) (form.py:984)
  c.jmass = (mix.m + (ray ? mix.c * dp.rm : 0.0)) / 24.0;
  c.rl = lf.lam_fac;
  c.rm = lf.mu_fac;
  c.rvl = ray ? dp.rk * lf.lam_fac : 0.0;
  c.rvm_e = ray ? dp.rk * lf.mu_fac : 0.0;
  c.rvm_eta = ray ? 0.0 : 0.5;
  c.rmass = 1.0 / 24.0;
  c.rvmass = ray ? dp.rm / 24.0 : 0.0;
  return c;
}

// Header and ring words of the fan tables (tables.build_fan_tables).
//   header: (brptr[n] - brptr[i0]) | deg << 12 | self_slot << 17 | ncell << 22 | closed << 27
//   ring:   staged vertex slot | CSR slot << 10 | local cell << 15
VF_HD int fan_hdr_b0(unsigned w) { return (int)(w & 0xfffu); }
VF_HD int fan_hdr_deg(unsigned w) { return (int)((w >> 12) & 0x1fu); }
VF_HD int fan_hdr_self(unsigned w) { return (int)((w >> 17) & 0x1fu); }
VF_HD int fan_hdr_ncell(unsigned w) { return (int)((w >> 22) & 0x1fu); }
VF_HD bool fan_hdr_closed(unsigned w) { return ((w >> 27) & 1u) != 0; }
VF_HD int fan_ent_vslot(unsigned w) { return (int)(w & 0x3ffu); }
VF_HD int fan_ent_cslot(unsigned w) { return (int)((w >> 10) & 0x1fu); }
VF_HD int fan_ent_cell(unsigned w) { return (int)((w >> 15) & 0xfffu); }

// 1 / (2 det): reciprocal seed + two Newton steps on the device (full double precision, no
// division slow path: ~5 instructions instead of ~10 and a branch); plain division on the host.
VF_HD double fan_half_rcp(double det) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(det));
  double e = fma(-det, r, 1.0);
  r = fma(r, e, r);
  e = fma(-det, r, 1.0);
  r = fma(r, e, r);
  return 0.5 * r;
#else
  return 0.5 / det;
#endif
}

struct FanBlock {
  double b00, b01, b10, b11;
};

// State of one ring vertex relative to the fan's centre n.
struct FanRing {
  double ex, ey;               // x_p - x_n
  double dux, duy, dvx, dvy;   // u_p - u_n, v_p - v_n  (u1 and v_nmk)
  double ax, ay;               // a_nmk at p
  int cs;                      // CSR slot of p in n's block row
};

struct FanAcc {
  double d00, d01, d11;  // diagonal block (symmetric), without ...
  double ds;             // ... its isotropic part, added to d00 and d11 at the end
  double r0, r1;         // residual
};

// One cell (n, p, q) of the fan: adds its share of the diagonal block and of the residual, ADDS its
// block (n, p) to P (which arrives holding the previous cell's share of that block) and returns
// its block (n, q) in Q.
// RAY = false drops the terms that exist only with Rayleigh damping (the stiffness-proportional
// viscous stress and the mass-proportional damping force): for the Kelvin-Voigt model their
// coefficients are zero, so the results are the same and ~10 of 128 operations are saved.
template <bool JAC, bool RES, bool RAY = true>
VF_HD void fan_cell(const FanRing& p, const FanRing& q, double emod, double eta, double rho,
                    const FanCoef& fc, const D2& vn, const D2& an, FanAcc& acc, FanBlock& P,
                    FanBlock& Q) {
  const double det = p.ex * q.ey - p.ey * q.ex;
  const double h = fan_half_rcp(det);
  // unnormalised gradients of p, q and n
  const double gpx = q.ey, gpy = -q.ex, gqx = -p.ey, gqy = p.ex;
  const double gnx = p.ey - q.ey, gny = q.ex - p.ex;
  if (JAC) {
    const double L = emod * fc.jl * h;
    const double M = (emod * fc.jm + eta * fc.jmv) * h;
    const double mb = rho * det * fc.jmass;
    const double anx = L * gnx, any = L * gny, bnx = M * gnx, bny = M * gny;
    // block (n, c)[i][k] = an_i gc_k + gc_i bn_k + delta_ik (bn . gc + mb (1 + delta_nc))
    // the isotropic part of the diagonal block is summed on its own and added once per node
    acc.ds += bnx * gnx + bny * gny + 2.0 * mb;
    acc.d00 = anx * gnx + (gnx * bnx + acc.d00);
    acc.d01 = anx * gny + (gnx * bny + acc.d01);
    acc.d11 = any * gny + (gny * bny + acc.d11);
    const double dgp = bnx * gpx + bny * gpy + mb;
    P.b00 = anx * gpx + (gpx * bnx + (dgp + P.b00));
    P.b01 = anx * gpy + (gpx * bny + P.b01);
    P.b10 = any * gpx + (gpy * bnx + P.b10);
    P.b11 = any * gpy + (gpy * bny + (dgp + P.b11));
    const double dgq = bnx * gqx + bny * gqy + mb;
    Q.b00 = anx * gqx + (gqx * bnx + dgq);
    Q.b01 = anx * gqy + gqx * bny;
    Q.b10 = any * gqx + gqy * bnx;
    Q.b11 = any * gqy + (gqy * bny + dgq);
  }
  if (RES) {
    // det * (mu grad u + vmu grad v) = dW_p (x) g_p + dW_q (x) g_q with dW = mu dU + vmu dV, and
    // det * div u for the isotropic part
    const double mu = emod * fc.rm, lam = emod * fc.rl;
    const double vmu = RAY ? emod * fc.rvm_e + eta * fc.rvm_eta : eta * fc.rvm_eta;
    const double wpx = mu * p.dux + vmu * p.dvx, wpy = mu * p.duy + vmu * p.dvy;
    const double wqx = mu * q.dux + vmu * q.dvx, wqy = mu * q.duy + vmu * q.dvy;
    const double t00 = wpx * gpx + wqx * gqx, t01 = wpx * gpy + wqx * gqy;
    const double t10 = wpy * gpx + wqy * gqx, t11 = wpy * gpy + wqy * gqy;
    const double divu = (p.dux * gpx + q.dux * gqx) + (p.duy * gpy + q.duy * gqy);
    const double iso = RAY ? lam * divu + (emod * fc.rvl) *
                                 ((p.dvx * gpx + q.dvx * gqx) + (p.dvy * gpy + q.dvy * gqy))
                           : lam * divu;
    const double s01 = t01 + t10;
    const double s00 = (t00 + t00) + iso, s11 = (t11 + t11) + iso;
    const double md = rho * det;
    const double ma = md * fc.rmass;
    // lumped sums over the cell's vertices with weight (1 + delta_an); v_p = dv_p + v_n
    const double sax = (an.x + an.x) + (p.ax + q.ax), say = (an.y + an.y) + (p.ay + q.ay);
    if (!RAY) {
      acc.r0 += h * (s00 * gnx + s01 * gny) + ma * sax;
      acc.r1 += h * (s01 * gnx + s11 * gny) + ma * say;
      return;
    }
    const double mv = md * fc.rvmass;
    const double svx = 4.0 * vn.x + (p.dvx + q.dvx), svy = 4.0 * vn.y + (p.dvy + q.dvy);
    acc.r0 += h * (s00 * gnx + s01 * gny) + (ma * sax + mv * svx);
    acc.r1 += h * (s01 * gnx + s11 * gny) + (ma * say + mv * svy);
  }
}

// ring(r)   -> word of row r of this node (0 = header, 1 + j = ring vertex j)
// vtx_xy(s) -> D2 coordinates of staged vertex s;  vtx_uva(s, u, v, a) -> nodal u1, v_nmk, a_nmk
// mat(c, emod, eta, rho) -> DG0 properties of the tile's local cell c
// Jtile: where the tile's slice of the CSR value array goes (the node's block row starts
// 4 (brptr[n] - brptr[i0]) doubles into it): shared memory in the kernel.
// The loop over the cells is unrolled by two with the roles of the two ring-vertex register sets
// swapped, so that "the new vertex becomes the previous one" costs no register moves.
template <bool JAC, bool RES, bool RAY = true, class Ring, class VtxXY, class VtxUVA, class Mat>
VF_HD void fan_walk_node(int nslot, const Ring& ring, const VtxXY& vtx_xy, const VtxUVA& vtx_uva,
                         const Mat& mat, const FanCoef& fc, double* Jtile, double* res_out) {
  const unsigned hdr = ring(0);
  const int deg = fan_hdr_deg(hdr), self = fan_hdr_self(hdr);
  const int ncell = fan_hdr_ncell(hdr);
  const bool closed = fan_hdr_closed(hdr);
  D2* row0 = reinterpret_cast<D2*>(Jtile) + 2 * fan_hdr_b0(hdr);
  D2* row1 = row0 + deg;

  const D2 xn = vtx_xy(nslot);
  D2 un = D2{0.0, 0.0}, vn = D2{0.0, 0.0}, an = D2{0.0, 0.0};
  if (RES) vtx_uva(nslot, un, vn, an);

  // loads ring entry `row` into rv and returns the local cell that FOLLOWS this vertex
  auto load_ring = [&](int row, FanRing& rv) -> int {
    const unsigned ent = ring(row);
    const int vs = fan_ent_vslot(ent);
    rv.cs = fan_ent_cslot(ent);
    const D2 x = vtx_xy(vs);
    rv.ex = x.x - xn.x;
    rv.ey = x.y - xn.y;
    if (RES) {
      D2 u, v, a;
      vtx_uva(vs, u, v, a);
      rv.dux = u.x - un.x; rv.duy = u.y - un.y;
      rv.dvx = v.x - vn.x; rv.dvy = v.y - vn.y;
      rv.ax = a.x; rv.ay = a.y;
    }
    return fan_ent_cell(ent);
  };
  FanRing A, B;
  A.dux = A.duy = A.dvx = A.dvy = A.ax = A.ay = 0.0;
  B = A;
  FanAcc acc = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  const FanBlock zero = {0.0, 0.0, 0.0, 0.0};
  FanBlock first = zero, cA, cB;
  double emod, eta, rho;
  auto put = [&](int slot, const FanBlock& b) {
    row0[slot] = D2{b.b00, b.b01};
    row1[slot] = D2{b.b10, b.b11};
  };

  // cell 0 (peeled): its (n, p_0) block waits for the end of the fan
  int cell = load_ring(1, A);
  const int cs_first = A.cs;
  mat(cell, emod, eta, rho);
  cell = load_ring(2, B);
  fan_cell<JAC, RES, RAY>(A, B, emod, eta, rho, fc, vn, an, acc, first, cB);
  int j = 1;
  for (; j + 1 < ncell; j += 2) {
    // cB holds cell j-1's share of block (n, B): cell j adds its own and the block is complete
    mat(cell, emod, eta, rho);
    cell = load_ring(2 + j, A);
    fan_cell<JAC, RES, RAY>(B, A, emod, eta, rho, fc, vn, an, acc, cB, cA);
    if (JAC) put(B.cs, cB);
    mat(cell, emod, eta, rho);
    cell = load_ring(3 + j, B);
    fan_cell<JAC, RES, RAY>(A, B, emod, eta, rho, fc, vn, an, acc, cA, cB);
    if (JAC) put(A.cs, cA);
  }
  int cs_last = B.cs;
  if (j < ncell) {  // odd remainder
    mat(cell, emod, eta, rho);
    load_ring(2 + j, A);
    fan_cell<JAC, RES, RAY>(B, A, emod, eta, rho, fc, vn, an, acc, cB, cA);
    if (JAC) put(B.cs, cB);
    cB = cA;
    cs_last = A.cs;
  }
  if (JAC) {
    if (closed) {  // the last cell meets the first: cs_last == cs_first
      row0[cs_first] = D2{first.b00 + cB.b00, first.b01 + cB.b01};
      row1[cs_first] = D2{first.b10 + cB.b10, first.b11 + cB.b11};
    } else {
      put(cs_first, first);
      put(cs_last, cB);
    }
    row0[self] = D2{acc.d00 + acc.ds, acc.d01};
    row1[self] = D2{acc.d01, acc.d11 + acc.ds};
  }
  if (RES) {
    res_out[0] = acc.r0;
    res_out[1] = acc.r1;
  }
}

}  // namespace vf
