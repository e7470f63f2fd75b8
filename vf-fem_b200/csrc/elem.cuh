// Element mathematics of the Newmark Kelvin-Voigt solid on P1 simplices (fp64).
//
// Stands in for the FFC/UFLACS-generated tabulate_tensor of the reference's UFL forms
// (/root/reference/src/femvf/equations/form.py:516-533 inertia, :540-572 elastic,
//  :965-990 Kelvin-Voigt, :733-756 follower pressure, :759-794 contact traction,
//  :800-855 membrane, :1067-1113 Newmark substitution; equations/uflcontinuum.py:9-26,
//  :73-88, :172-186).  Closed forms: SURVEY.md App. A.3.
//
// Everything here is __host__ __device__ so that tests/hostcheck can run the very same
// arithmetic on the CPU against the oracle.  The product path only ever calls it from
// CUDA kernels.
#pragma once

#if defined(__CUDACC__)
#define VF_HD __host__ __device__ __forceinline__
#else
#define VF_HD inline
#endif

namespace vf {

// Newmark constants of the reference (form.py:1083-1084): gamma = 1/2, beta = 1/4.
constexpr double kGamma = 0.5;
constexpr double kBeta = 0.25;

VF_HD double newmark_cv(double dt) { return kGamma / kBeta / dt; }          // d v1 / d u1
VF_HD double newmark_ca(double dt) { return 1.0 / kBeta / (dt * dt); }      // d a1 / d u1

// newmark.py:8-29
VF_HD double newmark_v(double u1, double u0, double v0, double a0, double dt) {
  return kGamma / kBeta / dt * (u1 - u0) - (kGamma / kBeta - 1.0) * v0 -
         dt * (kGamma / 2.0 / kBeta - 1.0) * a0;
}
// newmark.py:57-73
VF_HD double newmark_a(double u1, double u0, double v0, double a0, double dt) {
  return 1.0 / kBeta / (dt * dt) * (u1 - u0 - dt * v0) - (1.0 / 2.0 / kBeta - 1.0) * a0;
}

// Newmark update coefficients hoisted out of the element loops.  Each member is evaluated
// with exactly the operations of newmark.py:8-73, so v/a are bit-identical to calling
// newmark_v / newmark_a per entry.
struct NewmarkCoef {
  double cv, c_v0, c_a0v;  // v1 = cv (u1-u0) - c_v0 v0 - c_a0v a0
  double ca, c_a0a;        // a1 = ca (u1-u0-dt v0) - c_a0a a0
  double dt;
};

VF_HD NewmarkCoef newmark_coef(double dt) {
  NewmarkCoef c;
  c.cv = kGamma / kBeta / dt;
  c.c_v0 = kGamma / kBeta - 1.0;
  c.c_a0v = dt * (kGamma / 2.0 / kBeta - 1.0);
  c.ca = 1.0 / kBeta / (dt * dt);
  c.c_a0a = 1.0 / 2.0 / kBeta - 1.0;
  c.dt = dt;
  return c;
}
VF_HD double newmark_v(const NewmarkCoef& c, double u1, double u0, double v0, double a0) {
  return c.cv * (u1 - u0) - c.c_v0 * v0 - c.c_a0v * a0;
}
VF_HD double newmark_a(const NewmarkCoef& c, double u1, double u0, double v0, double a0) {
  return c.ca * (u1 - u0 - c.dt * v0) - c.c_a0a * a0;
}

// Weights of the matrices that make up a Jacobian-like assembly:
//   k K + c C + m M + p K_p, with Dirichlet rows applied when bc != 0.
// d F_u / d u1 is (1, cv, ca, 1, bc) (App. A.3); the state0 sensitivities (transient.py:408-421)
// are other mixes of the same matrices (C = d F_u / d v1, M = d F_u / d a1) without BCs.
struct JacMix {
  double k, c, m, p;
  int bc;
};
VF_HD JacMix jac_mix_du1(const NewmarkCoef& nc, bool is_static) {
  JacMix x;
  x.k = 1.0;
  x.c = is_static ? 0.0 : nc.cv;
  x.m = is_static ? 0.0 : nc.ca;
  x.p = 1.0;
  x.bc = 1;
  return x;
}

// Lame factors per unit modulus: lambda = emod * lam_fac, mu = emod * mu_fac
// (uflcontinuum.py:22-23); nu is a constant, so the divisions are done once.
struct LameFac {
  double lam_fac, mu_fac;
};
VF_HD LameFac lame_fac(double nu) {
  LameFac f;
  f.lam_fac = nu / (1.0 + nu) / (1.0 - 2.0 * nu);
  f.mu_fac = 1.0 / 2.0 / (1.0 + nu);
  return f;
}

template <int D>
struct CellGeo {
  double G[D + 1][D];  // constant shape-function gradients
  double vol;          // cell measure |K|
};

VF_HD void p1_geometry(const double (&x)[3][2], CellGeo<2>& g) {
  const double e1x = x[1][0] - x[0][0], e1y = x[1][1] - x[0][1];
  const double e2x = x[2][0] - x[0][0], e2y = x[2][1] - x[0][1];
  const double det = e1x * e2y - e1y * e2x;
  const double inv = 1.0 / det;
  g.vol = 0.5 * det;
  g.G[1][0] = e2y * inv;
  g.G[1][1] = -e2x * inv;
  g.G[2][0] = -e1y * inv;
  g.G[2][1] = e1x * inv;
  g.G[0][0] = -(g.G[1][0] + g.G[2][0]);
  g.G[0][1] = -(g.G[1][1] + g.G[2][1]);
}

VF_HD void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

VF_HD void p1_geometry(const double (&x)[4][3], CellGeo<3>& g) {
  double e1[3], e2[3], e3[3], c23[3], c31[3], c12[3];
  for (int i = 0; i < 3; ++i) {
    e1[i] = x[1][i] - x[0][i];
    e2[i] = x[2][i] - x[0][i];
    e3[i] = x[3][i] - x[0][i];
  }
  cross3(e2, e3, c23);
  cross3(e3, e1, c31);
  cross3(e1, e2, c12);
  const double det = e1[0] * c23[0] + e1[1] * c23[1] + e1[2] * c23[2];
  const double inv = 1.0 / det;
  g.vol = det / 6.0;
  for (int i = 0; i < 3; ++i) {
    g.G[1][i] = c23[i] * inv;
    g.G[2][i] = c31[i] * inv;
    g.G[3][i] = c12[i] * inv;
    g.G[0][i] = -(g.G[1][i] + g.G[2][i] + g.G[3][i]);
  }
}

// Damping model of the residual: Kelvin-Voigt (form.py:965-990: stress eta*eps(v)) or Rayleigh
// (form.py:918-956: rayleigh_m*rho*v body force + rayleigh_k*sigma_iso(eps(v))).
enum DampingKind { DAMP_KELVIN_VOIGT = 0, DAMP_RAYLEIGH = 1 };
struct Damping {
  int kind;
  double rm, rk;  // Rayleigh mass / stiffness factors
};

// Per-cell material factors already multiplied by |K|.  The viscous stress is written as
// 2 vmu eps(v) + vlam tr(eps(v)) I and the viscous body force as vmass-weighted nodal v, which
// covers both damping models.
struct CellCoef {
  double lamv;   // lambda |K|
  double muv;    // mu |K|
  double vlam;   // viscous lambda |K|   (Kelvin-Voigt: 0;          Rayleigh: rk lambda |K|)
  double vmu;    // viscous mu |K|       (Kelvin-Voigt: eta/2 |K|;  Rayleigh: rk mu |K|)
  double massv;  // rho |K| / ((d+1)(d+2))
  double vmass;  // damping mass         (Kelvin-Voigt: 0;          Rayleigh: rm massv)
};

template <int D>
VF_HD CellCoef cell_coef(double emod, const LameFac& lf, double eta, double rho, double vol,
                         const Damping& dp) {
  CellCoef c;
  c.lamv = emod * lf.lam_fac * vol;
  c.muv = emod * lf.mu_fac * vol;
  c.massv = rho * vol / double((D + 1) * (D + 2));
  if (dp.kind == DAMP_RAYLEIGH) {
    c.vlam = dp.rk * c.lamv;
    c.vmu = dp.rk * c.muv;
    c.vmass = dp.rm * c.massv;
  } else {
    c.vlam = 0.0;
    c.vmu = 0.5 * eta * vol;  // viscous stress is eta*eps(v), not 2 eta (form.py:984)
    c.vmass = 0.0;
  }
  return c;
}

// Block (a, c) of d F_u / d u1 from the cell integrals: 4M/dt^2 + 2C/dt + K  (App. A.3).
template <int D>
VF_HD void cell_block(const CellGeo<D>& g, const CellCoef& cf, const JacMix& mix, int a,
                      int c, double (&blk)[D][D]) {
  const double cv = mix.c, ca = mix.m;
  double gg = 0.0;
  for (int i = 0; i < D; ++i) gg += g.G[a][i] * g.G[c][i];
  const double lv = mix.k * cf.lamv + cv * cf.vlam;  // (mix.k == 1 reproduces lamv exactly)
  const double mv = mix.k * cf.muv + cv * cf.vmu;
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j)
      blk[i][j] = lv * g.G[a][i] * g.G[c][j] + mv * g.G[c][i] * g.G[a][j];
  const double dg = mv * gg + (ca * cf.massv + cv * cf.vmass) * (a == c ? 2.0 : 1.0);
  for (int i = 0; i < D; ++i) blk[i][i] += dg;
}

// Residual of the cell integrals at local node a.  U, V, A: nodal u1, v_nmk, a_nmk.
template <int D>
VF_HD void cell_residual(const CellGeo<D>& g, const CellCoef& cf, int a,
                         const double (&U)[D + 1][D], const double (&V)[D + 1][D],
                         const double (&A)[D + 1][D], double (&r)[D]) {
  double gu[D][D], gv[D][D];
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) {
      double su = 0.0, sv = 0.0;
      for (int b = 0; b <= D; ++b) {
        su += U[b][i] * g.G[b][j];
        sv += V[b][i] * g.G[b][j];
      }
      gu[i][j] = su;
      gv[i][j] = sv;
    }
  double tr = 0.0, trv = 0.0;
  for (int i = 0; i < D; ++i) {
    tr += gu[i][i];
    trv += gv[i][i];
  }
  for (int i = 0; i < D; ++i) {
    double s = 0.0;
    for (int j = 0; j < D; ++j) {
      double sig = cf.muv * (gu[i][j] + gu[j][i]) + cf.vmu * (gv[i][j] + gv[j][i]);
      if (i == j) sig += cf.lamv * tr + cf.vlam * trv;
      s += sig * g.G[a][j];
    }
    double m = 0.0, mvel = 0.0;
    for (int b = 0; b <= D; ++b) {
      m += (a == b ? 2.0 : 1.0) * A[b][i];
      mvel += (a == b ? 2.0 : 1.0) * V[b][i];
    }
    r[i] = s + cf.massv * m + cf.vmass * mvel;
  }
}

// --- exterior-facet terms ---------------------------------------------------------
// A facet is stored as (parent cell, local vertex o opposite the facet).  Its outward
// unit normal is -G_o / |G_o| and its measure is d |K| |G_o|.
template <int D>
VF_HD void facet_geometry(const CellGeo<D>& g, int o, double (&N)[D], double& meas) {
  double n2 = 0.0;
  for (int i = 0; i < D; ++i) n2 += g.G[o][i] * g.G[o][i];
  const double nrm = sqrt(n2);
  for (int i = 0; i < D; ++i) N[i] = -g.G[o][i] / nrm;
  meas = double(D) * g.vol * nrm;
}

// grad u1 from the nodal values of the cell
template <int D>
VF_HD void grad_u(const CellGeo<D>& g, const double (&U)[D + 1][D], double (&gu)[D][D]) {
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) {
      double s = 0.0;
      for (int b = 0; b <= D; ++b) s += U[b][i] * g.G[b][j];
      gu[i][j] = s;
    }
}

// c = cof(I + grad u) N      (uflcontinuum.py:172-186)
VF_HD void cof_normal(const double (&gu)[2][2], const double (&N)[2], double (&c)[2]) {
  const double F00 = 1.0 + gu[0][0], F01 = gu[0][1], F10 = gu[1][0], F11 = 1.0 + gu[1][1];
  c[0] = F11 * N[0] - F10 * N[1];
  c[1] = -F01 * N[0] + F00 * N[1];
}

VF_HD void cof_normal(const double (&gu)[3][3], const double (&N)[3], double (&c)[3]) {
  double F[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) F[i][j] = gu[i][j] + (i == j ? 1.0 : 0.0);
  for (int i = 0; i < 3; ++i) {
    double row[3];
    cross3(F[(i + 1) % 3], F[(i + 2) % 3], row);
    c[i] = row[0] * N[0] + row[1] * N[1] + row[2] * N[2];
  }
}

// d c / d U_b for c = cof(F) N  (App. A.3): 2D  t_b [[0,1],[-1,0]];  3D  -[F (N x G_b)]_x
VF_HD void dcof_normal(const double (&gu)[2][2], const double (&N)[2], const double* Gb,
                       double (&dc)[2][2]) {
  (void)gu;
  const double t = Gb[1] * N[0] - Gb[0] * N[1];
  dc[0][0] = 0.0;
  dc[0][1] = t;
  dc[1][0] = -t;
  dc[1][1] = 0.0;
}

VF_HD void dcof_normal(const double (&gu)[3][3], const double (&N)[3], const double* Gb,
                       double (&dc)[3][3]) {
  double nxg[3], q[3];
  cross3(N, Gb, nxg);
  for (int i = 0; i < 3; ++i) {
    double s = 0.0;
    for (int j = 0; j < 3; ++j) s += (gu[i][j] + (i == j ? 1.0 : 0.0)) * nxg[j];
    q[i] = s;
  }
  dc[0][0] = 0.0;
  dc[0][1] = q[2];
  dc[0][2] = -q[1];
  dc[1][0] = -q[2];
  dc[1][1] = 0.0;
  dc[1][2] = q[0];
  dc[2][0] = q[1];
  dc[2][1] = -q[0];
  dc[2][2] = 0.0;
}

// Cubic contact penalty (form.py:1173-1202).
VF_HD double positive_gap(double gap) {
  double pg = (gap + fabs(gap)) / 2.0;
  // gap == -inf gives nan above; the reference maps it to 0 (form.py:1184)
  if (gap < 0.0 && isinf(gap)) pg = 0.0;
  return pg;
}
// magnitude of tc = -k (g+)^3 n  ->  returns -k (g+)^3
VF_HD double contact_pressure(double gap, double k) {
  const double pg = positive_gap(gap);
  return -k * pg * pg * pg;
}
// the reference's per-DOF derivative factor: d tc_j / d u_j = -3 k (g+)^2 sign(g) n_j
VF_HD double contact_dpressure(double gap, double k) {
  const double pg = positive_gap(gap);
  const double sg = (gap > 0.0) ? 1.0 : ((gap < 0.0) ? -1.0 : 0.0);
  return -3.0 * k * pg * pg * sg;
}

// Membrane (form.py:812-855): with P = I - n n^T (n embedded in 3D, zero z in 2D),
//   S = 2 mu_m P eps P + lambda_pp tr(P eps P) P,   R_a = th |f| (P S P)[:d,:d] G_a .
// Since P is a projector, P S P = S.  Computes the d x d upper-left part of S for a given
// (not necessarily symmetric) displacement gradient gu.
template <int D>
VF_HD void membrane_stress(const double (&gu)[D][D], const double (&N)[D], double mu_m,
                           double lam_pp, double (&S)[D][D]) {
  double P[3][3], e[3][3], t[3][3], epp[3][3];
  double n3[3] = {0.0, 0.0, 0.0};
  for (int i = 0; i < D; ++i) n3[i] = N[i];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      P[i][j] = (i == j ? 1.0 : 0.0) - n3[i] * n3[j];
      e[i][j] = (i < D && j < D) ? 0.5 * (gu[i][j] + gu[j][i]) : 0.0;
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += P[i][k] * e[k][j];
      t[i][j] = s;
    }
  double tr = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += t[i][k] * P[k][j];
      epp[i][j] = s;
      if (i == j) tr += s;
    }
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) S[i][j] = 2.0 * mu_m * epp[i][j] + lam_pp * tr * P[i][j];
}

VF_HD void membrane_coef(double emod_m, double nu_m, double& mu_m, double& lam_pp) {
  mu_m = emod_m / 2.0 / (1.0 + nu_m);
  const double lam = emod_m * nu_m / (1.0 + nu_m) / (1.0 - 2.0 * nu_m);
  lam_pp = (emod_m == 0.0) ? 0.0 : 2.0 * mu_m * lam / (lam + 2.0 * mu_m);  // form.py:848-850
}


// ---- two-phase tile assembly (2D): per-element record + per-row accumulation -------------
// Record of one P1 triangle in shared memory:
//   [0..5]  G_a (a = 0,1,2; x,y)   [6] (lambda + cv lambda_v)|K|   [7] (mu + cv mu_v)|K|
//   [8]  (ca rho + cv rm rho)|K|/12   [9..14] cell residual at the 3 nodes   [15..17] pad
// The stride is 18 doubles = 144 B = 36 banks: 16-byte accesses of 8 consecutive records
// (a quarter warp) fall in 8 disjoint bank quads, where a 128-byte stride would put every
// record on the same banks (32-way conflict).
constexpr int kRec2D = 18;

struct
#if defined(__CUDACC__)
    __align__(16)
#else
    alignas(16)
#endif
        D2 {
  double x, y;
};

// Nodal state of one vertex as the residual needs it: u1 and the Newmark velocity /
// acceleration (form.py:1107-1111), each an (x, y) pair.
struct NodeUVA {
  D2 u, v, a;
};

VF_HD NodeUVA node_uva(const NewmarkCoef& nc, bool is_static, const D2& p1, const D2& p0,
                       const D2& pv, const D2& pa) {
  NodeUVA r;
  r.u = p1;
  if (is_static) {
    r.v = D2{0.0, 0.0};
    r.a = D2{0.0, 0.0};
  } else {
    r.v = D2{newmark_v(nc, p1.x, p0.x, pv.x, pa.x), newmark_v(nc, p1.y, p0.y, pv.y, pa.y)};
    r.a = D2{newmark_a(nc, p1.x, p0.x, pv.x, pa.x), newmark_a(nc, p1.y, p0.y, pv.y, pa.y)};
  }
  return r;
}

// Gather a vertex's state from the interleaved global vectors with 16-byte loads.
VF_HD NodeUVA gather_node_uva(const NewmarkCoef& nc, bool is_static, int node, const double* u1,
                              const double* u0, const double* v0, const double* a0) {
  const D2 p1 = reinterpret_cast<const D2*>(u1)[node];
  D2 p0 = D2{0.0, 0.0}, pv = D2{0.0, 0.0}, pa = D2{0.0, 0.0};
  if (!is_static) {
    p0 = reinterpret_cast<const D2*>(u0)[node];
    pv = reinterpret_cast<const D2*>(v0)[node];
    pa = reinterpret_cast<const D2*>(a0)[node];
  }
  return node_uva(nc, is_static, p1, p0, pv, pa);
}

// Record of one triangle.  `fetch(a)` returns the NodeUVA of the cell's a-th vertex; the nodal
// state is consumed vertex by vertex (gradients and lumped sums accumulated on the fly) to
// keep the live register set small.
template <class Fetch>
VF_HD void tri_record_t(const double (&x)[3][2], double emod, const LameFac& lf, double eta,
                        double rho, const Damping& dp, const JacMix& mix, bool with_res,
                        Fetch fetch, double* rec) {
  CellGeo<2> g;
  p1_geometry(x, g);
  const CellCoef cf = cell_coef<2>(emod, lf, eta, rho, g.vol, dp);
  const double cv = mix.c, ca = mix.m;
  D2* r2 = reinterpret_cast<D2*>(rec);  // 16-byte stores
  for (int a = 0; a < 3; ++a) r2[a] = D2{g.G[a][0], g.G[a][1]};
  r2[3] = D2{mix.k * cf.lamv + cv * cf.vlam, mix.k * cf.muv + cv * cf.vmu};
  const double mass_blk = ca * cf.massv + cv * cf.vmass;
  if (!with_res) {
    r2[4] = D2{mass_blk, 0.0};
    return;
  }
  double gu[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, gv[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  double As[2] = {0.0, 0.0}, Aa[3][2];  // lumped sums: massv a + vmass v
  for (int a = 0; a < 3; ++a) {
    const NodeUVA s = fetch(a);
    const double w1c[2] = {s.u.x, s.u.y}, vc[2] = {s.v.x, s.v.y}, ac[2] = {s.a.x, s.a.y};
    for (int c = 0; c < 2; ++c) {
      gu[c][0] += w1c[c] * g.G[a][0];
      gu[c][1] += w1c[c] * g.G[a][1];
      gv[c][0] += vc[c] * g.G[a][0];
      gv[c][1] += vc[c] * g.G[a][1];
      const double lump = cf.massv * ac[c] + cf.vmass * vc[c];
      As[c] += lump;
      Aa[a][c] = lump;
    }
  }
  const double tr = gu[0][0] + gu[1][1], trv = gv[0][0] + gv[1][1];
  const double iso = cf.lamv * tr + cf.vlam * trv;
  const double s00 = cf.muv * (gu[0][0] + gu[0][0]) + cf.vmu * (gv[0][0] + gv[0][0]) + iso;
  const double s11 = cf.muv * (gu[1][1] + gu[1][1]) + cf.vmu * (gv[1][1] + gv[1][1]) + iso;
  const double s01 = cf.muv * (gu[0][1] + gu[1][0]) + cf.vmu * (gv[0][1] + gv[1][0]);
  double r[3][2];
  for (int a = 0; a < 3; ++a) {
    r[a][0] = s00 * g.G[a][0] + s01 * g.G[a][1] + (As[0] + Aa[a][0]);
    r[a][1] = s01 * g.G[a][0] + s11 * g.G[a][1] + (As[1] + Aa[a][1]);
  }
  r2[4] = D2{mass_blk, r[0][0]};
  r2[5] = D2{r[0][1], r[1][0]};
  r2[6] = D2{r[1][1], r[2][0]};
  r2[7] = D2{r[2][1], 0.0};
}

// Pointer version: every vertex is gathered from the global vectors.
VF_HD void tri_record(const double (&x)[3][2], const int (&nd)[3], double emod,
                      const LameFac& lf, double eta, double rho, const Damping& dp,
                      const NewmarkCoef& nc, bool is_static, bool with_res, const double* u1,
                      const double* u0, const double* v0, const double* a0, double* rec) {
  tri_record_t(x, emod, lf, eta, rho, dp, jac_mix_du1(nc, is_static), with_res,
               [&](int a) { return gather_node_uva(nc, is_static, nd[a], u1, u0, v0, a0); }, rec);
}

// Block (a, c) of the cell matrix from a record; identical arithmetic to cell_block<2>.
VF_HD void tri_block(const double* rec, int a, int c, double (&b)[2][2]) {
  const D2* r2 = reinterpret_cast<const D2*>(rec);
  const D2 ga = r2[a], gc = r2[c], lm = r2[3];
  const double lamv = lm.x, mv = lm.y;
  const double dg = mv * (ga.x * gc.x + ga.y * gc.y) + rec[8] * (a == c ? 2.0 : 1.0);
  b[0][0] = lamv * ga.x * gc.x + mv * gc.x * ga.x + dg;
  b[0][1] = lamv * ga.x * gc.y + mv * gc.x * ga.y;
  b[1][0] = lamv * ga.y * gc.x + mv * gc.y * ga.x;
  b[1][1] = lamv * ga.y * gc.y + mv * gc.y * ga.y + dg;
}

// Scalar row `comp` of the blocks (a, a), (a, next), (a, prev) with next = (a+1)%3,
// prev = (a+2)%3 (the cell's counter-clockwise order), from a record.
VF_HD void tri_row_fan(const double* rec, int a, int comp, D2& w_self, D2& w_next, D2& w_prev) {
  const D2* r2 = reinterpret_cast<const D2*>(rec);
  // (a+1)%3 and (a+2)%3 from nibble tables: the three gradients are fetched by address, not
  // loaded all and permuted in registers
  const int nx = (0x021 >> (4 * a)) & 3, pv = (0x102 >> (4 * a)) & 3;
  const D2 lm = r2[3];
  const double mass = rec[8];
  const D2 gc[3] = {r2[a], r2[nx], r2[pv]};
  const D2 ga = gc[0];
  const double A = lm.x * (comp == 0 ? ga.x : ga.y);
  const double Bx = lm.y * ga.x, By = lm.y * ga.y;
  // the diagonal term goes to entry `comp` of the row: e0 * dg is exact (e0 is 0 or 1), so the
  // multiply-add rounds once like the plain addition it replaces
  const double e0 = comp == 0 ? 1.0 : 0.0, e1 = 1.0 - e0;
  D2 w[3];
  for (int c = 0; c < 3; ++c) {
    const double gc_i = comp == 0 ? gc[c].x : gc[c].y;
    double w0 = A * gc[c].x + gc_i * Bx;
    double w1 = A * gc[c].y + gc_i * By;
    const double dg = gc[c].x * Bx + gc[c].y * By + mass * (c == 0 ? 2.0 : 1.0);
    w0 = e0 * dg + w0;
    w1 = e1 * dg + w1;
    w[c] = D2{w0, w1};
  }
  w_self = w[0];
  w_next = w[1];
  w_prev = w[2];
}

// Scalar row `comp` of the three blocks (a, c = 0..2) of the cell matrix, from a record.
// Per-pair factors are hoisted (A = lv ga_i, B = mv ga); each block then costs 7 fp64 ops.
VF_HD void tri_row_blocks(const double* rec, int a, int comp, D2 (&w)[3]) {
  const D2* r2 = reinterpret_cast<const D2*>(rec);
  const D2 ga = r2[a], lm = r2[3];
  const double mass = rec[8];
  const double A = lm.x * (comp == 0 ? ga.x : ga.y);
  const double Bx = lm.y * ga.x, By = lm.y * ga.y;
  for (int c = 0; c < 3; ++c) {
    const D2 gc = r2[c];
    const double gc_i = comp == 0 ? gc.x : gc.y;
    double w0 = A * gc.x + gc_i * Bx;
    double w1 = A * gc.y + gc_i * By;
    const double dg = gc.x * Bx + gc.y * By + mass * (a == c ? 2.0 : 1.0);
    if (comp == 0) w0 += dg;
    else w1 += dg;
    w[c] = D2{w0, w1};
  }
}

}  // namespace vf
