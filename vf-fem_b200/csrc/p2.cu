// P2 (6-node, straight-sided) triangle residual + Jacobian assembly: the "P2 extension" of
// BASELINE.json's north_star / configs[2] ("~1M P2 triangles").  The reference is P1 only
// (/root/reference/src/femvf/equations/form.py:521-524, 545-550), so this stands in for the same
// dfn.assemble calls (models/assemblyutils.py:49-50, models/transient.py:363-406) on a P2
// function space; the weak forms are those of elem.cuh (form.py:516-533, 540-572, 965-990,
// 733-756, 1067-1113).  Self-contained object (vf_p2_*) with caller-owned device vectors.
//
// Closed forms instead of quadrature: on an affine triangle grad phi_a = sum_k D_ak(L) G_k with
// the three constant P1 gradients G_k and D linear in the barycentric coordinates, so with the
// exact reference tensor W_abkl = (1/|K|) int D_ak D_bl (p2_tables.h)
//     T_ab = sum_kl W_abkl G_k (x) G_l
//     K_ab = |K| (lam T_ab + mu T_ab^T + mu tr(T_ab) I),   C_ab = |K| eta/2 (T_ab^T + tr(T_ab) I),
//     M_ab = rho |K| m_ab I
// and J_ab = K_ab + cv C_ab + ca M_ab.  The follower pressure on the P2 edges is integrated with
// three Gauss points (exact: degree 5).  The oracle (oracle/fem_p2.py) evaluates everything by
// quadrature instead: two independent routes to the same numbers.
//
// One thread per node gathers its adjacent cells and accumulates its block row in a private
// shared-memory row (no atomics: bit-reproducible), which is then streamed to the CSR array.
// HBM-bound by bytes (B_asm = 8 nnz + 40 N + ...); this first version is latency / L2 bound.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/vffem_b200.h"
#include "elem.cuh"
#include "p2_tables.h"
#include "p2_node.cuh"

namespace vf {
int fail(const std::string& msg);
}

struct vf_p2 {
  int nn, ne, nfp, max_deg;
  long long nnzb;
  char* mem;
  // device tables
  double* xy;        // (nn, 2)
  int* cells;        // (ne, 6)
  int* brptr;        // nn + 1
  int* bcol;         // nnzb
  int* n2e_ptr;      // nn + 1
  int* n2e;          // pairs: cell * 8 + local node
  unsigned* n2e_slots;  // per pair: CSR slots of the cell's 6 nodes in this node's row, 5 bits each
  int* n2f_ptr;      // nn + 1
  int* n2f;          // pressure edge * 4 + local position (0, 1: vertices, 2: mid-edge node)
  int* n2f_pair;     // per entry: index of the (node, parent cell) pair in n2e
  int* pf_cell;      // nfp
  int* pf_loc;       // (nfp, 3) local nodes of the edge in the parent cell (va, vb, mid)
  double* pf_geo;    // (nfp, 3): outward unit normal, length
  unsigned char* fixed;  // nn
  int* order;        // thread -> node: vertex nodes first, then mid-edge nodes (uniform warps)
  double* uva;       // (nn, 6) packed nodal (u1, v_nmk, a_nmk) of the current assembly (version 2)
  int n_class0;      // number of vertex nodes in `order`
  int max_deg0, max_deg1;  // longest block row of each class
};

namespace {

using namespace vf;

// vf_p2_assemble runs version 2 of the kernel by default (measured on B200, 1.0 M P2 triangles:
// 1.24 -> 1.02 ms, identical bits; profiles/r2_variants_ab.json).  VF_P2_WARP=0 selects version 1.
constexpr bool kP2WarpDefault = true;

__global__ void __launch_bounds__(64) p2_assemble_kernel(vf_p2 P, P2Args A, int first, int count,
                                                         int max_deg) {
  extern __shared__ double s_rows[];
  // the reference tensors are indexed by the thread's local node: shared memory, not the
  // (warp-uniform) constant cache
  __shared__ double sW[6][6][3][3], sM[6][6];
  for (int t = threadIdx.x; t < 324; t += blockDim.x) (&sW[0][0][0][0])[t] = (&kP2W[0][0][0][0])[t];
  for (int t = threadIdx.x; t < 36; t += blockDim.x) (&sM[0][0])[t] = (&kP2M[0][0])[t];
  __syncthreads();
  const int tix = blockIdx.x * blockDim.x + threadIdx.x;
  if (tix >= count) return;
  const int i = P.order[first + tix];
  const int stride = 4 * max_deg + 2;  // +2: 16-byte aligned rows on different banks
  double* row = s_rows + (size_t)threadIdx.x * stride;
  const int b0 = P.brptr[i], deg = P.brptr[i + 1] - b0;
  // row layout in shared memory: block s = (r00, r01, r10, r11)
  if (A.jac)
    for (int s = 0; s < 4 * deg; ++s) row[s] = 0.0;
  double r0 = 0.0, r1 = 0.0;
  const NewmarkCoef nc = newmark_coef(A.dt);
  const double cv = nc.cv, ca = nc.ca;
  const LameFac lf = lame_fac(A.nu);

  for (int t = P.n2e_ptr[i]; t < P.n2e_ptr[i + 1]; ++t) {
    const int ref = P.n2e[t];
    const int e = ref >> 3, a = ref & 7;
    const unsigned slots = P.n2e_slots[t];
    int nd[6];
#pragma unroll
    for (int b = 0; b < 6; ++b) nd[b] = P.cells[6 * e + b];
    double x[3][2];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      x[k][0] = P.xy[2 * nd[k]];
      x[k][1] = P.xy[2 * nd[k] + 1];
    }
    // P1 gradients and area
    const double e1x = x[1][0] - x[0][0], e1y = x[1][1] - x[0][1];
    const double e2x = x[2][0] - x[0][0], e2y = x[2][1] - x[0][1];
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
    for (int b = 0; b < 6; ++b) {
      // T = sum_kl W_abkl G_k (x) G_l
      double T00 = 0.0, T01 = 0.0, T10 = 0.0, T11 = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int l = 0; l < 3; ++l) {
          const double w = sW[a][b][k][l];
          T00 += w * G[k][0] * G[l][0];
          T01 += w * G[k][0] * G[l][1];
          T10 += w * G[k][1] * G[l][0];
          T11 += w * G[k][1] * G[l][1];
        }
      const double tr = T00 + T11;
      const double mab = rho * vol * sM[a][b];
      // elastic block K and the symmetric-gradient part S = T^T + tr I shared with C
      const double S00 = T00 + tr, S01 = T10, S10 = T01, S11 = T11 + tr;
      const double K00 = vol * (lam * T00 + mu * S00), K01 = vol * (lam * T01 + mu * S01);
      const double K10 = vol * (lam * T10 + mu * S10), K11 = vol * (lam * T11 + mu * S11);
      const double ch = 0.5 * eta * vol;
      if (A.res) {
        const int n = nd[b];
        const double u1x = A.u1[2 * n], u1y = A.u1[2 * n + 1];
        const double u0x = A.u0[2 * n], u0y = A.u0[2 * n + 1];
        const double v0x = A.v0[2 * n], v0y = A.v0[2 * n + 1];
        const double a0x = A.a0[2 * n], a0y = A.a0[2 * n + 1];
        const double vx = newmark_v(nc, u1x, u0x, v0x, a0x), vy = newmark_v(nc, u1y, u0y, v0y, a0y);
        const double ax = newmark_a(nc, u1x, u0x, v0x, a0x), ay = newmark_a(nc, u1y, u0y, v0y, a0y);
        r0 += K00 * u1x + K01 * u1y + ch * (S00 * vx + S01 * vy) + mab * ax;
        r1 += K10 * u1x + K11 * u1y + ch * (S10 * vx + S11 * vy) + mab * ay;
      }
      if (A.jac) {
        const int s = (slots >> (5 * b)) & 31;
        const double cc = cv * ch, mm = ca * mab;
        row[4 * s + 0] += K00 + cc * S00 + mm;
        row[4 * s + 1] += K01 + cc * S01;
        row[4 * s + 2] += K10 + cc * S10;
        row[4 * s + 3] += K11 + cc * S11 + mm;
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
      // cof(F) n with F = I + grad u (2D: linear in grad u)
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
          // d c0 / d U_b,y = g_b,y nx - g_b,x ny ;  d c1 / d U_b,x = -g_b,y nx + g_b,x ny
          const double dd = g[b][1] * nx - g[b][0] * ny;
          row[4 * s + 1] += w * dd;
          row[4 * s + 2] -= w * dd;
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
        row[4 * s + 0] = self ? 1.0 : 0.0;
        row[4 * s + 1] = 0.0;
        row[4 * s + 2] = 0.0;
        row[4 * s + 3] = self ? 1.0 : 0.0;
      }
    }
  }
  if (A.res) {
    A.F[2 * i] = r0;
    A.F[2 * i + 1] = r1;
  }
  if (A.jac) {
    // scalar CSR order of the block row: [row 0: deg x (c0, c1)] [row 1: deg x (c0, c1)]
    double* out = A.J + (size_t)4 * b0;
    for (int s = 0; s < deg; ++s) {
      out[2 * s] = row[4 * s];
      out[2 * s + 1] = row[4 * s + 1];
      out[2 * deg + 2 * s] = row[4 * s + 2];
      out[2 * deg + 2 * s + 1] = row[4 * s + 3];
    }
  }
}

// Pre-pass of version 2 (optional, VF_P2_PACK=1): packed nodal (u1, v_nmk, a_nmk).
__global__ void p2_pack_state_kernel(P2Args A, int nn, double* __restrict__ uva) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < nn) p2_pack_state(A, newmark_coef(A.dt), n, uva);
}

// Second version of p2_assemble_kernel: same node-owner scheme, same summation order per entry
// (so the two agree to the last bits), three changes for the memory system:
//   * the private row is kept in the layout of the CSR array ([row 0: deg x (c0, c1)][row 1: ...])
//     at an ODD stride in doubles (rows of the 32 lanes start on different banks), and a warp
//     writes its 32 rows out one after another with all lanes on consecutive doubles: coalesced
//     256-byte stores instead of 32 partial sectors per store instruction;
//   * dphi_a/dL_k is non-zero only for k = a (vertex node a) or the two vertices of the edge
//     (mid-edge node a), so W_abkl has 1 / 2 / 4 structural non-zeros per (a, b) instead of 9:
//     the class of the launch (CLS 0: vertex nodes, 1: mid-edge nodes) fixes the count and only
//     those terms are summed, in the same (k, l) order;
//   * nodal pairs (x, y) are fetched by 16-byte loads.
template <int CLS>
__global__ void __launch_bounds__(64) p2_assemble_warp_kernel(vf_p2 P, P2Args A, int first,
                                                              int count, int max_deg) {
  extern __shared__ double s_rows[];
  __shared__ double sW[6][6][3][3], sM[6][6];
  for (int t = threadIdx.x; t < 324; t += blockDim.x) (&sW[0][0][0][0])[t] = (&kP2W[0][0][0][0])[t];
  for (int t = threadIdx.x; t < 36; t += blockDim.x) (&sM[0][0])[t] = (&kP2M[0][0])[t];
  __syncthreads();
  const int tix = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = tix < count;
  const int lane = threadIdx.x & 31;
  const int stride = (4 * max_deg) | 1;
  double* row = s_rows + (size_t)threadIdx.x * stride;
  int b0 = 0, deg = 0;
  if (active) {
    const int i = P.order[first + tix];
    b0 = P.brptr[i];
    deg = P.brptr[i + 1] - b0;
    P2View V;
    V.xy = P.xy; V.cells = P.cells; V.brptr = P.brptr; V.bcol = P.bcol;
    V.n2e_ptr = P.n2e_ptr; V.n2e = P.n2e; V.n2e_slots = P.n2e_slots;
    V.n2f_ptr = P.n2f_ptr; V.n2f = P.n2f; V.n2f_pair = P.n2f_pair;
    V.pf_cell = P.pf_cell; V.pf_loc = P.pf_loc; V.pf_geo = P.pf_geo; V.fixed = P.fixed;
    double r0, r1;
    p2_node_row<CLS>(V, A, &sW[0][0][0][0], &sM[0][0], i, row, r0, r1);
    if (A.res) reinterpret_cast<P2Pair*>(A.F)[i] = P2Pair{r0, r1};
  }
  if (A.jac) {
    __syncwarp();
    const double* wrow = s_rows + (size_t)(threadIdx.x - lane) * stride;
    for (int src = 0; src < 32; ++src) {
      const int n4 = 4 * __shfl_sync(0xffffffffu, deg, src);
      const int bs = __shfl_sync(0xffffffffu, b0, src);
      double* __restrict__ out = A.J + (size_t)4 * bs;
      const double* rs = wrow + (size_t)src * stride;
      for (int j = lane; j < n4; j += 32) out[j] = rs[j];
    }
  }
}

size_t align256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace

extern "C" {

int vf_p2_create(int nn, int ne, const double* coords_host, const int32_t* cells6_host,
                 const int32_t* brptr_host, const int32_t* bcol_host, const int32_t* n2e_ptr_host,
                 const int32_t* n2e_host, const uint32_t* n2e_slots_host,
                 const int32_t* n2f_ptr_host, const int32_t* n2f_host,
                 const int32_t* n2f_pair_host, int nfp, const int32_t* pf_cell_host,
                 const int32_t* pf_loc_host, const double* pf_geo_host, const uint8_t* fixed_host,
                 const int32_t* order_host, int n_class0, void* stream, vf_p2** out) {
  if (!out) return vf::fail("vf_p2_create: null output");
  if (nn <= 0 || ne <= 0 || !coords_host || !cells6_host || !brptr_host || !bcol_host ||
      !n2e_ptr_host || !n2e_host || !n2e_slots_host || !n2f_ptr_host || !fixed_host ||
      !order_host || n_class0 < 0 || n_class0 > nn)
    return vf::fail("vf_p2_create: missing tables");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return vf::fail("vf_p2_create: no CUDA device (there is no CPU fallback)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  vf_p2* P = new vf_p2{};
  P->nn = nn;
  P->ne = ne;
  P->nfp = nfp;
  P->nnzb = brptr_host[nn];
  const int npair = n2e_ptr_host[nn], nnf = n2f_ptr_host[nn];
  int max_deg = 0;
  for (int i = 0; i < nn; ++i) max_deg = std::max(max_deg, brptr_host[i + 1] - brptr_host[i]);
  if (max_deg > 31) {
    delete P;
    return vf::fail("vf_p2_create: a node couples to more than 31 nodes");
  }
  P->max_deg = max_deg;
  P->n_class0 = n_class0;
  P->max_deg0 = P->max_deg1 = 1;
  for (int t = 0; t < nn; ++t) {
    const int i = order_host[t];
    if (i < 0 || i >= nn) {
      delete P;
      return vf::fail("vf_p2_create: order is not a permutation of the nodes");
    }
    int& md = t < n_class0 ? P->max_deg0 : P->max_deg1;
    md = std::max(md, brptr_host[i + 1] - brptr_host[i]);
  }
  struct Item {
    void** dst;
    const void* src;
    size_t bytes;
  };
  std::vector<Item> items = {
      {(void**)&P->xy, coords_host, sizeof(double) * 2 * (size_t)nn},
      {(void**)&P->cells, cells6_host, sizeof(int) * 6 * (size_t)ne},
      {(void**)&P->brptr, brptr_host, sizeof(int) * ((size_t)nn + 1)},
      {(void**)&P->bcol, bcol_host, sizeof(int) * (size_t)P->nnzb},
      {(void**)&P->n2e_ptr, n2e_ptr_host, sizeof(int) * ((size_t)nn + 1)},
      {(void**)&P->n2e, n2e_host, sizeof(int) * (size_t)npair},
      {(void**)&P->n2e_slots, n2e_slots_host, sizeof(unsigned) * (size_t)npair},
      {(void**)&P->n2f_ptr, n2f_ptr_host, sizeof(int) * ((size_t)nn + 1)},
      {(void**)&P->n2f, n2f_host, sizeof(int) * (size_t)std::max(nnf, 1)},
      {(void**)&P->n2f_pair, n2f_pair_host, sizeof(int) * (size_t)std::max(nnf, 1)},
      {(void**)&P->pf_cell, pf_cell_host, sizeof(int) * (size_t)std::max(nfp, 1)},
      {(void**)&P->pf_loc, pf_loc_host, sizeof(int) * 3 * (size_t)std::max(nfp, 1)},
      {(void**)&P->pf_geo, pf_geo_host, sizeof(double) * 3 * (size_t)std::max(nfp, 1)},
      {(void**)&P->fixed, fixed_host, (size_t)nn},
      {(void**)&P->order, order_host, sizeof(int) * (size_t)nn},
      {(void**)&P->uva, nullptr, sizeof(double) * 6 * (size_t)nn},  // work space, not uploaded
  };
  size_t total = 0;
  for (auto& it : items) total += align256(it.bytes);
  if (cudaMalloc(&P->mem, total) != cudaSuccess) {
    delete P;
    return vf::fail("vf_p2_create: device allocation failed");
  }
  size_t off = 0;
  for (auto& it : items) {
    *it.dst = P->mem + off;
    const bool have = it.src != nullptr && !((it.dst == (void**)&P->n2f || it.dst == (void**)&P->n2f_pair) && nnf == 0) &&
                      !((it.dst == (void**)&P->pf_cell || it.dst == (void**)&P->pf_loc || it.dst == (void**)&P->pf_geo) && nfp == 0);
    if (have) cudaMemcpyAsync(P->mem + off, it.src, it.bytes, cudaMemcpyHostToDevice, st);
    off += align256(it.bytes);
  }
  if (cudaStreamSynchronize(st) != cudaSuccess) {
    cudaFree(P->mem);
    delete P;
    return vf::fail("vf_p2_create: upload failed");
  }
  *out = P;
  return 0;
}

void vf_p2_destroy(vf_p2* P) {
  if (!P) return;
  cudaFree(P->mem);
  delete P;
}

long long vf_p2_nnz(const vf_p2* P) { return P ? 4 * P->nnzb : 0; }

int vf_p2_assemble(vf_p2* P, int flags, double dt, double nu, const double* emod_dev,
                   const double* eta_dev, const double* rho_dev, const double* u1_dev,
                   const double* u0_dev, const double* v0_dev, const double* a0_dev,
                   const double* p1_dev, double* F_dev, double* J_dev, void* stream) {
  if (!P) return vf::fail("null P2 assembler");
  P2Args A;
  A.res = flags & 1;
  A.jac = (flags & 2) != 0;
  if (!A.res && !A.jac) return 0;
  if (!emod_dev || !eta_dev || !rho_dev || !u1_dev || !u0_dev || !v0_dev || !a0_dev || !p1_dev)
    return vf::fail("vf_p2_assemble: null input");
  if ((A.res && !F_dev) || (A.jac && !J_dev)) return vf::fail("vf_p2_assemble: null output");
  if (!(dt > 0.0)) return vf::fail("dt must be positive");
  A.emod = emod_dev; A.eta = eta_dev; A.rho = rho_dev;
  A.u1 = u1_dev; A.u0 = u0_dev; A.v0 = v0_dev; A.a0 = a0_dev; A.p1 = p1_dev;
  A.F = F_dev; A.J = J_dev; A.nu = nu; A.dt = dt;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(p2_assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(p2_assemble_warp_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(p2_assemble_warp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    attr_set = true;
  }
  // version 2 (coalesced row write-out, structural zeros of W skipped): VF_P2_WARP=0/1 overrides
  // the default; it needs 16-byte aligned nodal vectors (double2 loads), else version 1 runs
  const char* env_warp = getenv("VF_P2_WARP");  // read per call: tests switch it
  const bool want_warp = env_warp ? atoi(env_warp) != 0 : kP2WarpDefault;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool warp_ok = want_warp && al16(u1_dev) && al16(u0_dev) && al16(v0_dev) && al16(a0_dev) &&
                       (!A.res || al16(F_dev));
  // one launch per node class (vertex nodes: ~19 blocks per row, 6 cells; mid-edge nodes: 9 blocks,
  // 2 cells): uniform trip counts inside a warp, shared-memory rows sized per class
  const int block = 64;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (warp_ok && A.res) {
    // optional pre-pass (VF_P2_PACK=1): (u1, v_nmk, a_nmk) per node, packed for the residual
    // gathers of version 2.  Measured neutral (1.020 ms with and without: the kernel is not bound
    // by those gathers), so it is off by default
    const char* env_pack = getenv("VF_P2_PACK");
    if (env_pack && atoi(env_pack) != 0) {
      p2_pack_state_kernel<<<(P->nn + 255) / 256, 256, 0, st>>>(A, P->nn, P->uva);
      A.uva = P->uva;
    }
  }
  const int first[2] = {0, P->n_class0}, count[2] = {P->n_class0, P->nn - P->n_class0};
  const int mdeg[2] = {P->max_deg0, P->max_deg1};
  for (int c = 0; c < 2; ++c) {
    if (count[c] <= 0) continue;
    const int grid = (count[c] + block - 1) / block;
    if (warp_ok) {
      const size_t smem = sizeof(double) * (size_t)block * ((4 * mdeg[c]) | 1);
      if (c == 0) p2_assemble_warp_kernel<0><<<grid, block, smem, st>>>(*P, A, first[c], count[c], mdeg[c]);
      else p2_assemble_warp_kernel<1><<<grid, block, smem, st>>>(*P, A, first[c], count[c], mdeg[c]);
      continue;
    }
    const size_t smem = sizeof(double) * (size_t)block * (4 * mdeg[c] + 2);
    p2_assemble_kernel<<<grid, block, smem, st>>>(*P, A, first[c], count[c], mdeg[c]);
  }
  if (cudaGetLastError() != cudaSuccess) return vf::fail("vf_p2_assemble: launch failed");
  return 0;
}

}  // extern "C"
