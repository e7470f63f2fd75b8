// Device-resident Newton / GMRES / Newmark / FSI step executed by ONE thread block per
// ensemble member (single simulations are an ensemble of one).
//
// Stands in for, per time step of forward.integrate_steps
// (/root/reference/src/femvf/forward.py:169-184):
//   ExplicitFSIModel.solve_state1             models/transient.py:899-920
//   _set_ini_fluid_state / _set_fin_solid_state    :833-862, models/fsi.py:66-70
//   FenicsModel.solve_state1 + nonlineq.newton_solve   :441-468, solverconst.py:1-6
//   FenicsModel.solve_dres_dstate1 (PETSc LU)          :470-491
//   JaxModel.solve_state1                              :667-672
// The whole time loop runs inside one kernel launch: no host round trip per Newton
// iteration, per Krylov iteration or per time step.  The base problem is O(10^2..10^3)
// DOFs, i.e. latency bound; one CTA keeps the member's vectors hot in
// L1/L2 and the 148 SMs run 148+ members concurrently.
//
// Linear solver: left-preconditioned restarted GMRES(m) with block-Jacobi (d x d nodal
// blocks) and classical Gram-Schmidt with re-orthogonalisation (CGS2); all reductions are
// fixed-order warp-shuffle trees, so results are bit-reproducible for a given CTA size.
#pragma once

#include "engine_types.cuh"

namespace vf {

// ---- block-level primitives ------------------------------------------------------
struct BlockShared {
  double red[40];
  double h[kMaxRestart + 2];
  double h2[kMaxRestart + 2];
  double cs[kMaxRestart + 2];  // Givens rotations, rhs of the least-squares problem and its
  double sn[kMaxRestart + 2];  // solution: touched by one thread in a dependent chain, so
  double g[kMaxRestart + 2];   // they must not live in global memory
  double y[kMaxRestart + 2];
  double bc[8];  // broadcast scalars
  double B[kDenseNb][kDenseNb + 1];  // pivot block of the blocked Gauss-Jordan
  long long cyc[8];  // cycle counters: 0 assembly 1 spmv 2 dots+update 3 givens 4 solve tail 5 fluid 6 total
};

__device__ __forceinline__ double block_sum(double v, BlockShared& sh) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) sh.red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double t = (lane < nw) ? sh.red[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) sh.red[32] = t;
  }
  __syncthreads();
  return sh.red[32];
}

template <int D>
__device__ __forceinline__ void blk_spmv(const EngineDev& E, const double* __restrict__ J,
                                         const double* __restrict__ x, double* __restrict__ y) {
  for (int r = threadIdx.x; r < E.N; r += blockDim.x) {
    const int i = r / D, a = r - i * D;
    const int b0 = E.mesh.brptr[i], deg = E.mesh.brptr[i + 1] - b0;
    const double* row = J + (size_t)D * D * b0 + (size_t)a * D * deg;
    double s = 0.0;
    for (int k = 0; k < deg; ++k) {
      const int j = E.mesh.bcol[b0 + k];
#pragma unroll
      for (int c = 0; c < D; ++c) s += row[k * D + c] * x[D * j + c];
    }
    y[r] = s;
  }
}

// Solver working set of one member.  In the time loop the arrays live in shared memory when
// they fit (flags: 1 = vectors + Dinv, 2 = Krylov basis V, 4 = CSR values J); otherwise they
// are the member's global arrays (L2 resident).  The base problem (~300
// DOF) fits entirely, so a Krylov iteration never leaves the SM.
struct SolverWork {
  double *J, *F, *dx, *Dinv, *V, *w, *z, *t, *H, *cs, *sn, *g, *y;
  double *P, *pstate;  // dense fp64 work matrix (global) and {dt it was built for, refresh flag}
  float* Pf;           // transposed fp32 inverse
  double* Pscr;        // (warps x N) partial sums of the dense mat-vec
};

__device__ __forceinline__ SolverWork make_work(const EngineDev& E, double* mb, double* dsm,
                                                int flags) {
  const Layout& L = E.L;
  SolverWork W;
  W.J = mb + L.off[VF_J];
  W.F = mb + L.off[VF_F];
  W.dx = mb + L.off[VF_DX];
  W.Dinv = mb + L.Dinv;
  W.V = mb + L.V;
  W.w = mb + L.w;
  W.z = mb + L.z;
  W.t = mb + L.xk;
  W.H = mb + L.H;
  W.cs = mb + L.cs;
  W.sn = mb + L.sn;
  W.g = mb + L.g;
  W.y = mb + L.y;
  W.P = E.dense ? mb + L.Pinv : nullptr;
  W.Pf = E.dense ? reinterpret_cast<float*>(mb + L.Pf) : nullptr;
  W.Pscr = E.dense ? mb + L.Pscr : nullptr;
  W.pstate = E.dense ? mb + L.pstate : nullptr;
  size_t o = 0;
  auto take = [&](size_t n) {
    double* p = dsm + o;
    o += (n + 1) & ~size_t(1);
    return p;
  };
  const size_t N = E.N;
  if (flags & 1) {
    W.F = take(N);
    W.dx = take(N);
    W.w = take(N);
    W.z = take(N);
    W.t = take(N);
    W.Dinv = take((size_t)E.mesh.nn * E.d * E.d);
  }
  if (flags & 8) W.H = take((size_t)(E.restart + 1) * E.restart);
  if (flags & 2) W.V = take((size_t)(E.restart + 1) * N);
  if (flags & 4) W.J = take((size_t)E.nnz);
  return W;
}

// y = Dinv (J x): the left-preconditioned operator in one sweep, one thread per node
template <int D>
__device__ __forceinline__ void blk_spmv_prec(const EngineDev& E, const double* __restrict__ J,
                                              const double* __restrict__ Dinv,
                                              const double* __restrict__ x, double* __restrict__ y,
                                              const double* __restrict__ tcomb = nullptr) {
  for (int i = threadIdx.x; i < E.mesh.nn; i += blockDim.x) {
    const int b0 = E.mesh.brptr[i], deg = E.mesh.brptr[i + 1] - b0;
    const double* blk = J + (size_t)D * D * b0;
    double acc[D];
#pragma unroll
    for (int a = 0; a < D; ++a) acc[a] = 0.0;
    for (int k = 0; k < deg; ++k) {
      const int j = E.mesh.bcol[b0 + k];
      double xv[D];
#pragma unroll
      for (int c = 0; c < D; ++c) xv[c] = x[D * j + c];
#pragma unroll
      for (int a = 0; a < D; ++a)
#pragma unroll
        for (int c = 0; c < D; ++c) acc[a] += blk[a * D * deg + k * D + c] * xv[c];
    }
    const double* o = Dinv + (size_t)D * D * i;
#pragma unroll
    for (int a = 0; a < D; ++a) {
      double t = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) t += o[a * D + c] * acc[c];
      // tcomb: the Neumann recurrence y = tcomb + x - D^{-1} J x in the same sweep
      y[D * i + a] = tcomb ? tcomb[D * i + a] + x[D * i + a] - t : t;
    }
  }
}

// inverse of the d x d diagonal blocks (block-Jacobi preconditioner)
template <int D>
__device__ __forceinline__ void blk_compute_dinv(const EngineDev& E, const double* J, double* Dinv) {
  for (int i = threadIdx.x; i < E.mesh.nn; i += blockDim.x) {
    const int b0 = E.mesh.brptr[i], deg = E.mesh.brptr[i + 1] - b0;
    const int self = find_slot(E.mesh.bcol + b0, deg, i);
    const double* blk = J + (size_t)D * D * b0;
    double A[D][D];
    for (int a = 0; a < D; ++a)
      for (int c = 0; c < D; ++c) A[a][c] = blk[a * D * deg + self * D + c];
    double* o = Dinv + (size_t)D * D * i;
    if constexpr (D == 2) {
      const double det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
      const double inv = 1.0 / det;
      o[0] = A[1][1] * inv;
      o[1] = -A[0][1] * inv;
      o[2] = -A[1][0] * inv;
      o[3] = A[0][0] * inv;
    } else {
      // A^{-1} = [r1 x r2, r2 x r0, r0 x r1] / det  (columns), r_k the rows of A
      double c0[3], c1[3], c2[3];
      cross3(A[1], A[2], c0);
      cross3(A[2], A[0], c1);
      cross3(A[0], A[1], c2);
      const double det = A[0][0] * c0[0] + A[0][1] * c0[1] + A[0][2] * c0[2];
      const double inv = 1.0 / det;
      for (int k = 0; k < 3; ++k) {
        o[k * 3 + 0] = c0[k] * inv;
        o[k * 3 + 1] = c1[k] * inv;
        o[k * 3 + 2] = c2[k] * inv;
      }
    }
  }
}

template <int D>
__device__ __forceinline__ void blk_apply_dinv(const EngineDev& E, const double* Dinv,
                                               const double* v, double* z) {
  for (int r = threadIdx.x; r < E.N; r += blockDim.x) {
    const int i = r / D, a = r - i * D;
    const double* o = Dinv + (size_t)D * D * i + a * D;
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) s += o[c] * v[D * i + c];
    z[r] = s;
  }
}

// out[j] = V_j . w for j < nvec.  Each warp takes four basis vectors at a time and keeps four
// independent accumulators / shuffle trees in flight (the cost here is dependent-chain
// latency, not throughput); fixed order, so bit-reproducible.
__device__ __forceinline__ void blk_dots(const double* V, int N, int nvec, const double* w,
                                         double* out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j0 = wid * 4; j0 < nvec; j0 += nw * 4) {
    const int nj = min(4, nvec - j0);
    const double* v0 = V + (size_t)j0 * N;
    const double* v1 = v0 + (nj > 1 ? N : 0);
    const double* v2 = v0 + (nj > 2 ? 2 * (size_t)N : 0);
    const double* v3 = v0 + (nj > 3 ? 3 * (size_t)N : 0);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (int t = lane; t < N; t += 32) {
      const double wt = w[t];
      s0 += v0[t] * wt;
      s1 += v1[t] * wt;
      s2 += v2[t] * wt;
      s3 += v3[t] * wt;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, off);
      s1 += __shfl_xor_sync(0xffffffffu, s1, off);
      s2 += __shfl_xor_sync(0xffffffffu, s2, off);
      s3 += __shfl_xor_sync(0xffffffffu, s3, off);
    }
    if (lane == 0) {
      out[j0] = s0;
      if (nj > 1) out[j0 + 1] = s1;
      if (nj > 2) out[j0 + 2] = s2;
      if (nj > 3) out[j0 + 3] = s3;
    }
  }
}

// w[t] -= sum_{j<nvec} h[j] V_j[t] with four independent partial sums; returns the thread's
// partial of ||w||^2
__device__ __forceinline__ double blk_project_out(const double* V, int N, int nvec,
                                                  const double* h, double* w) {
  double p2 = 0.0;
  for (int t = threadIdx.x; t < N; t += blockDim.x) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int j = 0;
    for (; j + 3 < nvec; j += 4) {
      a0 += h[j] * V[(size_t)j * N + t];
      a1 += h[j + 1] * V[(size_t)(j + 1) * N + t];
      a2 += h[j + 2] * V[(size_t)(j + 2) * N + t];
      a3 += h[j + 3] * V[(size_t)(j + 3) * N + t];
    }
    for (; j < nvec; ++j) a0 += h[j] * V[(size_t)j * N + t];
    const double s = w[t] - ((a0 + a1) + (a2 + a3));
    w[t] = s;
    p2 += s * s;
  }
  return p2;
}

// ---- dense inverse preconditioner ---------------------------------------------------
// For the small systems this kernel is built for (one CTA per member, N of a few hundred) the
// Newton matrix barely changes from step to step: K, C, M are constant and only the follower
// pressure / contact blocks move.  Its inverse is therefore formed ONCE per launch and reused
// as the left preconditioner of GMRES, which then converges in a few iterations instead of
// ten with the polynomial preconditioner.  This is the sparse-LU stand-in for PETSc's direct
// solve at these sizes (transient.py:487).
//   * inversion: in-place BLOCKED Gauss-Jordan in fp64 on a dense work matrix in global
//     memory (L2): pivots are eliminated kDenseNb at a time, the pivot rows and columns are
//     staged in shared memory, and every other entry is read and written once per block with
//     kDenseNb fused multiply-adds in between (N^3 flops, N^3 / kDenseNb * 16 bytes of L2
//     traffic).  No pivoting: the matrix is a Dirichlet-row-modified SPD matrix plus a small
//     perturbation.
//   * application: the inverse is kept TRANSPOSED in fp32 (it is only a preconditioner; GMRES
//     still converges the fp64 system to its tolerance), warps split the summation range,
//     lanes own output entries (coalesced 4-byte loads, several independent loads in flight),
//     partial sums meet in a small global scratch.

// A := inverse of the matrix held in A (row-major N x N, global).  rows/cols: shared (or
// global) panels of kDenseNb * N doubles each.
__device__ void blk_block_gauss_jordan(double* A, int N, double* rows, double* cols,
                                       double (&B)[kDenseNb][kDenseNb + 1]) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int k0 = 0; k0 < N; k0 += kDenseNb) {
    const int nb = min(kDenseNb, N - k0);
    // 1. stage the pivot rows and the (old) pivot columns
    for (int t = threadIdx.x; t < N; t += blockDim.x)
      for (int p = 0; p < nb; ++p) {
        rows[p * N + t] = A[(size_t)(k0 + p) * N + t];
        cols[p * N + t] = A[(size_t)t * N + k0 + p];
      }
    __syncthreads();
    // 2. B = A11^{-1} (nb x nb), unblocked Gauss-Jordan by one thread on a tiny matrix
    if (threadIdx.x == 0) {
      for (int p = 0; p < nb; ++p)
        for (int q = 0; q < nb; ++q) B[p][q] = rows[p * N + k0 + q];
      for (int k = 0; k < nb; ++k) {
        const double piv = 1.0 / B[k][k];
        for (int q = 0; q < nb; ++q) B[k][q] = (q == k) ? piv : B[k][q] * piv;
        for (int p = 0; p < nb; ++p) {
          if (p == k) continue;
          const double f = B[p][k];
          for (int q = 0; q < nb; ++q) B[p][q] = (q == k) ? -f * piv : B[p][q] - f * B[k][q];
        }
      }
    }
    __syncthreads();
    // 3. new pivot rows A12' = B A12 (columns outside the block), A11' = B; thread per column
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
      double r[kDenseNb];
      for (int p = 0; p < nb; ++p) r[p] = rows[p * N + j];
      const bool inblk = j >= k0 && j < k0 + nb;
      for (int p = 0; p < nb; ++p) {
        double v;
        if (inblk) {
          v = B[p][j - k0];
        } else {
          v = 0.0;
          for (int q = 0; q < nb; ++q) v += B[p][q] * r[q];
        }
        rows[p * N + j] = v;
        A[(size_t)(k0 + p) * N + j] = v;
      }
    }
    // 4. new pivot columns A21' = -A21 B (rows outside the block); thread per row.  The OLD
    //    columns stay in shared memory for the rank-nb update
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      if (i >= k0 && i < k0 + nb) continue;
      double c[kDenseNb];
      for (int p = 0; p < nb; ++p) c[p] = cols[p * N + i];
      for (int q = 0; q < nb; ++q) {
        double v = 0.0;
        for (int p = 0; p < nb; ++p) v -= c[p] * B[p][q];
        A[(size_t)i * N + k0 + q] = v;
      }
    }
    __syncthreads();
    // 5. A22 -= A21 A12': warps take groups of 4 rows, lanes take columns; the rows' column
    //    entries sit in registers, the new pivot rows come from shared memory, and the 4 loads
    //    of a column chunk are independent (latency of the L2 round trip is overlapped)
    for (int i0 = 4 * wid; i0 < N; i0 += 4 * nw) {
      double c[4][kDenseNb];
      bool live[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u;
        live[u] = i < N && !(i >= k0 && i < k0 + nb);
#pragma unroll
        for (int p = 0; p < kDenseNb; ++p) c[u][p] = (live[u] && p < nb) ? cols[p * N + i] : 0.0;
      }
      for (int j = lane; j < N; j += 32) {
        if (j >= k0 && j < k0 + nb) continue;
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = live[u] ? A[(size_t)(i0 + u) * N + j] : 0.0;
#pragma unroll
        for (int p = 0; p < kDenseNb; ++p) {
          const double rp = p < nb ? rows[p * N + j] : 0.0;
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] -= c[u][p] * rp;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (live[u]) A[(size_t)(i0 + u) * N + j] = v[u];
      }
    }
    __syncthreads();
  }
}

// Pf := float((J^{-1})^T), N x N row-major.  A: fp64 work matrix (N x N, global).
template <int D>
__device__ void blk_dense_inverse(const EngineDev& E, const double* __restrict__ J, double* A,
                                  float* Pf, double* rows, double* cols,
                                  double (&B)[kDenseNb][kDenseNb + 1]) {
  const int N = E.N;
  const size_t NN = (size_t)N * N;
  for (size_t t = threadIdx.x; t < NN; t += blockDim.x) A[t] = 0.0;
  __syncthreads();
  // scatter the CSR values transposed: A[col][row] = J[row][col], so that A^{-1} = (J^{-1})^T
  for (int r = threadIdx.x; r < N; r += blockDim.x) {
    const int i = r / D, a = r - i * D;
    const int b0 = E.mesh.brptr[i], deg = E.mesh.brptr[i + 1] - b0;
    const double* row = J + (size_t)D * D * b0 + (size_t)a * D * deg;
    for (int k = 0; k < deg; ++k) {
      const int j = E.mesh.bcol[b0 + k];
#pragma unroll
      for (int c = 0; c < D; ++c) A[(size_t)(D * j + c) * N + r] = row[k * D + c];
    }
  }
  __syncthreads();
  blk_block_gauss_jordan(A, N, rows, cols, B);
  const int ldp = dense_ldp(N);
  for (int i = threadIdx.x >> 5; i < N; i += blockDim.x >> 5)
    for (int j = threadIdx.x & 31; j < ldp; j += 32)
      Pf[(size_t)i * ldp + j] = j < N ? (float)A[(size_t)i * N + j] : 0.0f;
  __syncthreads();
}

// out = J^{-1} r with the transposed fp32 inverse: out[t] = sum_j Pf[j][t] r[j].
// One SM reaches its L2 bandwidth only with tens of KB in flight: every lane owns 4
// consecutive outputs per 128-entry chunk (16-byte loads), two chunks and four rows are
// unrolled, i.e. 8 independent 512-byte warp loads per iteration.
// scratch: (warps x ldp) doubles in global memory.
__device__ __forceinline__ void blk_dense_mv(const float* __restrict__ Pf, int N,
                                             const double* __restrict__ r, double* scratch,
                                             double* __restrict__ out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int ldp = dense_ldp(N);
  const int per = (N + nw - 1) / nw;
  const int j0 = wid * per, j1 = min(N, j0 + per);
  for (int c0 = 0; c0 < ldp; c0 += 256) {        // two chunks of 128 outputs per pass
    const int ta = c0 + 4 * lane, tb = c0 + 128 + 4 * lane;
    const bool ha = ta < ldp, hb = tb < ldp;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
    int j = j0;
    for (; j + 3 < j1; j += 4) {
      float4 xa[4], xb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* row = Pf + (size_t)(j + u) * ldp;
        xa[u] = ha ? *reinterpret_cast<const float4*>(row + ta) : make_float4(0, 0, 0, 0);
        xb[u] = hb ? *reinterpret_cast<const float4*>(row + tb) : make_float4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double rj = r[j + u];
        a0 += (double)xa[u].x * rj; a1 += (double)xa[u].y * rj;
        a2 += (double)xa[u].z * rj; a3 += (double)xa[u].w * rj;
        b0 += (double)xb[u].x * rj; b1 += (double)xb[u].y * rj;
        b2 += (double)xb[u].z * rj; b3 += (double)xb[u].w * rj;
      }
    }
    for (; j < j1; ++j) {
      const float* row = Pf + (size_t)j * ldp;
      const double rj = r[j];
      if (ha) {
        const float4 x = *reinterpret_cast<const float4*>(row + ta);
        a0 += (double)x.x * rj; a1 += (double)x.y * rj; a2 += (double)x.z * rj; a3 += (double)x.w * rj;
      }
      if (hb) {
        const float4 x = *reinterpret_cast<const float4*>(row + tb);
        b0 += (double)x.x * rj; b1 += (double)x.y * rj; b2 += (double)x.z * rj; b3 += (double)x.w * rj;
      }
    }
    double* sc = scratch + (size_t)wid * ldp;
    if (ha) { sc[ta] = a0; sc[ta + 1] = a1; sc[ta + 2] = a2; sc[ta + 3] = a3; }
    if (hb) { sc[tb] = b0; sc[tb + 1] = b1; sc[tb + 2] = b2; sc[tb + 3] = b3; }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < N; t += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += scratch[(size_t)w * ldp + t];
    out[t] = s;
  }
}

// w = M^{-1} J v with M^{-1} = (I + N + ... + N^p) D^{-1}, N = I - D^{-1} J: the block-Jacobi
// preconditioner accelerated by a truncated Neumann series (M^{-1} J = I - N^{p+1}).  For the
// mass-dominated Newmark Jacobian the spectral radius of N is ~0.6, so p = 3 shrinks the
// Krylov space from ~25 to ~9 vectors and with it the O(k^2) orthogonalisation work, which
// dominates an in-CTA GMRES.  t, z: scratch vectors.
template <int D>
__device__ __forceinline__ void blk_apply_op(const EngineDev& E, const SolverWork& W, int p,
                                             const double* v, double* w) {
  if (p < 0) {
    blk_spmv<D>(E, W.J, v, W.t);
    __syncthreads();
    blk_dense_mv(W.Pf, E.N, W.t, W.Pscr, w);
    __syncthreads();
    return;
  }
  if (p == 0) {
    blk_spmv_prec<D>(E, W.J, W.Dinv, v, w);
    __syncthreads();
    return;
  }
  double* t = W.t;
  blk_spmv_prec<D>(E, W.J, W.Dinv, v, t);  // t = D^{-1} J v
  __syncthreads();
  const double* cur = t;
  for (int q = 0; q < p; ++q) {
    // ping-pong between z and w so that the last term lands in w and dst never aliases cur
    double* dst = ((p - q) & 1) ? w : W.z;
    blk_spmv_prec<D>(E, W.J, W.Dinv, cur, dst, t);  // dst = t + cur - D^{-1} J cur
    __syncthreads();
    cur = dst;
  }
}

// M^{-1} r for a residual r (used for the initial / restart residual)
template <int D>
__device__ __forceinline__ void blk_apply_prec(const EngineDev& E, const SolverWork& W, int p,
                                               const double* r, double* out) {
  if (p < 0) {
    blk_dense_mv(W.Pf, E.N, r, W.Pscr, out);
    __syncthreads();
    return;
  }
  blk_apply_dinv<D>(E, W.Dinv, r, out);  // z0 = D^{-1} r
  __syncthreads();
  if (p == 0) return;
  double* z0 = W.t;
  for (int i = threadIdx.x; i < E.N; i += blockDim.x) z0[i] = out[i];
  __syncthreads();
  for (int q = 0; q < p; ++q) {
    // out <- z0 + out - D^{-1} J out   (needs a scratch for the product: W.z)
    blk_spmv_prec<D>(E, W.J, W.Dinv, out, W.z);
    __syncthreads();
    for (int i = threadIdx.x; i < E.N; i += blockDim.x) out[i] = z0[i] + out[i] - W.z[i];
    __syncthreads();
  }
}

// GMRES(m) on J x = b for one member; x starts at 0.  Left-preconditioned with the
// block-Jacobi inverse M^{-1}: the Krylov space is built for M^{-1} J, whose rows are
// equilibrated (Dirichlet identity rows and mass-dominated rows both have unit diagonal
// blocks), and convergence is tested on ||M^{-1}(b - J x)|| / ||M^{-1} b||.  Returns the
// iteration count; *resid_out is the final preconditioned residual norm, *bnorm_out
// ||M^{-1} b||.  Every thread of the block must call it.
template <int D>
__device__ int blk_gmres(const EngineDev& E, const SolverWork& W, const double* b, double* x,
                         const SolverOpts& opt, BlockShared& sh, double* resid_out,
                         double* bnorm_out, bool dense = false) {
  const int N = E.N;
  const int m = E.restart;
  const double* J = W.J;
  double* V = W.V;
  double* w = W.w;
  double* H = W.H;  // column-major, leading dimension m+1
  double* cs = sh.cs;
  double* sn = sh.sn;
  double* g = sh.g;
  double* y = sh.y;
  const int ldh = m + 1;

  // r0 = M^{-1} b  (x0 = 0)
  // static problems have no mass term: the spectral radius of N approaches (or exceeds) 1 and
  // the Neumann acceleration does not pay, so it is only used for the transient Jacobian
  int pdeg = dense ? -1 : (opt.is_static ? 0 : opt.poly_degree);
  for (int t = threadIdx.x; t < N; t += blockDim.x) x[t] = 0.0;
  __syncthreads();
  blk_apply_prec<D>(E, W, pdeg, b, w);
  double part = 0.0;
  for (int t = threadIdx.x; t < N; t += blockDim.x) part += w[t] * w[t];
  const double bnorm = sqrt(block_sum(part, sh));
  *bnorm_out = bnorm;
  if (bnorm == 0.0) {
    *resid_out = 0.0;
    return 0;
  }
  double tol = fmax(opt.gmres_rel_tol * bnorm, opt.gmres_abs_tol);
  double beta = bnorm;
  double resid = bnorm;
  int iters = 0;
  bool first = true;
  while (true) {
    if (!first) {
      if (pdeg != 0) {
        // a full cycle did not converge: the Neumann series is not contracting for this
        // matrix (or the dense inverse is stale beyond repair) -- fall back to plain
        // block-Jacobi (norms are re-based on the new M)
        pdeg = 0;
        __syncthreads();
        blk_apply_prec<D>(E, W, 0, b, w);
        double pb = 0.0;
        for (int t = threadIdx.x; t < N; t += blockDim.x) pb += w[t] * w[t];
        const double bn = sqrt(block_sum(pb, sh));
        *bnorm_out = bn;
        tol = fmax(opt.gmres_rel_tol * bn, opt.gmres_abs_tol);
      }
      // explicit restart residual r = M^{-1} (b - J x)
      __syncthreads();
      // V_0 is free during a restart: use it for the unpreconditioned residual
      blk_spmv<D>(E, J, x, V);
      __syncthreads();
      for (int t = threadIdx.x; t < N; t += blockDim.x) V[t] = b[t] - V[t];
      __syncthreads();
      blk_apply_prec<D>(E, W, pdeg, V, w);
      double p2 = 0.0;
      for (int t = threadIdx.x; t < N; t += blockDim.x) p2 += w[t] * w[t];
      beta = sqrt(block_sum(p2, sh));
      resid = beta;
      if (beta <= tol) break;
    }
    first = false;
    {
      const double inv = 1.0 / beta;
      for (int t = threadIdx.x; t < N; t += blockDim.x) V[t] = w[t] * inv;
    }
    if (threadIdx.x == 0) g[0] = beta;
    __syncthreads();

    int k = 0;
    bool done = false;
    for (; k < m && !done; ++k) {
      const double* vk = V + (size_t)k * N;
      long long t0 = clock64();
      blk_apply_op<D>(E, W, pdeg, vk, w);
      long long t1 = clock64();
      // CGS2: two classical Gram-Schmidt passes
      blk_dots(V, N, k + 1, w, sh.h);
      __syncthreads();
      blk_project_out(V, N, k + 1, sh.h, w);
      __syncthreads();
      blk_dots(V, N, k + 1, w, sh.h2);
      __syncthreads();
      const double p2 = blk_project_out(V, N, k + 1, sh.h2, w);
      const double hk1 = sqrt(block_sum(p2, sh));
      long long t2 = clock64();
      if (threadIdx.x == 0) {
        // Hessenberg column, previous Givens rotations, new rotation
        // the running pair (a, b) of the column stays in registers and the operands of the
        // next rotation are fetched before the store of the current one: one thread, one
        // dependent chain -- only the multiply-adds remain on it
        double* hc = H + (size_t)k * ldh;
        double a = sh.h[0] + sh.h2[0];
        double cj = cs[0], sj = sn[0];
        double b = k > 0 ? sh.h[1] + sh.h2[1] : 0.0;
        for (int j = 0; j < k; ++j) {
          const double cn = cs[j + 1], sn_next = sn[j + 1];       // (index k is rewritten below)
          const double bn = j + 2 <= k ? sh.h[j + 2] + sh.h2[j + 2] : 0.0;
          const double t0 = cj * a + sj * b;
          const double b2 = -sj * a + cj * b;
          hc[j] = t0;
          a = b2;
          b = bn;
          cj = cn;
          sj = sn_next;
        }
        const double hkk = a;
        const double denom = sqrt(hkk * hkk + hk1 * hk1);
        const double c = (denom == 0.0) ? 1.0 : hkk / denom;
        const double s = (denom == 0.0) ? 0.0 : hk1 / denom;
        cs[k] = c;
        sn[k] = s;
        hc[k] = c * hkk + s * hk1;
        g[k + 1] = -s * g[k];
        g[k] = c * g[k];
        sh.bc[0] = fabs(g[k + 1]);
      }
      __syncthreads();
      resid = sh.bc[0];
      ++iters;
      if (hk1 > 0.0) {
        double* vn = V + (size_t)(k + 1) * N;
        const double inv = 1.0 / hk1;
        for (int t = threadIdx.x; t < N; t += blockDim.x) vn[t] = w[t] * inv;
      }
      if (resid <= tol || iters >= opt.gmres_max_iter || hk1 == 0.0) done = true;
      __syncthreads();
      if (threadIdx.x == 0) {
        const long long t3 = clock64();
        sh.cyc[1] += t1 - t0;
        sh.cyc[2] += t2 - t1;
        sh.cyc[3] += t3 - t2;
      }
    }
    long long t4 = clock64();
    // y = H^{-1} g  (k x k upper triangular) by column-oriented back substitution on one
    // warp (k dependent steps instead of k^2/2), then x += V y
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      for (int i = k - 1; i >= 0; --i) {
        const double yi = g[i] / H[(size_t)i * ldh + i];
        __syncwarp();
        for (int j = lane; j < i; j += 32) g[j] -= H[(size_t)i * ldh + j] * yi;
        if (lane == 0) y[i] = yi;
        __syncwarp();
      }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < N; t += blockDim.x) {
      double s = 0.0;
      for (int j = 0; j < k; ++j) s += y[j] * V[(size_t)j * N + t];
      x[t] += s;
    }
    __syncthreads();
    if (threadIdx.x == 0) sh.cyc[4] += clock64() - t4;
    if (resid <= tol || iters >= opt.gmres_max_iter) break;
  }
  *resid_out = resid;
  return iters;
}


// FenicsModel.solve_state1: Newton on F_u(u1) = 0 starting from the guess held in VF_U1,
// then v1, a1 from the Newmark relations (App. C, Q2).
// Residual (+ Jacobian) of the whole member mesh by the two-phase record algorithm of
// asm_tile2_kernel, inside the CTA: one 144-byte record per cell (thread per cell), then one
// thread per scalar row walks the vertex fan and completes its CSR row / residual entry from
// the records, then the boundary nodes add facet terms and Dirichlet rows.  Every cell's
// geometry, material and Newmark arithmetic is done once instead of once per adjacent vertex.
// recs: ne * kRec2D doubles (the Krylov basis storage, idle during assembly).
template <bool JAC>
__device__ void blk_assemble_records(const EngineDev& E, const PropView& pv, const StateView& sv,
                                     double* recs, double* Jv, double* F) {
  constexpr int D = 2;
  const MeshView& m = E.mesh;
  const NewmarkCoef nc = newmark_coef(sv.dt);
  const LameFac lf = lame_fac(pv.scal[SC_NU]);
  const Damping dp = prop_damping(pv);
  const bool is_static = sv.is_static != 0;
  for (int e = threadIdx.x; e < m.ne; e += blockDim.x) {
    int nd[3];
    double x[3][2];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      nd[a] = m.cells[(size_t)a * m.ne + e];
      const D2 c2 = reinterpret_cast<const D2*>(m.xy)[nd[a]];
      x[a][0] = c2.x;
      x[a][1] = c2.y;
    }
    tri_record(x, nd, pv.emod[e], lf, pv.eta[e], pv.rho[e], dp, nc, is_static, true, sv.u1, sv.u0,
               sv.v0, sv.a0, recs + (size_t)e * kRec2D);
  }
  __syncthreads();
  for (int r = threadIdx.x; r < E.N; r += blockDim.x) {
    const int n = r >> 1, comp = r & 1;
    const int b0 = m.brptr[n], deg = m.brptr[n + 1] - b0;
    double* row = Jv + (size_t)D * D * b0 + comp * D * deg;
    const int qb = m.n2e_ptr[n], qe = m.n2e_ptr[n + 1];
    double racc = 0.0;
    if (qe > qb) {
      unsigned info = E.gpair[qb];
      const double* rec = recs + (size_t)(info & 0xfffu) * kRec2D;
      int a = (info >> 12) & 3;
      D2 diag = D2{0.0, 0.0}, carry = D2{0.0, 0.0}, first = D2{0.0, 0.0};
      int slot_first = 0, slot_carry = 0;
      if (JAC) {
        tri_row_fan(rec, a, comp, diag, first, carry);
        slot_first = (info >> 20) & 63;
        slot_carry = (info >> 26) & 63;
      }
      racc = rec[9 + 2 * a + comp];
      for (int q = qb + 1; q < qe; ++q) {
        info = E.gpair[q];
        rec = recs + (size_t)(info & 0xfffu) * kRec2D;
        a = (info >> 12) & 3;
        if (JAC) {
          D2 ws, wn, wp;
          tri_row_fan(rec, a, comp, ws, wn, wp);
          diag.x += ws.x;
          diag.y += ws.y;
          *reinterpret_cast<D2*>(row + D * ((info >> 20) & 63)) = D2{carry.x + wn.x, carry.y + wn.y};
          carry = wp;
          slot_carry = (info >> 26) & 63;
        }
        racc += rec[9 + 2 * a + comp];
      }
      if (JAC) {
        if (slot_carry == slot_first) {
          *reinterpret_cast<D2*>(row + D * slot_first) = D2{first.x + carry.x, first.y + carry.y};
        } else {
          *reinterpret_cast<D2*>(row + D * slot_first) = first;
          *reinterpret_cast<D2*>(row + D * slot_carry) = carry;
        }
        *reinterpret_cast<D2*>(row + D * ((info >> 14) & 63)) = diag;
      }
    }
    F[r] = racc;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < E.n_touch; t += blockDim.x) {
    const int i = E.touch[t];
    double res[D] = {F[D * i], F[D * i + 1]};
    assemble_node_facets_bc<D, JAC, true>(i, m, pv, sv, Jv + (size_t)D * D * m.brptr[i], res);
    F[D * i] = res[0];
    F[D * i + 1] = res[1];
  }
  __syncthreads();
}

template <int D>
__device__ void blk_solve_solid(const EngineDev& E, double* mb, const SolverWork& W, double dt,
                                const SolverOpts& opt, BlockShared& sh, bool allow_dense) {
  const Layout& L = E.L;
  const int N = E.N, nn = E.mesh.nn;
  double* u1 = mb + L.off[VF_U1];
  double* F = W.F;
  double* Jv = W.J;
  double* dx = W.dx;
  double* info = mb + L.off[VF_INFO];
  const PropView pv = member_props<D>(E, mb);
  StateView sv;
  sv.u1 = u1;
  sv.u0 = opt.is_static ? u1 : mb + L.off[VF_U0];
  sv.v0 = mb + L.off[VF_V0];
  sv.a0 = mb + L.off[VF_A0];
  sv.p1 = mb + L.off[VF_P1];
  sv.dt = dt;
  sv.is_static = opt.is_static;
  sv.mix = jac_mix_du1(newmark_coef(sv.dt), opt.is_static != 0);

  // record-based assembly when the tables exist and the records fit the Krylov basis storage
  const bool use_records = E.gpair != nullptr &&
                           (size_t)E.mesh.ne * kRec2D <= (size_t)(E.restart + 1) * N;
  int k = 0;
  double r0 = 0.0, abs_err = 0.0, rel_err = 0.0;
  int gm_iters = 0;
  double gm_resid = 0.0, gm_bnorm = 0.0;
  while (true) {
    // residual (and, in the first iteration, the Jacobian in the same sweep)
    const long long ta = clock64();
    double part = 0.0;
    if (D == 2 && use_records) {
      if (k == 0) blk_assemble_records<true>(E, pv, sv, W.V, Jv, F);
      else blk_assemble_records<false>(E, pv, sv, W.V, Jv, F);
      for (int t = threadIdx.x; t < N; t += blockDim.x) part += F[t] * F[t];
    } else {
      for (int i = threadIdx.x; i < nn; i += blockDim.x) {
        double res[D];
        if (k == 0)
          assemble_node<D, true, true>(i, E.mesh, pv, sv, Jv + (size_t)D * D * E.mesh.brptr[i], res);
        else
          assemble_node<D, false, true>(i, E.mesh, pv, sv, nullptr, res);
        for (int c = 0; c < D; ++c) {
          F[D * i + c] = res[c];
          part += res[c] * res[c];
        }
      }
    }
    abs_err = sqrt(block_sum(part, sh));
    if (threadIdx.x == 0) sh.cyc[0] += clock64() - ta;
    if (k == 0) r0 = abs_err;
    rel_err = (r0 > 0.0) ? abs_err / r0 : 0.0;
    if (abs_err <= opt.newton_abs_tol || rel_err <= opt.newton_rel_tol ||
        k >= opt.newton_max_iter)
      break;
    if (k > 0) {
      if (D == 2 && use_records) {
        blk_assemble_records<true>(E, pv, sv, W.V, Jv, F);   // (F is rewritten identically)
      } else {
        for (int i = threadIdx.x; i < nn; i += blockDim.x) {
          double res[D];
          assemble_node<D, true, false>(i, E.mesh, pv, sv, Jv + (size_t)D * D * E.mesh.brptr[i], res);
        }
      }
    }
    __syncthreads();
    blk_compute_dinv<D>(E, Jv, W.Dinv);
    __syncthreads();
    bool dense = false;
    if (W.P && allow_dense && opt.is_static) {
      // static solve: the inverse of the actual Newton matrix, rebuilt at the first solve of
      // every launch (member_kernel raises pstate[1]) and whenever the last solve with it needed
      // more than a few iterations (contact has changed the matrix)
      if (W.pstate[0] != -1.0 || W.pstate[1] != 0.0) {
        const long long tp = clock64();
        __syncthreads();
        blk_dense_inverse<D>(E, Jv, W.P, W.Pf, W.V, W.V + (size_t)kDenseNb * N, sh.B);
        if (threadIdx.x == 0) {
          W.pstate[0] = -1.0;
          W.pstate[1] = 0.0;
          W.pstate[2] = 0.0;  // the transient inverse (below) is gone
          sh.cyc[7] += clock64() - tp;
        }
        __syncthreads();
      }
      dense = true;
    } else if (W.P && allow_dense) {
      // time loop: the inverse of the STATE-INDEPENDENT part ca M + cv C + K (+ membrane) with
      // Dirichlet rows -- a pure function of (dt, properties), so it survives launches and the
      // trajectory does not depend on when it was built.  It is validated once per launch by a
      // checksum of the properties and rebuilt when that or dt changes.
      bool fresh = sh.bc[6] != 0.0 && fabs(W.pstate[2] - dt) <= 1e-6 * dt;
      if (!fresh) {
        double cpart = 0.0;
        for (int e = threadIdx.x; e < E.mesh.ne; e += blockDim.x) {
          const double we = 1.0 + 1e-3 * e;
          cpart += we * (pv.emod[e] + 0.5 * pv.eta[e] + 0.25 * pv.rho[e]);
          if (pv.membrane) cpart += we * (pv.emod_m[e] + 0.5 * pv.nu_m[e] + 0.25 * pv.th_m[e]);
        }
        double chk = block_sum(cpart, sh);
        for (int q = 0; q < SC_COUNT; ++q) {
          const double v = pv.scal[q];
          if (v == v && fabs(v) < 1e300) chk += (q + 1) * v;  // (ycontact may be +-inf)
        }
        const bool ok = W.pstate[3] == chk && fabs(W.pstate[2] - dt) <= 1e-6 * dt &&
                        W.pstate[0] != -1.0;
        __syncthreads();
        if (!ok) {
          const long long tp = clock64();
          StateView sc = sv;
          sc.mix.p = 0.0;
          PropView pc = pv;
          pc.contact = 0;
          for (int i = threadIdx.x; i < nn; i += blockDim.x) {
            double res[D];
            assemble_node<D, true, false>(i, E.mesh, pc, sc, Jv + (size_t)D * D * E.mesh.brptr[i], res);
          }
          __syncthreads();
          blk_dense_inverse<D>(E, Jv, W.P, W.Pf, W.V, W.V + (size_t)kDenseNb * N, sh.B);
          // put the actual Newton matrix back
          for (int i = threadIdx.x; i < nn; i += blockDim.x) {
            double res[D];
            assemble_node<D, true, false>(i, E.mesh, pv, sv, Jv + (size_t)D * D * E.mesh.brptr[i], res);
          }
          __syncthreads();
          if (threadIdx.x == 0) {
            W.pstate[0] = 0.0;
            W.pstate[2] = dt;
            W.pstate[3] = chk;
            sh.cyc[7] += clock64() - tp;
          }
        }
        if (threadIdx.x == 0) sh.bc[6] = 1.0;
        __syncthreads();
      }
      dense = true;
    }
    const int its = blk_gmres<D>(E, W, F, dx, opt, sh, &gm_resid, &gm_bnorm, dense);
    gm_iters += its;
    if (dense && opt.is_static && its > 4 && threadIdx.x == 0) W.pstate[1] = 1.0;
    __syncthreads();
    for (int t = threadIdx.x; t < N; t += blockDim.x) u1[t] -= dx[t];
    __syncthreads();
    ++k;
  }
  // Newmark velocity / acceleration
  if (!opt.is_static) {
    const double* u0 = mb + L.off[VF_U0];
    const double* v0 = mb + L.off[VF_V0];
    const double* a0 = mb + L.off[VF_A0];
    double* v1 = mb + L.off[VF_V1];
    double* a1 = mb + L.off[VF_A1];
    for (int t = threadIdx.x; t < N; t += blockDim.x) {
      v1[t] = newmark_v(u1[t], u0[t], v0[t], a0[t], dt);
      a1[t] = newmark_a(u1[t], u0[t], v0[t], a0[t], dt);
    }
  }
  if (threadIdx.x == 0) {
    info[INFO_NUM_ITER] = double(k);
    info[INFO_ABS_ERR] = abs_err;
    info[INFO_REL_ERR] = rel_err;
    info[INFO_GMRES_ITERS] = double(gm_iters);
    info[INFO_GMRES_RESID] = gm_resid;
    info[INFO_BNORM] = gm_bnorm;
  }
  __syncthreads();
}

// _set_fin_solid_state (transient.py:836-848) + fluid solve: area from u1, Bernoulli -> q1, pf1
template <int D>
__device__ void blk_fluid(const EngineDev& E, double* mb, BlockShared& sh) {
  const Layout& L = E.L;
  const double* u1 = mb + L.off[VF_U1];
  double* area = mb + L.off[VF_AREA];
  const double ymid = (mb + L.off[VF_SCAL])[SC_YMID];
  for (int k = threadIdx.x; k < E.n_fsi; k += blockDim.x) {
    const int i = E.fsi_solid[k];
    const double y = E.mesh.xyz[(size_t)1 * E.mesh.nn + i] + u1[D * i + 1];
    area[E.fsi_fluid[k]] = 2.0 * (ymid - y);
  }
  __syncthreads();
  const int wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int f = wid; f < E.n_fluid; f += nw) {
    bernoulli_channel(E.fluid_kind, E.idx_sep, E.ns, E.s + (size_t)f * E.ns,
                      area + (size_t)f * E.ns, (mb + L.off[VF_PSUB])[f],
                      (mb + L.off[VF_PSUP])[f], mb + L.off[VF_FPROP] + (size_t)f * FP_COUNT,
                      mb + L.off[VF_Q1] + f, mb + L.off[VF_PF1] + (size_t)f * E.ns);
  }
  __syncthreads();
  // glottal-width signal: min over the fluid area (postprocess/solid.py:487-501)
  double mn = CUDART_INF;
  for (int k = threadIdx.x; k < E.n_fluid * E.ns; k += blockDim.x) mn = fmin(mn, area[k]);
  mn = -warp_max(-mn);
  const int lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh.red[wid] = mn;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = sh.red[0];
    for (int q = 1; q < nw; ++q) t = fmin(t, sh.red[q]);
    (mb + L.off[VF_INFO])[INFO_MIN_AREA] = t;
  }
  __syncthreads();
}

}  // namespace vf

