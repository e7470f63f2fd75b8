// Banded LU of J_uu in a bandwidth-reducing ordering: the direct "sparse LU" stand-in for the
// reference's dfn.solve(A, x, b, 'petsc') (/root/reference/src/femvf/models/transient.py:487,
// static.py:140) on meshes of up to a few 1e4 DOF that do not fit the one-CTA solver.  No
// cuSPARSE / cuSOLVER: fill, factorisation and triangular solves are the kernels below.
//
// The DOFs are renumbered on the host (reverse Cuthill-McKee on the node graph, gridsolve.py), the
// matrix is scattered into row-major band storage AB[i][j - i + b] (b = half bandwidth) and
// factorised in place without pivoting (unit-diagonal L below, U on and above the diagonal): J_uu
// is the sum of a positive definite stiffness / mass part, the follower-pressure block and
// identity rows, for which the elimination is stable; the caller adds one step of iterative
// refinement with the device SpMV.
// One CTA of 1024 threads walks the pivots in panels of 16 (band_lu_kernel).  Cost N b^2
// multiply-adds: 4 118 DOF (b = 101) 0.04 G, 16 074 DOF (b = 197) 0.6 G.
#include "engine_internal.h"

namespace vf {

namespace {

__global__ void band_fill_kernel(MeshView m, int d, const double* __restrict__ J,
                                 const int* __restrict__ perm, double* __restrict__ AB, int b) {
  // one thread per (node block row, block): scatters its d x d entries
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int nn = m.nn;
  const long long nblk = m.brptr[nn];
  if (t >= nblk) return;
  // row of block t: binary search in brptr
  int lo = 0, hi = nn;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (m.brptr[mid] <= t) lo = mid; else hi = mid;
  }
  const int i = lo, b0 = m.brptr[i], deg = m.brptr[i + 1] - b0, s = (int)(t - b0);
  const int j = m.bcol[t];
  const int W = 2 * b + 1;
  const double* blk = J + (size_t)d * d * b0;
  for (int a = 0; a < d; ++a)
    for (int c = 0; c < d; ++c) {
      const int pi = perm[d * i + a], pj = perm[d * j + c];
      AB[(size_t)pi * W + (pj - pi + b)] = blk[a * d * deg + s * d + c];
    }
}

// Blocked right-looking factorisation, kNB pivots per pass, one CTA of 1024 threads:
//   A  the panel -- the kNB pivot rows (full band width) and the kNB pivot columns below them --
//      is loaded into shared memory and factorised there (sequential over the pivots, parallel
//      over rows / columns);
//   B  multipliers and finished U rows go back to the band;
//   C  the trailing (b x b) window gets its rank-kNB update: every element is read and written
//      ONCE per pass with up to kNB multiply-adds from shared memory.  (Updating it pivot by pivot
//      costs one L2 round trip per element and pivot: 8000 cycles per pivot at b = 100.)
constexpr int kNB = 16;

__global__ void __launch_bounds__(1024, 1) band_lu_kernel(double* __restrict__ AB, int N, int b) {
  extern __shared__ double s_band[];
  const int W = 2 * b + 1;
  const int LW = b + kNB;                 // rows of the column panel / columns of the row panel
  double* Up = s_band;                    // Up[p][c]: row k0 + p, column k0 + c   (kNB x LW)
  double* Lp = s_band + (size_t)kNB * LW; // Lp[r][p]: row k0 + kNB + r, column k0 + p   (b x kNB)
  const int tid = threadIdx.x, nt = blockDim.x;
  auto at = [&](int r, int c) -> double& { return AB[(size_t)r * W + (c - r + b)]; };
  for (int k0 = 0; k0 < N; k0 += kNB) {
    const int np = min(kNB, N - k0);              // pivots of this pass
    const int nc = min(LW, N - k0);               // columns k0 .. k0 + nc - 1 touched by the pass
    const int nr = min(b, N - k0 - np);           // rows below the pivot rows touched by the pass
    // ---- A: load the panel (entries outside the band are zero)
    for (int t = tid; t < np * LW; t += nt) {
      const int p = t / LW, c = t % LW;
      Up[p * LW + c] = (c < nc && c - p <= b && p - c <= b) ? at(k0 + p, k0 + c) : 0.0;
    }
    for (int t = tid; t < nr * kNB; t += nt) {
      const int r = t / kNB, p = t % kNB;
      const int row = k0 + np + r, col = k0 + p;
      Lp[r * kNB + p] = (p < np && row - col <= b) ? at(row, col) : 0.0;
    }
    __syncthreads();
    for (int p = 0; p < np; ++p) {
      const double inv = 1.0 / Up[p * LW + p];
      // multipliers of column p: pivot rows below p (kept in Up's lower triangle) and Lp rows
      for (int t = tid; t < (np - 1 - p) + nr; t += nt) {
        if (t < np - 1 - p) Up[(p + 1 + t) * LW + p] *= inv;
        else Lp[(t - (np - 1 - p)) * kNB + p] *= inv;
      }
      __syncthreads();
      // eliminate: pivot rows q > p over their whole width, Lp rows over the panel columns > p
      const int wq = nc - (p + 1);
      for (int t = tid; t < (np - 1 - p) * wq; t += nt) {
        const int q = p + 1 + t / wq, c = p + 1 + t % wq;
        Up[q * LW + c] -= Up[q * LW + p] * Up[p * LW + c];
      }
      const int wl = np - 1 - p;
      for (int t = tid; t < nr * wl; t += nt) {
        const int r = t / wl, c = p + 1 + t % wl;
        Lp[r * kNB + c] -= Lp[r * kNB + p] * Up[p * LW + c];
      }
      __syncthreads();
    }
    // ---- B: write the panel back
    for (int t = tid; t < np * LW; t += nt) {
      const int p = t / LW, c = t % LW;
      if (c < nc && c - p <= b && p - c <= b) at(k0 + p, k0 + c) = Up[p * LW + c];
    }
    for (int t = tid; t < nr * kNB; t += nt) {
      const int r = t / kNB, p = t % kNB;
      const int row = k0 + np + r, col = k0 + p;
      if (p < np && row - col <= b) at(row, col) = Lp[r * kNB + p];
    }
    // ---- C: rank-np update of the trailing window (rows and columns k0 + np .. k0 + np + nr - 1)
    for (int t = tid; t < nr * nr; t += nt) {
      const int r = t / nr, c = t % nr;
      if (r - c > b || c - r > b) continue;
      double acc = 0.0;
#pragma unroll 4
      for (int p = 0; p < np; ++p) acc += Lp[r * kNB + p] * Up[p * LW + np + c];
      at(k0 + np + r, k0 + np + c) -= acc;
    }
    __syncthreads();
  }
}

// x := U^-1 L^-1 x in the band ordering (one CTA), blocked like the factorisation: the kNB x kNB
// triangular block of a panel is solved by one warp (lane q owns row q, the finished unknown is
// broadcast by shuffle), then every other row of the band gets its rank-kNB update from kNB
// CONTIGUOUS entries of its row -- N / kNB barrier pairs instead of N.
// SMEM: the vector lives in shared memory during the sweeps.
template <bool SMEM>
__global__ void __launch_bounds__(1024, 1) band_solve_kernel(const double* __restrict__ AB,
                                                            double* __restrict__ xg, int N, int b) {
  extern __shared__ double s_x[];
  const int W = 2 * b + 1;
  double* x = SMEM ? s_x : xg;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
  auto at = [&](int r, int c) -> double { return AB[(size_t)r * W + (c - r + b)]; };
  if (SMEM) {
    for (int i = tid; i < N; i += nt) s_x[i] = xg[i];
    __syncthreads();
  }
  // ---- forward: L y = x, unit lower triangular
  for (int k0 = 0; k0 < N; k0 += kNB) {
    const int np = min(kNB, N - k0);
    if (tid < 32) {
      // the lane's row of the triangular block first (independent loads: one round trip)
      double lrow[kNB];
#pragma unroll
      for (int p = 0; p < kNB; ++p) lrow[p] = (p < lane && lane < np) ? at(k0 + lane, k0 + p) : 0.0;
      double xv = lane < np ? x[k0 + lane] : 0.0;
#pragma unroll
      for (int p = 0; p < kNB; ++p) {
        const double xp = __shfl_sync(0xffffffffu, xv, p);
        xv -= lrow[p] * xp;
      }
      if (lane < np) x[k0 + lane] = xv;
    }
    __syncthreads();
    const int nr = min(b, N - k0 - np);
    for (int r = tid; r < nr; r += nt) {
      const int row = k0 + np + r;
      double acc = 0.0;
      for (int p = 0; p < np; ++p)
        if (row - (k0 + p) <= b) acc += at(row, k0 + p) * x[k0 + p];
      x[row] -= acc;
    }
    __syncthreads();
  }
  // ---- backward: U z = y
  const int last = ((N - 1) / kNB) * kNB;
  for (int k0 = last; k0 >= 0; k0 -= kNB) {
    const int np = min(kNB, N - k0);
    if (tid < 32) {
      double urow[kNB];
#pragma unroll
      for (int p = 0; p < kNB; ++p)
        urow[p] = (p >= lane && p < np && lane < np) ? at(k0 + lane, k0 + p) : 0.0;
      double dinv = 1.0;
#pragma unroll
      for (int p = 0; p < kNB; ++p)
        if (p == lane && lane < np) dinv = 1.0 / urow[p];
      double xv = lane < np ? x[k0 + lane] : 0.0;
#pragma unroll
      for (int p = kNB - 1; p >= 0; --p) {
        if (lane == p) xv *= dinv;
        const double xp = __shfl_sync(0xffffffffu, xv, p);
        if (lane < p) xv -= urow[p] * xp;
      }
      if (lane < np) x[k0 + lane] = xv;
    }
    __syncthreads();
    const int nr = min(b, k0);
    for (int r = tid; r < nr; r += nt) {
      const int row = k0 - 1 - r;
      double acc = 0.0;
      for (int p = 0; p < np; ++p)
        if ((k0 + p) - row <= b) acc += at(row, k0 + p) * x[k0 + p];
      x[row] -= acc;
    }
    __syncthreads();
  }
  if (SMEM) {
    for (int i = tid; i < N; i += nt) xg[i] = s_x[i];
  }
}

__global__ void band_permute_kernel(const double* __restrict__ src, double* __restrict__ dst,
                                    const int* __restrict__ perm, int N, int to_band) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  if (to_band) dst[perm[i]] = src[i];
  else dst[i] = src[perm[i]];
}

}  // namespace

void band_release(vf_engine* e) {
  if (e->band.mem) cudaFree(e->band.mem);
  e->band = BandState{};
}

}  // namespace vf

using namespace vf;

extern "C" {

int vf_band_setup(vf_engine* e, const int32_t* perm_host, int half_bandwidth, void* stream) {
  if (!e) return fail("null engine");
  if (!perm_host || half_bandwidth < 1) return fail("vf_band_setup: bad ordering");
  const size_t N = (size_t)e->desc.dim * e->desc.nn;
  if ((size_t)half_bandwidth >= N) return fail("vf_band_setup: bandwidth exceeds the matrix size");
  const size_t W = 2 * (size_t)half_bandwidth + 1;
  const size_t b_ab = align_up(sizeof(double) * N * W, 256);
  if (b_ab > ((size_t)8 << 30)) return fail("vf_band_setup: band storage above 8 GB");
  const size_t b_perm = align_up(sizeof(int) * N, 256), b_x = align_up(sizeof(double) * N, 256);
  band_release(e);
  char* mem = nullptr;
  VF_CUDA(cudaMalloc(&mem, b_ab + b_perm + b_x));
  BandState& S = e->band;
  S.mem = mem;
  S.AB = reinterpret_cast<double*>(mem);
  S.perm = reinterpret_cast<int*>(mem + b_ab);
  S.x = reinterpret_cast<double*>(mem + b_ab + b_perm);
  S.b = half_bandwidth;
  S.N = (int)N;
  cudaStream_t st = as_stream(stream);
  VF_CUDA(cudaMemcpyAsync(S.perm, perm_host, sizeof(int) * N, cudaMemcpyHostToDevice, st));
  VF_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int vf_band_factor(vf_engine* e, int member, void* stream) {
  if (!e) return fail("null engine");
  if (member < 0 || member >= e->desc.n_members) return fail("member out of range");
  BandState& S = e->band;
  if (!S.mem) return fail("vf_band_factor: call vf_band_setup first");
  cudaStream_t st = as_stream(stream);
  const size_t W = 2 * (size_t)S.b + 1;
  VF_CUDA(cudaMemsetAsync(S.AB, 0, sizeof(double) * (size_t)S.N * W, st));
  const long long nblk = e->brptr[e->desc.nn];
  const int block = 256;
  band_fill_kernel<<<(unsigned)((nblk + block - 1) / block), block, 0, st>>>(
      e->dev.mesh, e->desc.dim, member_array(e, VF_J, member), S.perm, S.AB, S.b);
  const size_t smem = sizeof(double) * ((size_t)kNB * (S.b + kNB) + (size_t)S.b * kNB);
  if (smem > 200 * 1024) return fail("vf_band_factor: bandwidth too large for the factor kernel");
  VF_CUDA(cudaFuncSetAttribute(band_lu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  band_lu_kernel<<<1, 1024, smem, st>>>(S.AB, S.N, S.b);
  e->launches += 2;
  VF_CUDA(cudaGetLastError());
  return 0;
}

int vf_band_solve(vf_engine* e, const double* b_dev, double* x_dev, void* stream) {
  if (!e) return fail("null engine");
  BandState& S = e->band;
  if (!S.mem) return fail("vf_band_solve: call vf_band_setup first");
  if (!b_dev || !x_dev) return fail("vf_band_solve: null vector");
  cudaStream_t st = as_stream(stream);
  const int block = 256, grid = (S.N + block - 1) / block;
  band_permute_kernel<<<grid, block, 0, st>>>(b_dev, S.x, S.perm, S.N, 1);
  const size_t smem = sizeof(double) * (size_t)S.N;
  if (smem <= 200 * 1024) {
    VF_CUDA(cudaFuncSetAttribute(band_solve_kernel<true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    band_solve_kernel<true><<<1, 1024, smem, st>>>(S.AB, S.x, S.N, S.b);
  } else {
    band_solve_kernel<false><<<1, 1024, 0, st>>>(S.AB, S.x, S.N, S.b);
  }
  band_permute_kernel<<<grid, block, 0, st>>>(S.x, x_dev, S.perm, S.N, 0);
  e->launches += 3;
  VF_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
