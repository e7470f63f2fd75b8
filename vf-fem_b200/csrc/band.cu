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
// One CTA of 1024 threads walks the N pivots; every step updates a (b x b) window that lives in
// L2.  Cost N b^2 multiply-adds: 4 118 DOF / b = 150: ~4 ms; 16 074 DOF / b = 300: ~60 ms.
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

__global__ void __launch_bounds__(1024, 1) band_lu_kernel(double* __restrict__ AB, int N, int b) {
  extern __shared__ double s_band[];
  double* sl = s_band;          // multipliers of the current column
  double* su = s_band + b + 1;  // row k of U right of the diagonal
  const int W = 2 * b + 1;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int k = 0; k < N - 1; ++k) {
    const int nr = min(b, N - 1 - k);
    const double inv = 1.0 / AB[(size_t)k * W + b];
    for (int i = threadIdx.x + 1; i <= nr; i += blockDim.x) {
      const size_t at = (size_t)(k + i) * W + (b - i);
      const double l = AB[at] * inv;
      AB[at] = l;
      sl[i] = l;
      su[i] = AB[(size_t)k * W + b + i];
    }
    __syncthreads();
    for (int i = ty + 1; i <= nr; i += 32) {
      const double l = sl[i];
      if (l != 0.0) {
        double* row = AB + (size_t)(k + i) * W + (b - i);
        for (int j = tx + 1; j <= nr; j += 32) row[j] -= l * su[j];
      }
    }
    __syncthreads();
  }
}

// x := U^-1 L^-1 x in the band ordering (one CTA).  SMEM: the vector lives in shared memory for
// the N sequential steps (a few hundred cycles each from global memory, ~100 from shared).
template <bool SMEM>
__global__ void __launch_bounds__(1024, 1) band_solve_kernel(const double* __restrict__ AB,
                                                            double* __restrict__ xg, int N, int b) {
  extern __shared__ double s_x[];
  const int W = 2 * b + 1;
  double* x = SMEM ? s_x : xg;
  if (SMEM) {
    for (int i = threadIdx.x; i < N; i += blockDim.x) s_x[i] = xg[i];
    __syncthreads();
  }
  for (int k = 0; k < N - 1; ++k) {
    const int nr = min(b, N - 1 - k);
    const double xk = x[k];
    for (int i = threadIdx.x + 1; i <= nr; i += blockDim.x)
      x[k + i] -= AB[(size_t)(k + i) * W + (b - i)] * xk;
    __syncthreads();
  }
  for (int k = N - 1; k >= 0; --k) {
    const double xk = x[k] / AB[(size_t)k * W + b];   // every thread: same value, no broadcast
    const int nr = min(b, k);
    for (int i = threadIdx.x + 1; i <= nr; i += blockDim.x)
      x[k - i] -= AB[(size_t)(k - i) * W + (b + i)] * xk;
    __syncthreads();
    if (threadIdx.x == 0) x[k] = xk;
  }
  if (SMEM) {
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) xg[i] = s_x[i];
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
  const size_t smem = sizeof(double) * 2 * ((size_t)S.b + 1);
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
