// Multicolour block ILU(0) preconditioner of J_uu on the node-block CSR pattern: the stand-in for
// the reference's sparse LU (dfn.solve(A, x, b, 'petsc'),
// /root/reference/src/femvf/models/transient.py:487, static.py:140) on meshes that do not fit
// one CTA.  No cuSPARSE: factorisation and both triangular sweeps are kernels of this file.
//
// The nodes of the solved range are coloured on the host (tables.color_node_graph) so that no
// two nodes of a colour share a cell.  Eliminating colour after colour, every block row of a
// colour depends only on rows of earlier colours: the factorisation is `ncolors` launches with
// one thread per block row, and M^-1 r = U^-1 L^-1 r is ncolors - 1 forward and ncolors backward
// launches, each a sparse block mat-vec over the rows of one colour (4 lanes per row, like
// spmv_kernel).  L (unit block diagonal) and U share the storage and the pattern of J; the
// inverses of the diagonal blocks of U are kept apart.
// HBM traffic of one application: every block of the factor once (8 nnz bytes), the column index
// and the colour of every block twice, O(N) vector data -- about 1.3 SpMV.
#include "engine_internal.h"

namespace vf {

namespace {

template <int D>
__device__ __forceinline__ void load_block(const double* rowblk, int deg, int s, double (&B)[D][D]) {
#pragma unroll
  for (int a = 0; a < D; ++a)
#pragma unroll
    for (int c = 0; c < D; ++c) B[a][c] = rowblk[a * D * deg + s * D + c];
}
template <int D>
__device__ __forceinline__ void store_block(double* rowblk, int deg, int s, const double (&B)[D][D]) {
#pragma unroll
  for (int a = 0; a < D; ++a)
#pragma unroll
    for (int c = 0; c < D; ++c) rowblk[a * D * deg + s * D + c] = B[a][c];
}
template <int D>
__device__ __forceinline__ void invert_block(const double (&A)[D][D], double* o) {
  if constexpr (D == 2) {
    const double inv = 1.0 / (A[0][0] * A[1][1] - A[0][1] * A[1][0]);
    o[0] = A[1][1] * inv;
    o[1] = -A[0][1] * inv;
    o[2] = -A[1][0] * inv;
    o[3] = A[0][0] * inv;
  } else {
    double c0[3], c1[3], c2[3];
    cross3(A[1], A[2], c0);
    cross3(A[2], A[0], c1);
    cross3(A[0], A[1], c2);
    const double inv = 1.0 / (A[0][0] * c0[0] + A[0][1] * c0[1] + A[0][2] * c0[2]);
    for (int k = 0; k < 3; ++k) {
      o[k * 3 + 0] = c0[k] * inv;
      o[k * 3 + 1] = c1[k] * inv;
      o[k * 3 + 2] = c2[k] * inv;
    }
  }
}

// Rows of colour c (IKJ variant): for every lower-colour neighbour k, in elimination order,
// L_ik = A_ik U_kk^-1, then A_ij -= L_ik U_kj for the blocks (k, j) of U that exist in row i.
template <int D>
__global__ void ilu_factor_color_kernel(MeshView m, double* __restrict__ LU,
                                        double* __restrict__ Dinv, const int* __restrict__ color,
                                        const int* __restrict__ rows, int count, int c) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const int i = rows[t];
  const int b0 = m.brptr[i], deg = m.brptr[i + 1] - b0;
  const int* col_i = m.bcol + b0;
  double* row_i = LU + (size_t)D * D * b0;
  for (int cc = 0; cc < c; ++cc) {
    for (int s = 0; s < deg; ++s) {
      const int k = col_i[s];
      if (color[k] != cc) continue;
      double B[D][D], Lik[D][D];
      load_block<D>(row_i, deg, s, B);
      const double* dk = Dinv + (size_t)D * D * k;
#pragma unroll
      for (int a = 0; a < D; ++a)
#pragma unroll
        for (int b = 0; b < D; ++b) {
          double v = 0.0;
#pragma unroll
          for (int q = 0; q < D; ++q) v += B[a][q] * dk[q * D + b];
          Lik[a][b] = v;
        }
      store_block<D>(row_i, deg, s, Lik);
      const int bk0 = m.brptr[k], degk = m.brptr[k + 1] - bk0;
      const int* col_k = m.bcol + bk0;
      const double* row_k = LU + (size_t)D * D * bk0;
      for (int t2 = 0; t2 < degk; ++t2) {
        const int j = col_k[t2];
        if (color[j] <= cc) continue;  // L part of row k, or outside the solved range (-1)
        int sj = -1;
        for (int q = 0; q < deg; ++q)
          if (col_i[q] == j) {
            sj = q;
            break;
          }
        if (sj < 0) continue;  // ILU(0): fill outside the pattern is dropped
        double U[D][D], A[D][D];
        load_block<D>(row_k, degk, t2, U);
        load_block<D>(row_i, deg, sj, A);
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
          for (int b = 0; b < D; ++b) {
            double v = A[a][b];
#pragma unroll
            for (int q = 0; q < D; ++q) v -= Lik[a][q] * U[q][b];
            A[a][b] = v;
          }
        store_block<D>(row_i, deg, sj, A);
      }
    }
  }
  int self = 0;
  for (int q = 0; q < deg; ++q)
    if (col_i[q] == i) self = q;
  double A[D][D];
  load_block<D>(row_i, deg, self, A);
  invert_block<D>(A, Dinv + (size_t)D * D * i);
}

// One triangular sweep over the rows of colour c, LANES lanes per block row.
//   FWD:  z_i -= sum over blocks (i, k) with 0 <= colour(k) < c of L_ik z_k
//   !FWD: z_i  = U_ii^-1 (z_i - sum over blocks (i, j) with colour(j) > c of U_ij z_j)
template <int D, int LANES, bool FWD>
__global__ void ilu_sweep_color_kernel(MeshView m, const double* __restrict__ LU,
                                       const double* __restrict__ Dinv,
                                       const int* __restrict__ color,
                                       const int* __restrict__ rows, int count, int c,
                                       double* __restrict__ z) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  const int lane = threadIdx.x % LANES;
  const bool valid = t < count;
  const int i = valid ? rows[t] : 0;
  double acc[D];
#pragma unroll
  for (int a = 0; a < D; ++a) acc[a] = 0.0;
  if (valid) {
    const int b0 = m.brptr[i], deg = m.brptr[i + 1] - b0;
    const double* row_i = LU + (size_t)D * D * b0;
    for (int s = lane; s < deg; s += LANES) {
      const int k = m.bcol[b0 + s];
      const int ck = color[k];
      if (FWD ? (ck < 0 || ck >= c) : (ck <= c)) continue;
      double zk[D];
#pragma unroll
      for (int q = 0; q < D; ++q) zk[q] = z[D * k + q];
#pragma unroll
      for (int a = 0; a < D; ++a)
#pragma unroll
        for (int q = 0; q < D; ++q) acc[a] += row_i[a * D * deg + s * D + q] * zk[q];
    }
  }
#pragma unroll
  for (int off = LANES / 2; off > 0; off >>= 1)
#pragma unroll
    for (int a = 0; a < D; ++a) acc[a] += __shfl_down_sync(0xffffffffu, acc[a], off, LANES);
  if (valid && lane == 0) {
    double v[D];
#pragma unroll
    for (int a = 0; a < D; ++a) v[a] = z[D * i + a] - acc[a];
    if (FWD) {
#pragma unroll
      for (int a = 0; a < D; ++a) z[D * i + a] = v[a];
    } else {
      const double* o = Dinv + (size_t)D * D * i;
#pragma unroll
      for (int a = 0; a < D; ++a) {
        double w = 0.0;
#pragma unroll
        for (int q = 0; q < D; ++q) w += o[a * D + q] * v[q];
        z[D * i + a] = w;
      }
    }
  }
}

}  // namespace

void ilu_release(vf_engine* e) {
  if (e->ilu.mem) cudaFree(e->ilu.mem);
  e->ilu = IluState{};
}

}  // namespace vf

using namespace vf;

extern "C" {

int vf_ilu_setup(vf_engine* e, int node0, int node1, int ncolors, const int32_t* rows_host,
                 const int32_t* color_ptr_host, const int32_t* color_host, void* stream) {
  if (!e) return fail("null engine");
  const int nn = e->desc.nn;
  if (node0 < 0 || node1 > nn || node0 >= node1) return fail("vf_ilu_setup: node range out of bounds");
  if (ncolors <= 0 || ncolors > 64 || !rows_host || !color_ptr_host || !color_host)
    return fail("vf_ilu_setup: bad colouring");
  if (color_ptr_host[0] != 0 || color_ptr_host[ncolors] != node1 - node0)
    return fail("vf_ilu_setup: colour classes do not cover the node range");
  ilu_release(e);
  cudaStream_t st = as_stream(stream);
  const int d = e->desc.dim;
  const size_t nnzb = (size_t)e->brptr[nn];
  const size_t b_lu = align_up(sizeof(double) * d * d * nnzb, 256);
  const size_t b_dinv = align_up(sizeof(double) * d * d * (size_t)nn, 256);
  const size_t b_color = align_up(sizeof(int) * (size_t)nn, 256);
  const size_t b_rows = align_up(sizeof(int) * (size_t)(node1 - node0), 256);
  char* mem = nullptr;
  VF_CUDA(cudaMalloc(&mem, b_lu + b_dinv + b_color + b_rows));
  IluState& S = e->ilu;
  S.mem = mem;
  S.LU = reinterpret_cast<double*>(mem);
  S.Dinv = reinterpret_cast<double*>(mem + b_lu);
  S.color = reinterpret_cast<int*>(mem + b_lu + b_dinv);
  S.rows = reinterpret_cast<int*>(mem + b_lu + b_dinv + b_color);
  S.node0 = node0;
  S.node1 = node1;
  S.ncolors = ncolors;
  S.color_ptr.assign(color_ptr_host, color_ptr_host + ncolors + 1);
  VF_CUDA(cudaMemcpyAsync(S.color, color_host, sizeof(int) * (size_t)nn, cudaMemcpyHostToDevice, st));
  VF_CUDA(cudaMemcpyAsync(S.rows, rows_host, sizeof(int) * (size_t)(node1 - node0),
                          cudaMemcpyHostToDevice, st));
  VF_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int vf_ilu_factor(vf_engine* e, int member, void* stream) {
  if (!e) return fail("null engine");
  if (member < 0 || member >= e->desc.n_members) return fail("member out of range");
  IluState& S = e->ilu;
  if (!S.mem) return fail("vf_ilu_factor: call vf_ilu_setup first");
  cudaStream_t st = as_stream(stream);
  const int d = e->desc.dim;
  const size_t nvals = (size_t)d * d * (size_t)e->brptr[e->desc.nn];
  VF_CUDA(cudaMemcpyAsync(S.LU, member_array(e, VF_J, member), sizeof(double) * nvals,
                          cudaMemcpyDeviceToDevice, st));
  for (int c = 0; c < S.ncolors; ++c) {
    const int count = S.color_ptr[c + 1] - S.color_ptr[c];
    if (count <= 0) continue;
    const int block = 128, grid = (count + block - 1) / block;
    if (d == 2)
      ilu_factor_color_kernel<2><<<grid, block, 0, st>>>(e->dev.mesh, S.LU, S.Dinv, S.color,
                                                          S.rows + S.color_ptr[c], count, c);
    else
      ilu_factor_color_kernel<3><<<grid, block, 0, st>>>(e->dev.mesh, S.LU, S.Dinv, S.color,
                                                          S.rows + S.color_ptr[c], count, c);
    e->launches += 1;
  }
  VF_CUDA(cudaGetLastError());
  return 0;
}

int vf_ilu_apply(vf_engine* e, const double* r_dev, double* z_dev, void* stream) {
  if (!e) return fail("null engine");
  IluState& S = e->ilu;
  if (!S.mem) return fail("vf_ilu_apply: call vf_ilu_setup first");
  if (!r_dev || !z_dev) return fail("vf_ilu_apply: null vector");
  cudaStream_t st = as_stream(stream);
  const int d = e->desc.dim;
  const size_t n0 = (size_t)d * S.node0, n = (size_t)d * (S.node1 - S.node0);
  if (r_dev != z_dev)
    VF_CUDA(cudaMemcpyAsync(z_dev + n0, r_dev + n0, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
  constexpr int kLanes2 = 4, kLanes3 = 8;
  const int block = 256;
  auto launch = [&](int c, bool fwd) {
    const int count = S.color_ptr[c + 1] - S.color_ptr[c];
    if (count <= 0) return;
    const int* rows = S.rows + S.color_ptr[c];
    if (d == 2) {
      const int grid = (int)(((size_t)count * kLanes2 + block - 1) / block);
      if (fwd)
        ilu_sweep_color_kernel<2, kLanes2, true><<<grid, block, 0, st>>>(
            e->dev.mesh, S.LU, S.Dinv, S.color, rows, count, c, z_dev);
      else
        ilu_sweep_color_kernel<2, kLanes2, false><<<grid, block, 0, st>>>(
            e->dev.mesh, S.LU, S.Dinv, S.color, rows, count, c, z_dev);
    } else {
      const int grid = (int)(((size_t)count * kLanes3 + block - 1) / block);
      if (fwd)
        ilu_sweep_color_kernel<3, kLanes3, true><<<grid, block, 0, st>>>(
            e->dev.mesh, S.LU, S.Dinv, S.color, rows, count, c, z_dev);
      else
        ilu_sweep_color_kernel<3, kLanes3, false><<<grid, block, 0, st>>>(
            e->dev.mesh, S.LU, S.Dinv, S.color, rows, count, c, z_dev);
    }
    e->launches += 1;
  };
  for (int c = 1; c < S.ncolors; ++c) launch(c, true);
  for (int c = S.ncolors - 1; c >= 0; --c) launch(c, false);
  VF_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
