// Grid-wide Krylov building blocks (single large mesh, optionally one partition of it) and
// their C-ABI entry points: CSR SpMV, block-Jacobi, fused dot products / updates.
#include "engine_internal.h"
#include "dev_util.cuh"

namespace vf {

// y = J x.  L lanes cooperate on one node block row (d scalar rows share their columns).
template <int D, int LANES>
__global__ void spmv_kernel(MeshView m, const double* __restrict__ J,
                            const double* __restrict__ x, double* __restrict__ y, int node0,
                            int node1, size_t pf_bytes) {
  const int gt = blockIdx.x * blockDim.x + threadIdx.x;
  const int node = node0 + gt / LANES;
  const int lane = gt % LANES;
  const bool valid = node < node1;
  double acc[D];
#pragma unroll
  for (int a = 0; a < D; ++a) acc[a] = 0.0;
  if (valid) {
    const int b0 = m.brptr[node], deg = m.brptr[node + 1] - b0;
    const double* blk = J + (size_t)D * D * b0;
    if (pf_bytes > 0) {
      // J, bcol and brptr are contiguous streams consumed in block order: every thread asks L2
      // for the line a fixed distance ahead of the one it is about to read, so the union of the
      // requests is the stream itself, shifted -- later CTAs then find their three dependent
      // loads (brptr -> bcol -> values) in L2 instead of paying three DRAM round trips
      const size_t vend = (size_t)D * D * m.brptr[m.nn] * sizeof(double);
      const size_t voff = (size_t)((const char*)(blk + D * lane) - (const char*)J) + pf_bytes;
      if (lane < deg) {
#pragma unroll
        for (int a = 0; a < D; ++a) {   // one request per scalar row of the block row
          const size_t o = voff + (size_t)a * D * deg * sizeof(double);
          if (o < vend) prefetch_l2((const char*)J + o);
        }
      }
      if (lane == 0) {
        const size_t ahead = pf_bytes / (D * D * sizeof(double));   // blocks
        if ((size_t)b0 + ahead < (size_t)m.brptr[m.nn]) prefetch_l2(m.bcol + b0 + ahead);
        const size_t nahead = ahead / 7;                             // nodes (7 blocks per row)
        if ((size_t)node + nahead < (size_t)m.nn) prefetch_l2(m.brptr + node + nahead);
      }
    }
    for (int k = lane; k < deg; k += LANES) {
      const int j = __ldg(m.bcol + b0 + k);
      if (D == 2) {
        const double2 xv = *reinterpret_cast<const double2*>(x + 2 * j);
        const double2 r0 = __ldcs(reinterpret_cast<const double2*>(blk + 2 * k));
        const double2 r1 = __ldcs(reinterpret_cast<const double2*>(blk + 2 * deg + 2 * k));
        acc[0] += r0.x * xv.x + r0.y * xv.y;
        acc[1] += r1.x * xv.x + r1.y * xv.y;
      } else {
        double xv[D];
#pragma unroll
        for (int c = 0; c < D; ++c) xv[c] = x[D * j + c];
#pragma unroll
        for (int a = 0; a < D; ++a) {
          const double* row = blk + (size_t)a * D * deg + k * D;
#pragma unroll
          for (int c = 0; c < D; ++c) acc[a] += __ldcs(row + c) * xv[c];
        }
      }
    }
  }
#pragma unroll
  for (int off = LANES / 2; off > 0; off >>= 1)
#pragma unroll
    for (int a = 0; a < D; ++a) acc[a] += __shfl_down_sync(0xffffffffu, acc[a], off, LANES);
  if (valid && lane == 0) {
#pragma unroll
    for (int a = 0; a < D; ++a) y[D * node + a] = acc[a];
  }
}


// ---- grid-wide Krylov building blocks (single large mesh, optionally one partition of it) ----

// Block-Jacobi inverse of the d x d diagonal blocks for node rows [node0, node1)
template <int D>
__global__ void block_jacobi_kernel(MeshView m, const double* __restrict__ J,
                                    double* __restrict__ Dinv, int node0, int node1) {
  const int i = node0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= node1) return;
  const int b0 = m.brptr[i], deg = m.brptr[i + 1] - b0;
  const int self = find_slot(m.bcol + b0, deg, i);
  const double* blk = J + (size_t)D * D * b0;
  double A[D][D];
  for (int a = 0; a < D; ++a)
    for (int c = 0; c < D; ++c) A[a][c] = blk[a * D * deg + self * D + c];
  double* o = Dinv + (size_t)D * D * i;
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

// z = Dinv r on DOFs of node rows [node0, node1)
template <int D>
__global__ void apply_block_jacobi_kernel(const double* __restrict__ Dinv,
                                          const double* __restrict__ r, double* __restrict__ z,
                                          int node0, int node1) {
  const int i = node0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= node1) return;
  const double* o = Dinv + (size_t)D * D * i;
  double v[D];
#pragma unroll
  for (int c = 0; c < D; ++c) v[c] = r[D * i + c];
#pragma unroll
  for (int a = 0; a < D; ++a) {
    double t = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) t += o[a * D + c] * v[c];
    z[D * i + a] = t;
  }
}
constexpr int kDotBlock = 256;
// Partial dot products of w with NV (<= nvec) Krylov vectors in ONE pass over the data: each
// thread keeps the NV accumulators of its elements in registers (w is read once, every V_j
// once, coalesced), then the block reduces them in a fixed order (deterministic).
template <int NV>
__global__ void __launch_bounds__(kDotBlock) multidot_partial_kernel(
    const double* __restrict__ V, size_t ldv, int nvec, const double* __restrict__ w, size_t n,
    double* __restrict__ partial) {
  __shared__ double red[kDotBlock / 32][NV];
  double acc[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) acc[j] = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const double wi = w[i];
#pragma unroll
    for (int j = 0; j < NV; ++j)
      if (j < nvec) acc[j] += V[(size_t)j * ldv + i] * wi;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const double s = warp_sum(acc[j]);
    if (lane == 0) red[wid][j] = s;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nvec; j += blockDim.x) {
    double t = 0.0;
    for (int q = 0; q < kDotBlock / 32; ++q) t += red[q][j];
    partial[(size_t)blockIdx.x * nvec + j] = t;
  }
}

__global__ void multidot_final_kernel(const double* __restrict__ partial, int nblocks, int nvec,
                                      double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nvec) return;
  double t = 0.0;
  for (int b = 0; b < nblocks; ++b) t += partial[(size_t)b * nvec + j];
  out[j] = t;
}

// w -= sum_j h[j] V_j   (h on the device)
__global__ void multi_axpy_kernel(const double* __restrict__ V, size_t ldv, int nvec,
                                  const double* __restrict__ h, double* __restrict__ w, size_t n) {
  extern __shared__ double hs[];
  for (int j = threadIdx.x; j < nvec; j += blockDim.x) hs[j] = h[j];
  __syncthreads();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int j = 0; j < nvec; ++j) s += hs[j] * V[(size_t)j * ldv + i];
    w[i] -= s;
  }
}

// y = x / sqrt(s) with s = *s2 - sum_i sub[i]^2 read from device memory (y = 0 when s <= 0):
// normalises a Krylov vector by a norm that never visits the host.  The subtraction is the
// Pythagorean update of the second Gram-Schmidt pass; s is also stored to *s_out for the host.
__global__ void scale_rsqrt_kernel(const double* x, const double* __restrict__ s2,
                                   const double* __restrict__ sub, int nsub, double* s_out,
                                   double* y, size_t n) {  // y may alias x
  double v = *s2;
  for (int i = 0; i < nsub; ++i) v -= sub[i] * sub[i];
  const double f = v > 0.0 ? 1.0 / sqrt(v) : 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    y[i] = f * x[i];
  if (s_out && blockIdx.x == 0 && threadIdx.x == 0) *s_out = v;
}

// y = alpha x + beta y
__global__ void axpby_kernel(double alpha, const double* __restrict__ x, double beta,
                             double* __restrict__ y, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    y[i] = alpha * x[i] + (beta == 0.0 ? 0.0 : beta * y[i]);
}

}  // namespace vf

using namespace vf;

extern "C" {

int vf_spmv_rows(vf_engine* e, int member, const double* x_dev, double* y_dev, int node0,
                 int node1, void* stream) {
  if (!e) return fail("null engine");
  if (member < 0 || member >= e->desc.n_members) return fail("member out of range");
  if (node0 < 0 || node1 > e->desc.nn || node0 > node1) return fail("node range out of bounds");
  if (node0 == node1) return 0;
  cudaStream_t st = as_stream(stream);
  const double* J = member_array(e, VF_J, member);
  const int nrows = node1 - node0;
  const int block = 256;
  // L2 prefetch distance of the value stream (VF_SPMV_PF_MB; 0 disables).  Measured on the
  // 5.6e7-nnz matrix: 0 -> 0.1536, 4 MB -> 0.1471, 16 MB -> 0.1495, 64 MB -> 0.1707 ms; only
  // worth it when the matrix does not sit in L2 anyway
  static const char* env_pf = getenv("VF_SPMV_PF_MB");
  const size_t pf_mb = env_pf ? (size_t)std::max(atoi(env_pf), 0) : 4;
  const size_t jbytes = (size_t)e->dev.nnz * sizeof(double);
  const size_t pf_bytes = jbytes > ((size_t)64 << 20) ? pf_mb << 20 : 0;
  if (e->desc.dim == 2) {
    // lanes per node block row (VF_SPMV_LANES).  Triangles have ~7 blocks per row; measured on
    // the 5.6e7-nnz matrix: 1 -> 0.405, 2 -> 0.195, 4 -> 0.1345, 8 -> 0.1476, 16 -> 0.270 ms
    static const char* env_ln = getenv("VF_SPMV_LANES");
    const int ln = env_ln ? atoi(env_ln) : 4;
    if (ln == 2) {
      const int grid = (int)(((size_t)nrows * 2 + block - 1) / block);
      spmv_kernel<2, 2><<<grid, block, 0, st>>>(e->dev.mesh, J, x_dev, y_dev, node0, node1, pf_bytes);
    } else if (ln == 1) {
      const int grid = (int)(((size_t)nrows + block - 1) / block);
      spmv_kernel<2, 1><<<grid, block, 0, st>>>(e->dev.mesh, J, x_dev, y_dev, node0, node1, pf_bytes);
    } else if (ln == 4) {
      const int grid = (int)(((size_t)nrows * 4 + block - 1) / block);
      spmv_kernel<2, 4><<<grid, block, 0, st>>>(e->dev.mesh, J, x_dev, y_dev, node0, node1, pf_bytes);
    } else if (ln == 16) {
      const int grid = (int)(((size_t)nrows * 16 + block - 1) / block);
      spmv_kernel<2, 16><<<grid, block, 0, st>>>(e->dev.mesh, J, x_dev, y_dev, node0, node1, pf_bytes);
    } else {
      constexpr int LN = 8;
      const int grid = (int)(((size_t)nrows * LN + block - 1) / block);
      spmv_kernel<2, LN><<<grid, block, 0, st>>>(e->dev.mesh, J, x_dev, y_dev, node0, node1, pf_bytes);
    }
  } else {
    // tetrahedra: ~15 blocks per row (VF_SPMV_LANES3)
    static const char* env_l3 = getenv("VF_SPMV_LANES3");
    const int l3 = env_l3 ? atoi(env_l3) : 16;
    if (l3 == 8) {
      const int grid = (int)(((size_t)nrows * 8 + block - 1) / block);
      spmv_kernel<3, 8><<<grid, block, 0, st>>>(e->dev.mesh, J, x_dev, y_dev, node0, node1, pf_bytes);
    } else if (l3 == 4) {
      const int grid = (int)(((size_t)nrows * 4 + block - 1) / block);
      spmv_kernel<3, 4><<<grid, block, 0, st>>>(e->dev.mesh, J, x_dev, y_dev, node0, node1, pf_bytes);
    } else {
      constexpr int LN = 16;
      const int grid = (int)(((size_t)nrows * LN + block - 1) / block);
      spmv_kernel<3, LN><<<grid, block, 0, st>>>(e->dev.mesh, J, x_dev, y_dev, node0, node1, pf_bytes);
    }
  }
  e->launches += 1;
  VF_CUDA(cudaGetLastError());
  return 0;
}

int vf_spmv(vf_engine* e, int member, const double* x_dev, double* y_dev, void* stream) {
  if (!e) return fail("null engine");
  return vf_spmv_rows(e, member, x_dev, y_dev, 0, e->desc.nn, stream);
}

int vf_block_jacobi_setup(vf_engine* e, int member, int node0, int node1, void* stream) {
  if (!e) return fail("null engine");
  if (member < 0 || member >= e->desc.n_members) return fail("member out of range");
  if (node0 < 0 || node1 > e->desc.nn || node0 >= node1) return fail("node range out of bounds");
  cudaStream_t st = as_stream(stream);
  const double* J = member_array(e, VF_J, member);
  double* Dinv = e->dev.members + (size_t)member * e->dev.L.stride + e->dev.L.Dinv;
  const int block = 128, grid = (node1 - node0 + block - 1) / block;
  if (e->desc.dim == 2)
    block_jacobi_kernel<2><<<grid, block, 0, st>>>(e->dev.mesh, J, Dinv, node0, node1);
  else
    block_jacobi_kernel<3><<<grid, block, 0, st>>>(e->dev.mesh, J, Dinv, node0, node1);
  e->launches += 1;
  VF_CUDA(cudaGetLastError());
  return 0;
}

int vf_block_jacobi_apply(vf_engine* e, int member, const double* r_dev, double* z_dev, int node0,
                          int node1, void* stream) {
  if (!e) return fail("null engine");
  if (member < 0 || member >= e->desc.n_members) return fail("member out of range");
  if (node0 < 0 || node1 > e->desc.nn || node0 >= node1) return fail("node range out of bounds");
  cudaStream_t st = as_stream(stream);
  const double* Dinv = e->dev.members + (size_t)member * e->dev.L.stride + e->dev.L.Dinv;
  const int block = 256, grid = (node1 - node0 + block - 1) / block;
  if (e->desc.dim == 2)
    apply_block_jacobi_kernel<2><<<grid, block, 0, st>>>(Dinv, r_dev, z_dev, node0, node1);
  else
    apply_block_jacobi_kernel<3><<<grid, block, 0, st>>>(Dinv, r_dev, z_dev, node0, node1);
  e->launches += 1;
  VF_CUDA(cudaGetLastError());
  return 0;
}

int vf_multidot(vf_engine* e, const double* V_dev, size_t ldv, int nvec, const double* w_dev,
                size_t n, double* out_dev, double* scratch_dev, size_t scratch_count,
                void* stream) {
  if (!e) return fail("null engine");
  if (nvec <= 0 || n == 0) return fail("vf_multidot: empty problem");
  cudaStream_t st = as_stream(stream);
  int nblocks = (int)std::min<size_t>(148 * 4, (n + 2047) / 2048);
  nblocks = std::max(nblocks, 1);
  if (scratch_count < (size_t)nblocks * nvec) return fail("vf_multidot: scratch too small");
  // vectors are processed in groups of at most 32 (register accumulators)
  for (int j0 = 0; j0 < nvec; j0 += 32) {
    const int nv = std::min(32, nvec - j0);
    const double* Vg = V_dev + (size_t)j0 * ldv;
    double* part = scratch_dev + (size_t)j0 * nblocks;
    if (nv <= 4)
      multidot_partial_kernel<4><<<nblocks, kDotBlock, 0, st>>>(Vg, ldv, nv, w_dev, n, part);
    else if (nv <= 8)
      multidot_partial_kernel<8><<<nblocks, kDotBlock, 0, st>>>(Vg, ldv, nv, w_dev, n, part);
    else if (nv <= 16)
      multidot_partial_kernel<16><<<nblocks, kDotBlock, 0, st>>>(Vg, ldv, nv, w_dev, n, part);
    else
      multidot_partial_kernel<32><<<nblocks, kDotBlock, 0, st>>>(Vg, ldv, nv, w_dev, n, part);
    multidot_final_kernel<<<(nv + 63) / 64, 64, 0, st>>>(part, nblocks, nv, out_dev + j0);
    e->launches += 2;
  }
  VF_CUDA(cudaGetLastError());
  return 0;
}

int vf_multi_axpy(vf_engine* e, const double* V_dev, size_t ldv, int nvec, const double* h_dev,
                  double* w_dev, size_t n, void* stream) {
  if (!e) return fail("null engine");
  if (nvec <= 0 || n == 0) return 0;
  cudaStream_t st = as_stream(stream);
  const int block = 256;
  const int grid = (int)std::min<size_t>(148 * 8, (n + block - 1) / block);
  multi_axpy_kernel<<<grid, block, sizeof(double) * nvec, st>>>(V_dev, ldv, nvec, h_dev, w_dev, n);
  e->launches += 1;
  VF_CUDA(cudaGetLastError());
  return 0;
}

int vf_axpby(vf_engine* e, double alpha, const double* x_dev, double beta, double* y_dev, size_t n,
             void* stream) {
  if (!e) return fail("null engine");
  if (n == 0) return 0;
  cudaStream_t st = as_stream(stream);
  const int block = 256;
  const int grid = (int)std::min<size_t>(148 * 8, (n + block - 1) / block);
  axpby_kernel<<<grid, block, 0, st>>>(alpha, x_dev, beta, y_dev, n);
  e->launches += 1;
  VF_CUDA(cudaGetLastError());
  return 0;
}

int vf_scale_rsqrt(vf_engine* e, const double* x_dev, const double* s2_dev, const double* sub_dev,
                   int nsub, double* s_out_dev, double* y_dev, size_t n, void* stream) {
  if (!e) return fail("null engine");
  if (n == 0) return 0;
  if (nsub < 0 || (nsub > 0 && !sub_dev)) return fail("vf_scale_rsqrt: bad subtraction list");
  if (s_out_dev == s2_dev) return fail("vf_scale_rsqrt: s_out must not alias s2");
  const int block = 256;
  const int grid = (int)std::min<size_t>(148 * 8, (n + block - 1) / block);
  scale_rsqrt_kernel<<<grid, block, 0, as_stream(stream)>>>(x_dev, s2_dev, sub_dev, nsub,
                                                             s_out_dev, y_dev, n);
  e->launches += 1;
  VF_CUDA(cudaGetLastError());
  return 0;
}


}  // extern "C"
