// vffem_b200: CUDA kernels (sm_100a) and the C ABI declared in include/vffem_b200.h.
//
// Kernels
//   asm_tile_kernel   K1+K2+K3: residual + Jacobian of one contiguous node tile; the tile's
//                     slice of the CSR value array is accumulated in shared memory by the
//                     owning threads and streamed out once with coalesced vector stores
//   spmv_kernel       K4: block-aware CSR SpMV, L lanes per node block row
//   fluid_kernel      K8: Bernoulli channels, one warp each
//   member_kernel     K5-K8: persistent Newton/GMRES/Newmark/FSI time loop, one CTA per member
// See DESIGN.md for the layout and the roofline of each.

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/vffem_b200.h"
#include "member_solver.cuh"

namespace vf {

static thread_local std::string g_err;

static int fail(const std::string& msg) {
  g_err = msg;
  return 1;
}

#define VF_CUDA(call)                                                            \
  do {                                                                           \
    cudaError_t _e = (call);                                                     \
    if (_e != cudaSuccess)                                                       \
      return fail(std::string(#call) + ": " + cudaGetErrorString(_e));           \
  } while (0)

// ---- grid-wide kernels (single large mesh) ----------------------------------------------

template <int D, bool JAC, bool RES>
__global__ void asm_tile_kernel(EngineDev E, int member, double dt, int is_static, JacMix mix,
                                const int* __restrict__ tile_start) {
  extern __shared__ double tile[];
  double* mb = E.members + (size_t)member * E.L.stride;
  const Layout& L = E.L;
  const int i0 = tile_start[blockIdx.x], i1 = tile_start[blockIdx.x + 1];
  const size_t base = (size_t)D * D * E.mesh.brptr[i0];
  const int nvals = int((size_t)D * D * E.mesh.brptr[i1] - base);
  const int i = i0 + threadIdx.x;
  if (i < i1) {
    PropView pv = member_props<D>(E, mb);
    StateView sv;
    sv.u1 = mb + L.off[VF_U1];
    sv.u0 = is_static ? sv.u1 : mb + L.off[VF_U0];
    sv.v0 = mb + L.off[VF_V0];
    sv.a0 = mb + L.off[VF_A0];
    sv.p1 = mb + L.off[VF_P1];
    sv.dt = dt;
    sv.is_static = is_static;
  sv.mix = mix;
    sv.mix = mix;
    double res[D];
    double* rowblk = JAC ? tile + ((size_t)D * D * E.mesh.brptr[i] - base) : nullptr;
    assemble_node<D, JAC, RES>(i, E.mesh, pv, sv, rowblk, res);
    if (RES) {
      double* F = mb + L.off[VF_F];
#pragma unroll
      for (int c = 0; c < D; ++c) F[D * i + c] = res[c];
    }
  }
  if (JAC) {
    __syncthreads();
    double* Jg = mb + L.off[VF_J] + base;
    if (D == 2) {
      // base and nvals are multiples of 4 doubles: 16-byte vector stores, fully coalesced
      double2* dst = reinterpret_cast<double2*>(Jg);
      const double2* src = reinterpret_cast<const double2*>(tile);
      for (int t = threadIdx.x; t < nvals / 2; t += blockDim.x) dst[t] = src[t];
    } else {
      for (int t = threadIdx.x; t < nvals; t += blockDim.x) Jg[t] = tile[t];
    }
  }
}



// Two-phase, element-centric tile assembly (triangles).  Replaces the thread-per-node gather
// for 2D: every cell touching the tile is processed ONCE per CTA.
//   phase 0  one 32-byte tile descriptor, then all index data of the tile (vertex quads,
//            packed pair info, slices of brptr / n2e_ptr) is fetched with independent,
//            coalesced loads -- a single dependent round trip instead of the chain
//            tile_start -> te_ptr -> te_elem -> cells -> nodal data
//   phase 1  thread per cell: 16-byte nodal gathers, geometry + material + cell residual
//            -> one 144-byte record in shared memory
//   phase 2  thread per scalar row (or per node) of the tile: walks the row's (node, cell)
//            pairs in the fixed n2e order, reads the records, and accumulates its CSR row
//            slice in shared memory (rows are private to their thread: no atomics,
//            bit-reproducible, same summation order as assemble_node)
//   phase 3  the tile's CSR slice and residual entries are streamed out with coalesced stores
// The exterior-facet terms and Dirichlet rows touch O(sqrt(N)) boundary nodes only and are
// applied afterwards by facet_bc_kernel, which keeps this kernel's register budget small.
// Shared memory: [records: max_tile_elems x 18][CSR slice][F: 2 x nodes][pair info][brptr][n2e_ptr].
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ int4 ldg_nc_v4(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ double ldg_nc_f64(const double* p) {
  double r;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}

template <bool JAC, bool RES, int ROW, int MAXT, int MINB, bool DIRECT = false>
__global__ void __launch_bounds__(MAXT, MINB) asm_tile2_kernel(
    EngineDev E, int member, NewmarkCoef nc_arg, int is_static, JacMix mix,
    const int4* __restrict__ tile_desc,
    const int4* __restrict__ te_quad, const unsigned* __restrict__ pair_info,
    const int* __restrict__ tile_halo, int max_tile_elems, int tile_max_values,
    int max_tile_pairs, int max_tile_nodes, int max_tile_verts, int pf_dist, int dbg_skip) {
  constexpr int D = 2;
  constexpr int kThreads = MAXT;  // always launched with exactly MAXT threads: loop strides
                                  // and trip counts are compile-time constants
  extern __shared__ double smem[];
  double* recs = smem;
  // region A: the CSR slice (phases 2-3) aliases the nodal staging area (phases 0-1)
  double* tileJ = recs + (size_t)max_tile_elems * kRec2D;
  D2* s_xy = reinterpret_cast<D2*>(tileJ);
  NodeUVA* s_uva = reinterpret_cast<NodeUVA*>(s_xy + max_tile_verts);
  // DIRECT: phase 2 stores its rows straight to HBM, the region only holds the staging area
  const size_t region_a = DIRECT ? (size_t)8 * max_tile_verts
                                 : max((size_t)tile_max_values, (size_t)8 * max_tile_verts);
  double* tileF = tileJ + region_a;
  unsigned* s_pair = reinterpret_cast<unsigned*>(tileF + D * max_tile_nodes);
  int* s_brptr = reinterpret_cast<int*>(s_pair + max_tile_pairs);
  int* s_n2e = s_brptr + max_tile_nodes + 1;
  double* mb = E.members + (size_t)member * E.L.stride;
  const Layout& L = E.L;
  const MeshView& m = E.mesh;

#ifdef VF_PHASE_PROF  // -DVF_PHASE_PROF + VF_DEBUG_SKIP=32: per-phase cycle counts into info[8..15]
  const bool prof = (dbg_skip & 32) && threadIdx.x == 0;
  long long tk0 = prof ? clock64() : 0, tk1 = 0, tk2 = 0, tk3 = 0, ta = 0, tb = 0, tc = 0;
#define VF_PROBE(x) x
#else
#define VF_PROBE(x)
#endif
  // ---- phase 0: descriptor, then every load of the tile in one dependent round ---------------
  const int4 d0 = tile_desc[3 * blockIdx.x], d1 = tile_desc[3 * blockIdx.x + 1];
  const int i0 = d0.x, te0 = d0.y, pr0 = d0.z, bbase = d0.w;
  const int h0 = d1.x, nT = d1.y & 0xffff, nH = (int)((unsigned)d1.y >> 16);
  const int nte = d1.z & 0xffff, npr = (int)((unsigned)d1.z >> 16);
  const int nV = nT + nH;
  const size_t base = (size_t)D * D * bbase;
  const int nvals = D * D * d1.w;
  const PropView pv = member_props<D>(E, mb);
  const double* u1 = mb + L.off[VF_U1];
  const double* u0 = is_static ? u1 : mb + L.off[VF_U0];
  const double* v0 = mb + L.off[VF_V0];
  const double* a0 = mb + L.off[VF_A0];

  VF_PROBE(if (prof) ta = clock64() + (i0 & 0);)
  // this thread's first cell (volatile load: issued here, not sunk below the barrier)
  int4 quad = make_int4(0, 0, 0, 0);
  const bool have = (int)threadIdx.x < nte;
  if (have) quad = ldg_nc_v4(te_quad + te0 + threadIdx.x);
  // index slices: asynchronous global->shared copies, no registers, waited for at the barrier
  for (int t = threadIdx.x; t < npr; t += kThreads) cp_async4(s_pair + t, pair_info + pr0 + t);
  for (int t = threadIdx.x; t <= nT; t += kThreads) {
    cp_async4(s_brptr + t, m.brptr + i0 + t);
    cp_async4(s_n2e + t, m.n2e_ptr + i0 + t);
  }
  // descriptor of the tile `pf_dist` CTAs ahead: its inputs are pulled into L2 by this CTA's
  // idle threads during phase 2, so that the later CTA's three dependent round trips hit L2
  __shared__ int s_far[12];
  const int far = blockIdx.x + pf_dist;
  const bool pf = pf_dist > 0 && far < (int)gridDim.x;
  if (pf && threadIdx.x < 12)
    cp_async4(s_far + threadIdx.x, reinterpret_cast<const int*>(tile_desc) + 12 * far + threadIdx.x);
  cp_async_commit();
  // Lame / Newmark coefficients: a handful of fp64 divisions, done once per CTA
  __shared__ LameFac s_lf;
  if (threadIdx.x == kThreads - 1) s_lf = lame_fac(pv.scal[SC_NU]);
  // stage the tile's vertices -- its own contiguous range, then the halo vertices of its
  // cells -- with 16-byte loads; v_nmk / a_nmk are evaluated once per vertex here instead of
  // once per (cell, vertex) in phase 1, and phase 1 reads shared memory only
  for (int t = threadIdx.x; t < nV; t += kThreads) {
    const int vtx = t < nT ? i0 + t : tile_halo[h0 + t - nT];
    // all global loads first, then the shared-memory stores: a store in between would order
    // the (generic-pointer) loads behind it and cost a second round trip
    const D2 c2 = reinterpret_cast<const D2*>(m.xy)[vtx];
    NodeUVA s3;
    if (RES) s3 = gather_node_uva(nc_arg, is_static != 0, vtx, u1, u0, v0, a0);
    s_xy[t] = c2;
    if (RES) s_uva[t] = s3;
  }
  VF_PROBE(if (prof) tb = clock64();)
  // the quad has arrived by now: the cell's material data, also before the barrier
  double emod_e = 0.0, eta_e = 0.0, rho_e = 0.0;
  if (have) {
    emod_e = ldg_nc_f64(pv.emod + quad.w);
    eta_e = ldg_nc_f64(pv.eta + quad.w);
    rho_e = ldg_nc_f64(pv.rho + quad.w);
  }
  cp_async_wait_all();
  VF_PROBE(if (prof) tc = clock64() + (__double_as_longlong(emod_e) & 0);)
  __syncthreads();
  VF_PROBE(if (prof) tk1 = clock64();)

  // ---- phase 1: one record per cell, from shared memory ----------------------------------------
  {
    const LameFac lf = s_lf;
    const Damping dp = prop_damping(pv);
    for (int q = threadIdx.x; q < nte && !(dbg_skip & 1); q += kThreads) {
      if (q != (int)threadIdx.x) {
        quad = te_quad[te0 + q];
        emod_e = pv.emod[quad.w];
        eta_e = pv.eta[quad.w];
        rho_e = pv.rho[quad.w];
      }
      const int nd[3] = {quad.x, quad.y, quad.z};  // local slots in the staged vertex list
      double x[3][2];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const D2 c2 = s_xy[nd[a]];
        x[a][0] = c2.x;
        x[a][1] = c2.y;
      }
      tri_record_t(
          x, emod_e, lf, eta_e, rho_e, dp, mix, RES,
          [&](int a) { return s_uva[nd[a]]; }, recs + (size_t)q * kRec2D);
    }
  }
  __syncthreads();
  VF_PROBE(if (prof) tk2 = clock64();)

  // ---- phase 2 ---------------------------------------------------------------------------------
  if (dbg_skip & 2) {
    // measurement aid (VF_DEBUG_SKIP): phase skipped
  } else if (ROW == 2) {
    // one thread per scalar row; the row's cells are visited counter-clockwise around the
    // vertex (tables.order_fans_2d), so every off-diagonal block is the sum of two
    // CONSECUTIVE cells: it is completed in registers and stored once -- no zero-fill and no
    // read-modify-write of the shared-memory slice
    for (int r = threadIdx.x; r < D * nT; r += kThreads) {
      const int n = r >> 1, comp = r & 1;
      const int b0 = s_brptr[n], deg = s_brptr[n + 1] - b0;
      double* row = (DIRECT ? mb + L.off[VF_J] + base : tileJ) + D * D * (b0 - bbase) +
                    comp * D * deg;
      const int qb = s_n2e[n] - pr0, qe = s_n2e[n + 1] - pr0;
      double racc = 0.0;
      if (qe > qb) {
        // first cell of the fan (peeled): nothing to complete yet
        unsigned info = s_pair[qb];
        const double* rec = recs + (size_t)(info & 0xfffu) * kRec2D;
        int a = (info >> 12) & 3;
        D2 diag = D2{0.0, 0.0}, carry = D2{0.0, 0.0}, first = D2{0.0, 0.0};
        int slot_first = 0, slot_carry = 0;
        if (JAC) {
          D2 wn;
          tri_row_fan(rec, a, comp, diag, wn, carry);
          first = wn;
          slot_first = (info >> 20) & 63;
          slot_carry = (info >> 26) & 63;
        }
        if (RES) racc = rec[9 + 2 * a + comp];
        for (int q = qb + 1; q < qe; ++q) {
          info = s_pair[q];
          rec = recs + (size_t)(info & 0xfffu) * kRec2D;
          a = (info >> 12) & 3;
          if (JAC) {
            D2 ws, wn, wp;
            tri_row_fan(rec, a, comp, ws, wn, wp);
            diag.x += ws.x;
            diag.y += ws.y;
            *reinterpret_cast<D2*>(row + D * ((info >> 20) & 63)) =
                D2{carry.x + wn.x, carry.y + wn.y};
            carry = wp;
            slot_carry = (info >> 26) & 63;
          }
          if (RES) racc += rec[9 + 2 * a + comp];
        }
        if (JAC) {
          if (slot_carry == slot_first) {  // closed fan: the last cell meets the first
            *reinterpret_cast<D2*>(row + D * slot_first) = D2{first.x + carry.x, first.y + carry.y};
          } else {
            *reinterpret_cast<D2*>(row + D * slot_first) = first;
            *reinterpret_cast<D2*>(row + D * slot_carry) = carry;
          }
          *reinterpret_cast<D2*>(row + D * ((info >> 14) & 63)) = diag;
        }
      }
      if (RES) tileF[r] = racc;
    }
    if (pf) {
      // threads without a row (or all, when every thread has one) share the far tile's lines
      const int idle0 = ((D * nT + 31) / 32) * 32;
      const bool some_idle = idle0 + 32 <= (int)kThreads;
      const int k = some_idle ? (int)threadIdx.x - idle0 : (int)threadIdx.x;
      const int nk = some_idle ? (int)kThreads - idle0 : (int)kThreads;
      if (k >= 0) {
        const int f_i0 = s_far[0], f_te0 = s_far[1], f_pr0 = s_far[2], f_h0 = s_far[4];
        const int f_nT = s_far[5] & 0xffff, f_nH = (int)((unsigned)s_far[5] >> 16);
        const int f_nte = s_far[6] & 0xffff, f_npr = (int)((unsigned)s_far[6] >> 16);
        const int f_e0 = s_far[8], f_en = s_far[9];
        auto pull = [&](const void* p, int nbytes) {
          const char* c = reinterpret_cast<const char*>(p);
          for (int off = k * 128; off < nbytes; off += nk * 128) prefetch_l2(c + off);
        };
        pull(te_quad + f_te0, 16 * f_nte);
        pull(pair_info + f_pr0, 4 * f_npr);
        pull(tile_halo + f_h0, 4 * f_nH);
        pull(m.brptr + f_i0, 4 * (f_nT + 1));
        pull(m.n2e_ptr + f_i0, 4 * (f_nT + 1));
        pull(m.xy + D * f_i0, 16 * f_nT);
        if (RES) {
          pull(u1 + D * f_i0, 16 * f_nT);
          if (!is_static) {
            pull(u0 + D * f_i0, 16 * f_nT);
            pull(v0 + D * f_i0, 16 * f_nT);
            pull(a0 + D * f_i0, 16 * f_nT);
          }
        }
        if (!(dbg_skip & 64)) {
          // nodal data of the far tile's halo vertices (gathered: one line per vertex and array)
          for (int h = k; h < f_nH; h += nk) {
            const int v = tile_halo[f_h0 + h];
            prefetch_l2(m.xy + D * v);
            if (RES) {
              prefetch_l2(u1 + D * v);
              if (!is_static) {
                prefetch_l2(u0 + D * v);
                prefetch_l2(v0 + D * v);
                prefetch_l2(a0 + D * v);
              }
            }
          }
        }
        pull(pv.emod + f_e0, 8 * f_en);
        pull(pv.eta + f_e0, 8 * f_en);
        pull(pv.rho + f_e0, 8 * f_en);
      }
    }
  } else if (ROW == 1) {
    // one thread per scalar row, read-modify-write accumulation (any cell order)
    for (int r = threadIdx.x; r < D * nT; r += kThreads) {
      const int n = r >> 1, comp = r & 1;
      const int b0 = s_brptr[n], deg = s_brptr[n + 1] - b0;
      double* row = tileJ + D * D * (b0 - bbase) + comp * D * deg;
      if (JAC) {
        const D2 z = D2{0.0, 0.0};
        for (int t = 0; t < deg; ++t) reinterpret_cast<D2*>(row)[t] = z;
      }
      double racc = 0.0;
      const int qe = s_n2e[n + 1] - pr0;
      for (int q = s_n2e[n] - pr0; q < qe; ++q) {
        const unsigned info = s_pair[q];
        const double* rec = recs + (size_t)(info & 0xfffu) * kRec2D;
        const int a = (info >> 12) & 3;
        if (JAC) {
          D2 wv[3];
          tri_row_fan(rec, a, comp, wv[0], wv[1], wv[2]);  // slots are (self, next, prev)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const int slot = (info >> (14 + 6 * c)) & 63;
            D2* dst = reinterpret_cast<D2*>(row + D * slot);
            D2 cur = *dst;
            cur.x += wv[c].x;
            cur.y += wv[c].y;
            *dst = cur;
          }
        }
        if (RES) racc += rec[9 + 2 * a + comp];
      }
      if (RES) tileF[r] = racc;
    }
  } else {
    // one thread per node (both scalar rows of its block row)
    for (int n = threadIdx.x; n < nT; n += kThreads) {
      const int b0 = s_brptr[n], deg = s_brptr[n + 1] - b0;
      double* row0 = tileJ + D * D * (b0 - bbase);
      double* row1 = row0 + D * deg;
      if (JAC) {
        const D2 z = D2{0.0, 0.0};
        for (int t = 0; t < D * deg; ++t) reinterpret_cast<D2*>(row0)[t] = z;
      }
      double r0 = 0.0, r1 = 0.0;
      const int qe = s_n2e[n + 1] - pr0;
      for (int q = s_n2e[n] - pr0; q < qe; ++q) {
        const unsigned info = s_pair[q];
        const double* rec = recs + (size_t)(info & 0xfffu) * kRec2D;
        const int a = (info >> 12) & 3;
        if (JAC) {
#pragma unroll
          for (int sft = 0; sft < 3; ++sft) {
            const int slot = (info >> (14 + 6 * sft)) & 63;  // slots are (self, next, prev)
            const int c = (a + sft) % 3;
            double b[2][2];
            tri_block(rec, a, c, b);
            D2* p0 = reinterpret_cast<D2*>(row0 + D * slot);
            D2* p1 = reinterpret_cast<D2*>(row1 + D * slot);
            D2 c0 = *p0, c1 = *p1;
            c0.x += b[0][0];
            c0.y += b[0][1];
            c1.x += b[1][0];
            c1.y += b[1][1];
            *p0 = c0;
            *p1 = c1;
          }
        }
        if (RES) {
          r0 += rec[9 + 2 * a];
          r1 += rec[10 + 2 * a];
        }
      }
      if (RES) {
        tileF[D * n] = r0;
        tileF[D * n + 1] = r1;
      }
    }
  }
  __syncthreads();
  VF_PROBE(if (prof) tk3 = clock64();)

  // ---- phase 3: coalesced write-out ---------------------------------------------------------------
  if (dbg_skip & 4) return;
  if (JAC && !DIRECT) {
    double2* dst = reinterpret_cast<double2*>(mb + L.off[VF_J] + base);
    const double2* src = reinterpret_cast<const double2*>(tileJ);
    for (int t = threadIdx.x; t < nvals / 2; t += kThreads) __stcs(dst + t, src[t]);
  }
  if (RES) {
    double* F = mb + L.off[VF_F] + (size_t)D * i0;
    for (int t = threadIdx.x; t < D * nT; t += kThreads) F[t] = tileF[t];
  }
#ifdef VF_PHASE_PROF
  if (prof) {
    const long long tk4 = clock64();
    double* info = mb + L.off[VF_INFO];
    atomicAdd(info + 8, (double)(tk1 - tk0));
    atomicAdd(info + 9, (double)(tk2 - tk1));
    atomicAdd(info + 10, (double)(tk3 - tk2));
    atomicAdd(info + 11, (double)(tk4 - tk3));
    atomicAdd(info + 12, 1.0);
    atomicAdd(info + 13, (double)(ta - tk0));
    atomicAdd(info + 14, (double)(tb - ta));
    atomicAdd(info + 15, (double)(tc - tb));
  }
#endif
#undef VF_PROBE
}


template <int D, bool JAC, bool RES>
__global__ void facet_bc_kernel(EngineDev E, int member, double dt, int is_static, JacMix mix,
                                const int* __restrict__ touch_nodes, int n_touch) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_touch) return;
  const int i = touch_nodes[t];
  double* mb = E.members + (size_t)member * E.L.stride;
  const Layout& L = E.L;
  const PropView pv = member_props<D>(E, mb);
  StateView sv;
  sv.u1 = mb + L.off[VF_U1];
  sv.u0 = is_static ? sv.u1 : mb + L.off[VF_U0];
  sv.v0 = mb + L.off[VF_V0];
  sv.a0 = mb + L.off[VF_A0];
  sv.p1 = mb + L.off[VF_P1];
  sv.dt = dt;
  sv.is_static = is_static;
  sv.mix = mix;
  double* F = mb + L.off[VF_F];
  double res[D];
#pragma unroll
  for (int c = 0; c < D; ++c) res[c] = RES ? F[D * i + c] : 0.0;
  assemble_node_facets_bc<D, JAC, RES>(i, E.mesh, pv, sv,
                                       mb + L.off[VF_J] + (size_t)D * D * E.mesh.brptr[i], res);
  if (RES) {
#pragma unroll
    for (int c = 0; c < D; ++c) F[D * i + c] = res[c];
  }
}

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

// partial[b][j] = sum over the block's chunk of V_j[i] w[i]; fixed-order reductions so the
// result is bit-reproducible; a second kernel adds the partials in block order.

// Thread-per-node gather writing the node's block row straight to the global CSR array (no
// shared-memory slice).  Used for tetrahedra, where a block row is ~1 KB: staging it in shared
// memory caps the resident threads at ~200 per SM, while here occupancy is bounded by
// registers only.  Rows are private to their thread, so the accumulation is still
// deterministic; the read-modify-write traffic stays in L1/L2.
template <int D, bool JAC, bool RES>
__global__ void __launch_bounds__(128, 3)
asm_node_global_kernel(EngineDev E, int member, double dt, int is_static, JacMix mix) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= E.mesh.nn) return;
  double* mb = E.members + (size_t)member * E.L.stride;
  const Layout& L = E.L;
  PropView pv = member_props<D>(E, mb);
  StateView sv;
  sv.u1 = mb + L.off[VF_U1];
  sv.u0 = is_static ? sv.u1 : mb + L.off[VF_U0];
  sv.v0 = mb + L.off[VF_V0];
  sv.a0 = mb + L.off[VF_A0];
  sv.p1 = mb + L.off[VF_P1];
  sv.dt = dt;
  sv.is_static = is_static;
  sv.mix = mix;
  double res[D];
  assemble_node<D, JAC, RES>(i, E.mesh, pv, sv,
                             JAC ? mb + L.off[VF_J] + (size_t)D * D * E.mesh.brptr[i] : nullptr, res);
  if (RES) {
    double* F = mb + L.off[VF_F];
#pragma unroll
    for (int c = 0; c < D; ++c) F[D * i + c] = res[c];
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

// d F_u / d p1 (transient.py:423-435): thread per pressure facet.  res_a += mw (1 + delta_ab)
// p_b cof(F) N for facet vertices a, b (assemble_node_facets_bc), so the (a, b) block is
// mw (1 + delta_ab) cof(F) N.  Dirichlet rows are not touched (the reference applies none).
template <int D>
__global__ void pressure_control_kernel(EngineDev E, int member, double* __restrict__ out) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= E.mesh.nfp) return;
  const MeshView& m = E.mesh;
  const double* mb = E.members + (size_t)member * E.L.stride;
  const double* u1 = mb + E.L.off[VF_U1];
  const int e = m.pf_cell[f], o = m.pf_opp[f];
  int nd[D + 1];
  double x[D + 1][D];
  load_cell<D>(m, e, nd, x);
  CellGeo<D> g;
  p1_geometry(x, g);
  double N[D], meas;
  facet_geometry<D>(g, o, N, meas);
  double U[D + 1][D], gu[D][D];
  gather_vec<D>(u1, nd, U);
  grad_u<D>(g, U, gu);
  const double mw = meas / double(D * (D + 1));
  double c[D];
  cof_normal(gu, N, c);
  double* dst = out + (size_t)f * D * D * D;
  int ia = 0;
  for (int a = 0; a <= D; ++a) {
    if (a == o) continue;
    int ib = 0;
    for (int b = 0; b <= D; ++b) {
      if (b == o) continue;
      const double w = mw * (a == b ? 2.0 : 1.0);
      for (int k = 0; k < D; ++k) dst[(ia * D + ib) * D + k] = w * c[k];
      ++ib;
    }
    ++ia;
  }
}

// Nodal Newmark residuals F_v = v1 - v_nmk(u1, u0, v0, a0), F_a = a1 - a_nmk(...)
// (transient.py:374-377): streaming, 6 reads + 2 writes per DOF.
__global__ void newmark_res_kernel(EngineDev E, int member, NewmarkCoef nc, double* fv,
                                   double* fa) {
  const double* mb = E.members + (size_t)member * E.L.stride;
  const double* u1 = mb + E.L.off[VF_U1];
  const double* v1 = mb + E.L.off[VF_V1];
  const double* a1 = mb + E.L.off[VF_A1];
  const double* u0 = mb + E.L.off[VF_U0];
  const double* v0 = mb + E.L.off[VF_V0];
  const double* a0 = mb + E.L.off[VF_A0];
  const size_t n = (size_t)E.mesh.dim * E.mesh.nn;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const double u = u1[i], p0 = u0[i], pv = v0[i], pa = a0[i];
    fv[i] = v1[i] - newmark_v(nc, u, p0, pv, pa);
    fa[i] = a1[i] - newmark_a(nc, u, p0, pv, pa);
  }
}

// Minimum fluid area ("glottal width", postprocess/solid.py:487-501) of a batch of stored
// displacement states: one CTA per state.  The area vector starts from the member's current
// fluid area (entries no solid DOF maps to keep their value), the mapped entries are
// 2 (ymid - y) of the deformed surface (transient.py:836-848), then a block-wide minimum.
template <int D>
__global__ void glottal_width_series_kernel(EngineDev E, int member, const double* __restrict__ u_hist,
                                            size_t ldu, double* __restrict__ out) {
  extern __shared__ double s_area[];
  __shared__ double s_red[32];
  const double* mb = E.members + (size_t)member * E.L.stride;
  const double* base = mb + E.L.off[VF_AREA];
  const double ymid = (mb + E.L.off[VF_SCAL])[SC_YMID];
  const double* u = u_hist + (size_t)blockIdx.x * ldu;
  const int na = E.n_fluid * E.ns;
  for (int k = threadIdx.x; k < na; k += blockDim.x) s_area[k] = base[k];
  __syncthreads();
  for (int k = threadIdx.x; k < E.n_fsi; k += blockDim.x) {
    const int i = E.fsi_solid[k];
    const double y = E.mesh.xyz[(size_t)1 * E.mesh.nn + i] + u[D * i + 1];
    s_area[E.fsi_fluid[k]] = 2.0 * (ymid - y);
  }
  __syncthreads();
  double mn = INFINITY;
  for (int k = threadIdx.x; k < na; k += blockDim.x) mn = fmin(mn, s_area[k]);
  for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mn;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < (int)(blockDim.x >> 5); ++q) mn = fmin(mn, s_red[q]);
    out[blockIdx.x] = mn;
  }
}

__global__ void fluid_kernel(EngineDev E, int member0) {
  double* mb = E.members + (size_t)(member0 + blockIdx.x) * E.L.stride;
  const Layout& L = E.L;
  const int wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int f = wid; f < E.n_fluid; f += nw) {
    bernoulli_channel(E.fluid_kind, E.idx_sep, E.ns, E.s + (size_t)f * E.ns,
                      mb + L.off[VF_AREA] + (size_t)f * E.ns, (mb + L.off[VF_PSUB])[f],
                      (mb + L.off[VF_PSUP])[f], mb + L.off[VF_FPROP] + (size_t)f * FP_COUNT,
                      mb + L.off[VF_Q1] + f, mb + L.off[VF_PF1] + (size_t)f * E.ns);
  }
}

// Staging <-> member blocks (host-buffer entry points move one contiguous buffer over PCIe
// and let the device do the per-member scatter/gather).
__global__ void pack_state_kernel(EngineDev E, double* staged, int to_members) {
  double* mb = E.members + (size_t)blockIdx.x * E.L.stride;
  const Layout& L = E.L;
  const int N = E.N, nq = E.n_fluid, np = E.n_fluid * E.ns;
  const size_t SS = (size_t)3 * N + nq + np;
  double* st = staged + (size_t)blockIdx.x * SS;
  const int ids[5] = {VF_U0, VF_V0, VF_A0, VF_Q0, VF_P0};
  const int cnt[5] = {N, N, N, nq, np};
  size_t o = 0;
  for (int k = 0; k < 5; ++k) {
    double* arr = mb + L.off[ids[k]];
    for (int t = threadIdx.x; t < cnt[k]; t += blockDim.x) {
      if (to_members) arr[t] = st[o + t];
      else st[o + t] = arr[t];
    }
    o += cnt[k];
  }
}

__global__ void pack_array_kernel(EngineDev E, const double* staged, int array_id, int count) {
  double* arr = E.members + (size_t)blockIdx.x * E.L.stride + E.L.off[array_id];
  const double* st = staged + (size_t)blockIdx.x * count;
  for (int t = threadIdx.x; t < count; t += blockDim.x) arr[t] = st[t];
}

// ---- persistent per-member kernel ---------------------------------------------------------

enum MemberMode { MODE_SOLVE_SOLID = 0, MODE_INTEGRATE = 1, MODE_LINEAR_SOLVE = 2 };

template <int D>
__device__ void write_history(const EngineDev& E, double* mb, double* hist_state,
                              double* hist_info, size_t row, bool zero_info) {
  const Layout& L = E.L;
  const int N = E.N, nq = E.n_fluid, np = E.n_fluid * E.ns;
  if (hist_state) {
    double* dst = hist_state + row * (size_t)(3 * N + nq + np);
    const double* u = mb + L.off[VF_U0];
    const double* v = mb + L.off[VF_V0];
    const double* a = mb + L.off[VF_A0];
    for (int t = threadIdx.x; t < N; t += blockDim.x) {
      dst[t] = u[t];
      dst[N + t] = v[t];
      dst[2 * N + t] = a[t];
    }
    for (int t = threadIdx.x; t < nq; t += blockDim.x) dst[3 * N + t] = (mb + L.off[VF_Q0])[t];
    for (int t = threadIdx.x; t < np; t += blockDim.x)
      dst[3 * N + nq + t] = (mb + L.off[VF_P0])[t];
  }
  if (hist_info && threadIdx.x == 0) {
    const double* info = mb + L.off[VF_INFO];
    double* dst = hist_info + row * 4;
    dst[0] = zero_info ? 0.0 : info[INFO_NUM_ITER];
    dst[1] = zero_info ? 0.0 : info[INFO_ABS_ERR];
    dst[2] = zero_info ? 0.0 : info[INFO_REL_ERR];
    dst[3] = info[INFO_MIN_AREA];
  }
}

template <int D, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
member_kernel(EngineDev E, int member0, int mode, int nsteps, const double* __restrict__ dts,
              int nctrl, const double* __restrict__ controls, SolverOpts opt, double dt_single,
              double* hist_state, double* hist_info, const double* lin_b, double* lin_x,
              int smem_flags) {
  __shared__ BlockShared sh;
  extern __shared__ double dsm[];
  const int b = member0 + blockIdx.x;
  double* mb = E.members + (size_t)b * E.L.stride;
  const Layout& L = E.L;
  const int N = E.N, nn = E.mesh.nn;
  // single solves keep the member's global J / F / dx (they are API-visible results)
  const SolverWork W = make_work(E, mb, dsm, mode == MODE_INTEGRATE ? smem_flags : 0);
  if (threadIdx.x < 8) sh.cyc[threadIdx.x] = 0;
  // the dense inverse never outlives a launch: results then depend on the inputs of this
  // launch only (properties may have been rewritten in between), and ensemble members and
  // single runs take the same path
  if (threadIdx.x == 0 && W.pstate) W.pstate[1] = 1.0;
  if (threadIdx.x == 0) sh.bc[6] = 0.0;  // transient inverse not validated in this launch yet
  const long long t_start = clock64();
  __syncthreads();

  if (mode == MODE_SOLVE_SOLID) {
    // a single transient solve is cheaper with the polynomial preconditioner than one
    // inversion; the stiff static problem (no mass term) is where the inverse pays
    blk_solve_solid<D>(E, mb, W, dt_single, opt, sh, opt.is_static != 0 || E.dense == 2);
    return;
  }
  if (mode == MODE_LINEAR_SOLVE) {
    double resid, bnorm;
    blk_compute_dinv<D>(E, W.J, W.Dinv);
    __syncthreads();
    const int it = blk_gmres<D>(E, W, lin_b, lin_x, opt, sh, &resid, &bnorm);
    if (threadIdx.x == 0) {
      double* info = mb + L.off[VF_INFO];
      info[INFO_GMRES_ITERS] = double(it);
      info[INFO_GMRES_RESID] = resid;
      info[INFO_BNORM] = bnorm;
    }
    return;
  }

  // MODE_INTEGRATE
  double* u0 = mb + L.off[VF_U0];
  double* v0 = mb + L.off[VF_V0];
  double* a0 = mb + L.off[VF_A0];
  double* u1 = mb + L.off[VF_U1];
  double* v1 = mb + L.off[VF_V1];
  double* a1 = mb + L.off[VF_A1];
  double* q0 = mb + L.off[VF_Q0];
  double* p0 = mb + L.off[VF_P0];
  double* q1 = mb + L.off[VF_Q1];
  double* pf1 = mb + L.off[VF_PF1];
  double* p1 = mb + L.off[VF_P1];
  double* psub = mb + L.off[VF_PSUB];
  double* psup = mb + L.off[VF_PSUP];
  const size_t hrow0 = (size_t)blockIdx.x * (size_t)(nsteps + 1);

  // row 0 of the history: the initial state with zero solver info (forward.py:75-86);
  // the min-area entry is evaluated from the initial displacement
  {
    const double ymid = (mb + L.off[VF_SCAL])[SC_YMID];
    double* area = mb + L.off[VF_AREA];
    for (int k = threadIdx.x; k < E.n_fsi; k += blockDim.x) {
      const int i = E.fsi_solid[k];
      area[E.fsi_fluid[k]] = 2.0 * (ymid - (E.mesh.xyz[(size_t)E.mesh.nn + i] + u0[D * i + 1]));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double mn = CUDART_INF;
      for (int k = 0; k < E.n_fluid * E.ns; ++k) mn = fmin(mn, area[k]);
      (mb + L.off[VF_INFO])[INFO_MIN_AREA] = mn;
    }
    __syncthreads();
    write_history<D>(E, mb, hist_state, hist_info, hrow0, true);
  }

  for (int n = 0; n < nsteps; ++n) {
    const double dt = dts[n];
    const int ci = min(n, nctrl - 1);
    // set_control (transient.py:797-802)
    for (int f = threadIdx.x; f < E.n_fluid; f += blockDim.x) {
      psub[f] = controls[((size_t)ci * 2 + 0) * E.n_fluid + f];
      psup[f] = controls[((size_t)ci * 2 + 1) * E.n_fluid + f];
    }
    // _set_ini_fluid_state: p1 := 0; p1[solid_dofs] = p0[fluid_dofs]  (transient.py:850-858)
    for (int i = threadIdx.x; i < nn; i += blockDim.x) p1[i] = 0.0;
    // initial guess for the final state = initial state (transient.py:904)
    for (int t = threadIdx.x; t < N; t += blockDim.x) u1[t] = u0[t];
    __syncthreads();
    for (int k = threadIdx.x; k < E.n_fsip; k += blockDim.x) p1[E.fsip_solid[k]] = p0[E.fsip_fluid[k]];
    __syncthreads();

    blk_solve_solid<D>(E, mb, W, dt, opt, sh, E.dense == 2);
    const long long tf = clock64();
    blk_fluid<D>(E, mb, sh);
    if (threadIdx.x == 0) sh.cyc[5] += clock64() - tf;

    // state0 <- state1 (forward.py:184)
    for (int t = threadIdx.x; t < N; t += blockDim.x) {
      u0[t] = u1[t];
      v0[t] = v1[t];
      a0[t] = a1[t];
    }
    for (int t = threadIdx.x; t < E.n_fluid; t += blockDim.x) q0[t] = q1[t];
    for (int t = threadIdx.x; t < E.n_fluid * E.ns; t += blockDim.x) p0[t] = pf1[t];
    __syncthreads();
    write_history<D>(E, mb, hist_state, hist_info, hrow0 + n + 1, false);
  }
  if (threadIdx.x == 0) {
    double* info = mb + L.off[VF_INFO];
    for (int q = 0; q < 6; ++q) info[8 + q] = double(sh.cyc[q]);
    info[14] = double(clock64() - t_start);
    info[15] = double(sh.cyc[7]);  // dense-inverse builds
  }
}

}  // namespace vf

// ======================================= host side ===========================================

using namespace vf;

// layout pinned for the ctypes binding (femvf_b200/_cabi.py, tests/test_cabi.py)
static_assert(sizeof(vf_solver_opts) == 56, "vf_solver_opts layout changed");
static_assert(VF_ARRAY_COUNT == 26, "vf_array_id changed: update _cabi.ARRAY_IDS");

struct vf_engine {
  vf_problem_desc desc;  // scalar fields only are valid after create
  EngineDev dev;
  char* arena;
  size_t arena_bytes;
  int* tile_start_dev;
  int* te_ptr_dev;
  int* te_elem_dev;
  unsigned* pair_info_dev;
  int4* tile_desc_dev;
  int4* te_quad_dev;
  int* tile_halo_dev;
  int* touch_dev;
  int n_touch;
  bool two_phase;
  bool fan_ok;
  std::vector<int32_t> brptr, bcol;
  std::vector<int32_t> pf_nodes;  // (nfp, dim): vertices of every pressure facet, parent-cell order
  int member_threads;
  int64_t launches;
};

namespace {

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Dense inverse preconditioner of the member solver (member_solver.cuh).  Returns 0 (off), 1 (used
// for static solves only: the default) or 2 (also in the time loop: VF_DENSE_PREC=1).
// Measured on config 1/2 (N = 296, profiles/README.md):
//   * static solve with contact (no mass term, stiff): 2326 -> 11 GMRES iterations,
//     18.2 -> 3.8 ms: this is the PETSc-LU stand-in, on by default when the storage is small;
//   * transient steps (inverse of the state-independent part, kept across launches): 10 -> 6
//     iterations, but one mat-vec with the 350 KB fp32 inverse costs ~24 k cycles on ONE SM
//     (L2 -> SM at ~15 B/clk with 8 warps) against ~10 k for the degree-3 polynomial step out
//     of shared memory: 5.9 k vs 7.1 k steps/s on the same box.  Opt-in only.
inline int dense_prec_mode(const vf_problem_desc& d) {
  const char* env = getenv("VF_DENSE_PREC");
  const int want = env ? atoi(env) : -1;
  if (want == 0) return 0;
  const size_t N = (size_t)d.dim * d.nn;
  if (N > (size_t)kMaxDenseN) return 0;
  // the pivot panels (2 * kDenseNb * N doubles) live in the Krylov basis storage
  if ((size_t)d.gmres_restart + 1 < 2 * (size_t)kDenseNb) return 0;
  const size_t bytes = 12 * N * N * (size_t)d.n_members;
  if (want == 1) return bytes <= ((size_t)24 << 30) ? 2 : 0;
  return bytes <= ((size_t)2 << 30) ? 1 : 0;
}
inline bool dense_prec_enabled(const vf_problem_desc& d) { return dense_prec_mode(d) != 0; }

struct ArenaPlan {
  // byte offsets of the shared tables
  size_t xyz, xy, cells, brptr, bcol, n2e_ptr, n2e, n2f_ptr, n2f, pf_cell, pf_opp, bc, tile_start, s,
      fsi_solid, fsi_fluid, fsip_solid, fsip_fluid, te_ptr, te_elem, pair_info, tile_desc, te_quad, tile_halo, touch, gpair, members, total;
  Layout L;
  long long nnz;
  int N;
};

ArenaPlan plan_arena(const vf_problem_desc& d) {
  ArenaPlan P{};
  const int nen = d.dim + 1;
  const long long nnzb = d.brptr_host ? d.brptr_host[d.nn] : 0;
  P.nnz = nnzb * d.dim * d.dim;
  P.N = d.dim * d.nn;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o = align_up(o + bytes, 256);
    return r;
  };
  const int n_n2e = d.n2e_ptr_host ? d.n2e_ptr_host[d.nn] : 0;
  const int n_n2f = d.n2f_ptr_host ? d.n2f_ptr_host[d.nn] : 0;
  P.xyz = take(sizeof(double) * d.dim * d.nn);
  P.xy = take(sizeof(double) * 2 * (d.dim == 2 ? d.nn : 1));
  P.cells = take(sizeof(int) * nen * d.ne);
  P.brptr = take(sizeof(int) * (d.nn + 1));
  P.bcol = take(sizeof(int) * nnzb);
  P.n2e_ptr = take(sizeof(int) * (d.nn + 1));
  P.n2e = take(sizeof(int) * n_n2e);
  P.n2f_ptr = take(sizeof(int) * (d.nn + 1));
  P.n2f = take(sizeof(int) * std::max(n_n2f, 1));
  P.pf_cell = take(sizeof(int) * std::max(d.nfp, 1));
  P.pf_opp = take(sizeof(int) * std::max(d.nfp, 1));
  P.bc = take(d.dim * d.nn);
  P.tile_start = take(sizeof(int) * (d.ntiles + 1));
  P.s = take(sizeof(double) * std::max(d.n_fluid * d.ns, 1));
  P.fsi_solid = take(sizeof(int) * std::max(d.n_fsi, 1));
  P.fsi_fluid = take(sizeof(int) * std::max(d.n_fsi, 1));
  P.fsip_solid = take(sizeof(int) * std::max(d.n_fsip, 1));
  P.fsip_fluid = take(sizeof(int) * std::max(d.n_fsip, 1));
  const int n_te = d.te_ptr_host ? d.te_ptr_host[d.ntiles] : 0;
  P.te_ptr = take(sizeof(int) * (d.ntiles + 1));
  P.te_elem = take(sizeof(int) * std::max(n_te, 1));
  P.pair_info = take(sizeof(unsigned) * std::max(d.te_ptr_host ? n_n2e : 0, 1));
  P.tile_desc = take(sizeof(int) * 12 * (d.te_ptr_host ? d.ntiles : 1));
  P.te_quad = take(sizeof(int) * 4 * std::max(n_te, 1));
  P.tile_halo = take(sizeof(int) * std::max(d.te_ptr_host ? d.n_tile_halo : 0, 1));
  P.touch = take(sizeof(int) * std::max(d.nn, 1));
  P.gpair = take(sizeof(unsigned) * std::max(d.gpair_host ? n_n2e : 0, 1));
  P.members = o;

  // member block (offsets in doubles, each array aligned to 16 doubles = 128 B)
  Layout& L = P.L;
  size_t m = 0;
  auto mtake = [&](size_t count) {
    size_t r = m;
    m = align_up(m + std::max<size_t>(count, 1), 16);
    return r;
  };
  const size_t N = P.N, nq = d.n_fluid, np = (size_t)d.n_fluid * d.ns, ne = d.ne;
  auto pub = [&](int id, size_t count) {
    L.off[id] = mtake(count);
    L.cnt[id] = count;
  };
  pub(VF_U0, N); pub(VF_V0, N); pub(VF_A0, N); pub(VF_Q0, nq); pub(VF_P0, np);
  pub(VF_U1, N); pub(VF_V1, N); pub(VF_A1, N); pub(VF_Q1, nq); pub(VF_PF1, np);
  pub(VF_PSUB, nq); pub(VF_PSUP, nq);
  pub(VF_P1, d.nn);
  pub(VF_AREA, np);
  pub(VF_RHO, ne); pub(VF_ETA, ne); pub(VF_EMOD, ne);
  pub(VF_EMOD_M, d.membrane ? ne : 1); pub(VF_NU_M, d.membrane ? ne : 1);
  pub(VF_TH_M, d.membrane ? ne : 1);
  pub(VF_SCAL, SC_COUNT);
  pub(VF_FPROP, nq * FP_COUNT);
  pub(VF_F, N);
  pub(VF_J, (size_t)P.nnz);
  pub(VF_DX, N);
  pub(VF_INFO, kInfoCount);
  const size_t mr = d.gmres_restart;
  L.Dinv = mtake((size_t)d.nn * d.dim * d.dim);
  L.V = mtake((mr + 1) * N);
  L.w = mtake(N);
  L.z = mtake(N);
  L.H = mtake((mr + 1) * mr);
  L.cs = mtake(mr);
  L.sn = mtake(mr);
  L.g = mtake(mr + 1);
  L.y = mtake(mr);
  L.xk = mtake(N);
  L.Pinv = 0;
  L.pstate = 0;
  L.Pf = 0;
  L.Pscr = 0;
  if (dense_prec_enabled(d)) {
    L.Pinv = mtake(N * N);
    const size_t ldp = (size_t)dense_ldp((int)N);
    L.Pf = mtake((N * ldp + 1) / 2);   // fp32, rows padded to 128 bytes
    L.Pscr = mtake(32 * ldp);          // up to 32 warps
    L.pstate = mtake(4);
  }
  L.stride = align_up(m, 32);
  P.total = P.members + sizeof(double) * L.stride * (size_t)d.n_members;
  return P;
}

int check_desc(const vf_problem_desc* d) {
  if (!d) return fail("null problem descriptor");
  if (d->dim != 2 && d->dim != 3) return fail("dim must be 2 or 3");
  if (d->nn <= 0 || d->ne <= 0) return fail("empty mesh");
  if (d->n_members <= 0) return fail("n_members must be positive");
  if (d->gmres_restart <= 0 || d->gmres_restart > kMaxRestart)
    return fail("gmres_restart must be in [1, 128]");
  if (d->n_fluid < 0 || d->ns < 0 || d->n_fsi < 0 || d->n_fsip < 0)
    return fail("negative fluid sizes");
  if (d->ntiles <= 0 || !d->tile_start_host) return fail("missing assembly tile partition");
  if (d->tile_threads <= 0 || d->tile_threads > 1024) return fail("invalid tile_threads");
  if ((size_t)d->tile_max_values * sizeof(double) > 227 * 1024)
    return fail("tile exceeds the 227 KB shared memory of an SM");
  return 0;
}

cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

double* member_array(vf_engine* e, int id, int member) {
  return e->dev.members + (size_t)member * e->dev.L.stride + e->dev.L.off[id];
}

// dynamic shared memory of asm_tile2_kernel: records, region A (CSR slice aliasing the nodal
// staging: coordinates + u/v/a = 64 bytes per own or halo vertex), F, index slices
size_t tile2_smem_bytes(const vf_problem_desc& d, bool direct = false) {
  const size_t idx_words = (size_t)d.max_tile_pairs + 2 * ((size_t)d.tile_threads + 1);
  const size_t region_a = direct ? (size_t)8 * d.max_tile_verts
                                 : std::max((size_t)d.tile_max_values, (size_t)8 * d.max_tile_verts);
  return sizeof(double) * ((size_t)d.max_tile_elems * kRec2D + region_a +
                           2 * (size_t)d.tile_threads + ((idx_words + 3) / 4) * 2);
}

SolverOpts to_opts(const vf_solver_opts* o) {
  SolverOpts s;
  if (o) {
    s.newton_abs_tol = o->newton_abs_tol;
    s.newton_rel_tol = o->newton_rel_tol;
    s.newton_max_iter = o->newton_max_iter;
    s.gmres_rel_tol = o->gmres_rel_tol;
    s.gmres_abs_tol = o->gmres_abs_tol;
    s.gmres_max_iter = o->gmres_max_iter;
    s.is_static = o->is_static;
    s.poly_degree = std::min(std::max(o->poly_degree, 0), 8);
  } else {
    s.newton_abs_tol = 1e-8;   // solverconst.py:1-6
    s.newton_rel_tol = 1e-10;
    s.newton_max_iter = 50;
    s.gmres_rel_tol = 1e-13;
    s.gmres_abs_tol = 0.0;
    s.gmres_max_iter = 2000;
    s.is_static = 0;
    s.poly_degree = 3;
  }
  return s;
}

template <int D>
int launch_member(vf_engine* e, int member0, int count, int mode, int nsteps, const double* dts,
                  int nctrl, const double* controls, const SolverOpts& opt, double dt_single,
                  double* hist_state, double* hist_info, const double* lin_b, double* lin_x,
                  cudaStream_t st) {
  // place the solver working set in shared memory when it fits (time loop only)
  int flags = 0;
  size_t smem = 0;
  int per_sm = 1;
  if (mode == MODE_INTEGRATE) {
    // resident CTAs per SM wanted for ensembles (VF_MEMBER_PER_SM, default 2): the shared
    // memory budget of one CTA shrinks accordingly and the plan below keeps what fits
    static const char* env_k = getenv("VF_MEMBER_PER_SM");
    per_sm = env_k ? std::min(std::max(atoi(env_k), 1), 4) : 2;
    if (count <= 148) per_sm = 1;
    const size_t N = e->dev.N;
    const size_t budget = per_sm == 1 ? 200 * 1024 : (227 * 1024) / per_sm - 2048;
    auto pad = [](size_t n) { return (n + 1) & ~size_t(1); };
    const size_t small = 8 * (5 * pad(N) + pad((size_t)e->desc.nn * D * D));
    const size_t basis = 8 * pad((size_t)(e->dev.restart + 1) * N);
    const size_t jac = 8 * pad((size_t)e->dev.nnz);
    const size_t hess = 8 * pad((size_t)(e->dev.restart + 1) * e->dev.restart);
    if (small <= budget) { flags |= 1; smem += small; }
    if ((flags & 1) && smem + hess <= budget) { flags |= 8; smem += hess; }
    if ((flags & 1) && smem + basis <= budget) { flags |= 2; smem += basis; }
    if ((flags & 2) && smem + jac <= budget) { flags |= 4; smem += jac; }
    static const char* env = getenv("VF_MEMBER_SMEM");
    if (env && atoi(env) == 0) { flags = 0; smem = 0; }
  }
#define VF_LAUNCH_MEMBER(NT_, MB_)                                                                \
  do {                                                                                            \
    VF_CUDA(cudaFuncSetAttribute(member_kernel<D, NT_, MB_>,                                      \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
    member_kernel<D, NT_, MB_><<<count, NT_, smem, st>>>(e->dev, member0, mode, nsteps, dts,      \
                                                        nctrl, controls, opt, dt_single,          \
                                                        hist_state, hist_info, lin_b, lin_x,      \
                                                        flags);                                   \
  } while (0)
  // two resident CTAs per SM when the working set leaves room for it and there are enough
  // members to use them (ensembles); one fat CTA otherwise
  if (e->member_threads == 256) {
    if (per_sm >= 4) VF_LAUNCH_MEMBER(256, 4);
    else if (per_sm == 3) VF_LAUNCH_MEMBER(256, 3);
    else if (per_sm == 2) VF_LAUNCH_MEMBER(256, 2);
    else VF_LAUNCH_MEMBER(256, 1);
  } else {
    VF_LAUNCH_MEMBER(512, 1);
  }
#undef VF_LAUNCH_MEMBER
  e->launches += 1;
  VF_CUDA(cudaGetLastError());
  return 0;
}

int launch_member_any(vf_engine* e, int member0, int count, int mode, int nsteps,
                      const double* dts, int nctrl, const double* controls,
                      const SolverOpts& opt, double dt_single, double* hist_state,
                      double* hist_info, const double* lin_b, double* lin_x, cudaStream_t st) {
  if (member0 < 0 || count <= 0 || member0 + count > e->desc.n_members)
    return fail("member range out of bounds");
  if (e->desc.dim == 2)
    return launch_member<2>(e, member0, count, mode, nsteps, dts, nctrl, controls, opt,
                            dt_single, hist_state, hist_info, lin_b, lin_x, st);
  return launch_member<3>(e, member0, count, mode, nsteps, dts, nctrl, controls, opt, dt_single,
                          hist_state, hist_info, lin_b, lin_x, st);
}

}  // namespace

extern "C" {

const char* vf_last_error(void) { return g_err.c_str(); }

int vf_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

size_t vf_arena_bytes(const vf_problem_desc* desc) {
  if (check_desc(desc)) return 0;
  return plan_arena(*desc).total;
}

int vf_create(const vf_problem_desc* desc, void* arena_dev, size_t arena_bytes, void* stream,
              vf_engine** out) {
  if (!out) return fail("null output handle");
  *out = nullptr;
  if (check_desc(desc)) return 1;
  if (vf_device_count() <= 0)
    return fail("no CUDA device available: vffem_b200 has no CPU fallback");
  const vf_problem_desc& d = *desc;
  ArenaPlan P = plan_arena(d);
  if (!arena_dev || arena_bytes < P.total) return fail("arena too small");
  if (reinterpret_cast<uintptr_t>(arena_dev) % 256 != 0) return fail("arena must be 256-byte aligned");
  cudaStream_t st = as_stream(stream);
  char* A = static_cast<char*>(arena_dev);
  const int nen = d.dim + 1;
  const int nnzb = d.brptr_host[d.nn];
  const int n_n2e = d.n2e_ptr_host[d.nn];
  const int n_n2f = d.n2f_ptr_host[d.nn];

  {
    // vf_integrate_host stages through cudaMallocAsync: keep freed blocks in the device's default
    // pool instead of returning them to the driver at every synchronisation point
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    (void)cudaGetLastError();
  }
  VF_CUDA(cudaMemsetAsync(A, 0, P.total, st));
  auto up = [&](size_t off, const void* src, size_t bytes) -> cudaError_t {
    if (bytes == 0) return cudaSuccess;
    return cudaMemcpyAsync(A + off, src, bytes, cudaMemcpyHostToDevice, st);
  };
  VF_CUDA(up(P.xyz, d.xyz_host, sizeof(double) * d.dim * d.nn));
  std::vector<double> xy;
  if (d.dim == 2) {
    xy.resize((size_t)2 * d.nn);
    for (int i = 0; i < d.nn; ++i) {
      xy[2 * (size_t)i] = d.xyz_host[i];
      xy[2 * (size_t)i + 1] = d.xyz_host[(size_t)d.nn + i];
    }
    VF_CUDA(up(P.xy, xy.data(), sizeof(double) * xy.size()));
  }
  VF_CUDA(up(P.cells, d.cells_host, sizeof(int) * nen * d.ne));
  VF_CUDA(up(P.brptr, d.brptr_host, sizeof(int) * (d.nn + 1)));
  VF_CUDA(up(P.bcol, d.bcol_host, sizeof(int) * nnzb));
  VF_CUDA(up(P.n2e_ptr, d.n2e_ptr_host, sizeof(int) * (d.nn + 1)));
  VF_CUDA(up(P.n2e, d.n2e_host, sizeof(int) * n_n2e));
  VF_CUDA(up(P.n2f_ptr, d.n2f_ptr_host, sizeof(int) * (d.nn + 1)));
  VF_CUDA(up(P.n2f, d.n2f_host, sizeof(int) * n_n2f));
  VF_CUDA(up(P.pf_cell, d.pf_cell_host, sizeof(int) * d.nfp));
  VF_CUDA(up(P.pf_opp, d.pf_opp_host, sizeof(int) * d.nfp));
  VF_CUDA(up(P.bc, d.bc_host, d.dim * d.nn));
  VF_CUDA(up(P.tile_start, d.tile_start_host, sizeof(int) * (d.ntiles + 1)));
  VF_CUDA(up(P.s, d.s_host, sizeof(double) * d.n_fluid * d.ns));
  VF_CUDA(up(P.fsi_solid, d.fsi_solid_host, sizeof(int) * d.n_fsi));
  VF_CUDA(up(P.fsi_fluid, d.fsi_fluid_host, sizeof(int) * d.n_fsi));
  VF_CUDA(up(P.fsip_solid, d.fsip_solid_host, sizeof(int) * d.n_fsip));
  VF_CUDA(up(P.fsip_fluid, d.fsip_fluid_host, sizeof(int) * d.n_fsip));
  const bool two_phase = d.te_ptr_host && d.te_elem_host && d.pair_info_host &&
                         d.tile_desc_host && d.te_quad_host && d.tile_halo_host && d.dim == 2 &&
                         d.tile2_threads > 0;
  std::vector<int> touch;
  for (int i = 0; i < d.nn; ++i) {
    bool t = d.n2f_ptr_host[i + 1] > d.n2f_ptr_host[i];
    for (int c = 0; c < d.dim && !t; ++c) t = d.bc_host[d.dim * i + c] != 0;
    if (t) touch.push_back(i);
  }
  VF_CUDA(up(P.touch, touch.data(), sizeof(int) * touch.size()));
  if (d.gpair_host) VF_CUDA(up(P.gpair, d.gpair_host, sizeof(unsigned) * n_n2e));
  if (two_phase) {
    VF_CUDA(up(P.te_ptr, d.te_ptr_host, sizeof(int) * (d.ntiles + 1)));
    VF_CUDA(up(P.te_elem, d.te_elem_host, sizeof(int) * d.te_ptr_host[d.ntiles]));
    VF_CUDA(up(P.pair_info, d.pair_info_host, sizeof(unsigned) * n_n2e));
    VF_CUDA(up(P.tile_desc, d.tile_desc_host, sizeof(int) * 12 * d.ntiles));
    VF_CUDA(up(P.te_quad, d.te_quad_host, sizeof(int) * 4 * d.te_ptr_host[d.ntiles]));
    VF_CUDA(up(P.tile_halo, d.tile_halo_host, sizeof(int) * d.n_tile_halo));
  }
  VF_CUDA(cudaStreamSynchronize(st));

  vf_engine* e = new vf_engine();
  e->desc = d;
  e->arena = A;
  e->arena_bytes = P.total;
  e->launches = 0;
  e->brptr.assign(d.brptr_host, d.brptr_host + d.nn + 1);
  e->bcol.assign(d.bcol_host, d.bcol_host + nnzb);
  e->pf_nodes.resize((size_t)d.nfp * d.dim);
  for (int f = 0; f < d.nfp; ++f) {
    int k = 0;
    for (int a = 0; a <= d.dim; ++a)
      if (a != d.pf_opp_host[f])
        e->pf_nodes[(size_t)f * d.dim + k++] = d.cells_host[(size_t)a * d.ne + d.pf_cell_host[f]];
  }
  // host pointers of the descriptor are not retained
  e->desc.xyz_host = nullptr; e->desc.cells_host = nullptr; e->desc.brptr_host = nullptr;
  e->desc.bcol_host = nullptr; e->desc.n2e_ptr_host = nullptr; e->desc.n2e_host = nullptr;
  e->desc.n2f_ptr_host = nullptr; e->desc.n2f_host = nullptr; e->desc.pf_cell_host = nullptr;
  e->desc.pf_opp_host = nullptr; e->desc.bc_host = nullptr; e->desc.tile_start_host = nullptr;
  e->desc.s_host = nullptr; e->desc.fsi_solid_host = nullptr; e->desc.fsi_fluid_host = nullptr;
  e->desc.fsip_solid_host = nullptr; e->desc.fsip_fluid_host = nullptr;
  e->desc.te_ptr_host = nullptr; e->desc.te_elem_host = nullptr; e->desc.pair_info_host = nullptr;
  e->desc.tile_desc_host = nullptr; e->desc.te_quad_host = nullptr;
  e->desc.tile_halo_host = nullptr;
  e->desc.gpair_host = nullptr;
  e->two_phase = two_phase;
  e->fan_ok = d.fan_ok != 0;
  e->touch_dev = reinterpret_cast<int*>(A + P.touch);
  e->n_touch = (int)touch.size();
  e->te_ptr_dev = reinterpret_cast<int*>(A + P.te_ptr);
  e->te_elem_dev = reinterpret_cast<int*>(A + P.te_elem);
  e->pair_info_dev = reinterpret_cast<unsigned*>(A + P.pair_info);
  e->tile_desc_dev = reinterpret_cast<int4*>(A + P.tile_desc);
  e->te_quad_dev = reinterpret_cast<int4*>(A + P.te_quad);
  e->tile_halo_dev = reinterpret_cast<int*>(A + P.tile_halo);

  EngineDev& E = e->dev;
  E.mesh.dim = d.dim; E.mesh.nn = d.nn; E.mesh.ne = d.ne; E.mesh.nfp = d.nfp;
  E.mesh.xyz = reinterpret_cast<const double*>(A + P.xyz);
  E.mesh.xy = d.dim == 2 ? reinterpret_cast<const double*>(A + P.xy) : nullptr;
  E.mesh.cells = reinterpret_cast<const int*>(A + P.cells);
  E.mesh.brptr = reinterpret_cast<const int*>(A + P.brptr);
  E.mesh.bcol = reinterpret_cast<const int*>(A + P.bcol);
  E.mesh.n2e_ptr = reinterpret_cast<const int*>(A + P.n2e_ptr);
  E.mesh.n2e = reinterpret_cast<const int*>(A + P.n2e);
  E.mesh.n2f_ptr = reinterpret_cast<const int*>(A + P.n2f_ptr);
  E.mesh.n2f = reinterpret_cast<const int*>(A + P.n2f);
  E.mesh.pf_cell = reinterpret_cast<const int*>(A + P.pf_cell);
  E.mesh.pf_opp = reinterpret_cast<const int*>(A + P.pf_opp);
  E.mesh.bc = reinterpret_cast<const unsigned char*>(A + P.bc);
  E.d = d.dim; E.N = P.N; E.n_fluid = d.n_fluid; E.ns = d.ns; E.n_fsi = d.n_fsi;
  E.fluid_kind = d.fluid_kind; E.idx_sep = d.idx_sep; E.contact = d.contact;
  E.membrane = d.membrane; E.damping = d.damping; E.restart = d.gmres_restart; E.nnz = P.nnz;
  E.dense = dense_prec_mode(d);
  // record-based in-CTA assembly: 2D, fan-ordered adjacency, cell index fits the packed word
  const char* env_rec = getenv("VF_MEMBER_RECORDS");
  const bool rec_on = !(env_rec && atoi(env_rec) == 0);
  E.gpair = (rec_on && d.gpair_host && d.dim == 2 && d.fan_ok && d.ne < 4096)
                ? reinterpret_cast<const unsigned*>(A + P.gpair) : nullptr;
  E.touch = reinterpret_cast<const int*>(A + P.touch);
  E.n_touch = (int)touch.size();
  E.s = reinterpret_cast<const double*>(A + P.s);
  E.fsi_solid = reinterpret_cast<const int*>(A + P.fsi_solid);
  E.fsi_fluid = reinterpret_cast<const int*>(A + P.fsi_fluid);
  E.fsip_solid = reinterpret_cast<const int*>(A + P.fsip_solid);
  E.fsip_fluid = reinterpret_cast<const int*>(A + P.fsip_fluid);
  E.n_fsip = d.n_fsip;
  E.members = reinterpret_cast<double*>(A + P.members);
  E.L = P.L;
  e->tile_start_dev = reinterpret_cast<int*>(A + P.tile_start);
  e->member_threads = (P.N <= 2048) ? 256 : 512;

  // opt in to large dynamic shared memory for the tile kernel
  const int smem = d.tile_max_values * (int)sizeof(double);
  if (d.dim == 2) {
    VF_CUDA(cudaFuncSetAttribute(asm_tile_kernel<2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    VF_CUDA(cudaFuncSetAttribute(asm_tile_kernel<2, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  } else {
    VF_CUDA(cudaFuncSetAttribute(asm_tile_kernel<3, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    VF_CUDA(cudaFuncSetAttribute(asm_tile_kernel<3, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  if (two_phase) {
    const int smem2 = (int)tile2_smem_bytes(d);
    if (smem2 > 227 * 1024) {
      delete e;
      return fail("two-phase tile exceeds the 227 KB shared memory of an SM");
    }
#define VF_SMEM2(K) VF_CUDA(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2))
#define VF_SMEM2_ALL(J_, R_, ROW_)                          \
  VF_SMEM2((asm_tile2_kernel<J_, R_, ROW_, 128, 8>));       \
  VF_SMEM2((asm_tile2_kernel<J_, R_, ROW_, 192, 5>));       \
  VF_SMEM2((asm_tile2_kernel<J_, R_, ROW_, 256, 4>));       \
  VF_SMEM2((asm_tile2_kernel<J_, R_, ROW_, 320, 3>))
    VF_SMEM2_ALL(true, true, 0);
    VF_SMEM2_ALL(true, true, 1);
    VF_SMEM2_ALL(true, true, 2);
    VF_SMEM2_ALL(true, false, 0);
    VF_SMEM2_ALL(true, false, 1);
    VF_SMEM2_ALL(true, false, 2);
    VF_SMEM2_ALL(false, true, 1);
#undef VF_SMEM2_ALL
#undef VF_SMEM2
  }
  *out = e;
  return 0;
}

void vf_destroy(vf_engine* e) { delete e; }

int vf_array_info(const vf_engine* e, int array_id, int member, size_t* byte_offset,
                  size_t* count) {
  if (!e) return fail("null engine");
  if (array_id < 0 || array_id >= VF_ARRAY_COUNT) return fail("invalid array id");
  if (member < 0 || member >= e->desc.n_members) return fail("member out of range");
  const char* p = reinterpret_cast<const char*>(
      e->dev.members + (size_t)member * e->dev.L.stride + e->dev.L.off[array_id]);
  if (byte_offset) *byte_offset = size_t(p - e->arena);
  if (count) *count = e->dev.L.cnt[array_id];
  return 0;
}

int vf_upload(vf_engine* e, int array_id, int member, const double* src_host, size_t count,
              void* stream) {
  size_t off, cnt;
  if (vf_array_info(e, array_id, member, &off, &cnt)) return 1;
  if (count != cnt) return fail("vf_upload: size mismatch for array " + std::to_string(array_id));
  VF_CUDA(cudaMemcpyAsync(e->arena + off, src_host, sizeof(double) * cnt, cudaMemcpyHostToDevice,
                          as_stream(stream)));
  VF_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return 0;
}

int vf_download(vf_engine* e, int array_id, int member, double* dst_host, size_t count,
                void* stream) {
  size_t off, cnt;
  if (vf_array_info(e, array_id, member, &off, &cnt)) return 1;
  if (count != cnt) return fail("vf_download: size mismatch for array " + std::to_string(array_id));
  VF_CUDA(cudaMemcpyAsync(dst_host, e->arena + off, sizeof(double) * cnt, cudaMemcpyDeviceToHost,
                          as_stream(stream)));
  VF_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return 0;
}

int64_t vf_nnz(const vf_engine* e) { return e ? e->dev.nnz : 0; }

int vf_csr_pattern(const vf_engine* e, int32_t* rowptr, int32_t* colidx) {
  if (!e || !rowptr || !colidx) return fail("null argument");
  const int d = e->desc.dim, nn = e->desc.nn;
  int64_t pos = 0;
  for (int i = 0; i < nn; ++i) {
    const int b0 = e->brptr[i], deg = e->brptr[i + 1] - b0;
    for (int a = 0; a < d; ++a) {
      rowptr[d * i + a] = (int32_t)pos;
      for (int k = 0; k < deg; ++k)
        for (int b = 0; b < d; ++b) colidx[pos++] = d * e->bcol[b0 + k] + b;
    }
  }
  rowptr[d * nn] = (int32_t)pos;
  return 0;
}

namespace {
int assemble_impl(vf_engine* e, int member, int flags, double dt, int is_static, const JacMix& mix,
                  void* stream);
}

int vf_assemble(vf_engine* e, int member, int flags, double dt, int is_static, void* stream) {
  if (!e) return fail("null engine");
  return assemble_impl(e, member, flags, dt, is_static,
                       jac_mix_du1(newmark_coef(dt), is_static != 0), stream);
}

int vf_assemble_mix(vf_engine* e, int member, double dt, const double* coef4, int apply_bc,
                    void* stream) {
  if (!e) return fail("null engine");
  if (!coef4) return fail("null coefficient array");
  JacMix mix;
  mix.k = coef4[0];
  mix.c = coef4[1];
  mix.m = coef4[2];
  mix.p = coef4[3];
  mix.bc = apply_bc ? 1 : 0;
  return assemble_impl(e, member, 2, dt, 0, mix, stream);
}

namespace {
int assemble_impl(vf_engine* e, int member, int flags, double dt, int is_static, const JacMix& mix,
                  void* stream) {
  if (member < 0 || member >= e->desc.n_members) return fail("member out of range");
  const bool res = flags & 1, jac = flags & 2;
  if (!res && !jac) return 0;
  cudaStream_t st = as_stream(stream);
  const int grid = e->desc.ntiles, block = e->desc.tile_threads;
  if (e->two_phase) {
    const vf_problem_desc& d = e->desc;
    const size_t smem2 = tile2_smem_bytes(d);
#define VF_LAUNCH_ASM2(J_, R_, ROW_, MT_, MB_)                                                     \
  asm_tile2_kernel<J_, R_, ROW_, MT_, MB_><<<grid, MT_, smem2, st>>>(                  \
      e->dev, member, newmark_coef(dt), is_static, mix, e->tile_desc_dev, e->te_quad_dev,          \
      e->pair_info_dev, e->tile_halo_dev, d.max_tile_elems, d.tile_max_values, d.max_tile_pairs,   \
      d.tile_threads, d.max_tile_verts, pf_dist, dbg_skip)
    const int dbg_skip = getenv("VF_DEBUG_SKIP") ? atoi(getenv("VF_DEBUG_SKIP")) : 0;
    // L2 prefetch distance in tiles: one wave of resident CTAs (148 SMs x 3 CTAs; measured flat
    // between one and two waves, worse below and far above: profiles/README.md)
    const int pf_dist = getenv("VF_PF_DIST") ? atoi(getenv("VF_PF_DIST")) : 3 * 148;
    int v_row = getenv("VF_TILE2_ROW") ? atoi(getenv("VF_TILE2_ROW")) : 2;
    if (v_row == 2 && !e->fan_ok) v_row = 1;
    // occupancy class by CTA size: small CTAs run many per SM so that their phases overlap
    const int nt = d.tile2_threads;
#define VF_ASM2_BY_SIZE(J_, R_, ROW_)                                                              \
  do {                                                                                            \
    if (nt <= 128) VF_LAUNCH_ASM2(J_, R_, ROW_, 128, 8);                                          \
    else if (nt <= 192) VF_LAUNCH_ASM2(J_, R_, ROW_, 192, 5);                                     \
    else if (nt <= 256) VF_LAUNCH_ASM2(J_, R_, ROW_, 256, 4);                                     \
    else VF_LAUNCH_ASM2(J_, R_, ROW_, 320, 3);                                                    \
  } while (0)
#define VF_ASM2_BY_MODE(J_, R_)                                                                    \
  do {                                                                                            \
    if (v_row == 2) VF_ASM2_BY_SIZE(J_, R_, 2);                                                   \
    else if (v_row == 1) VF_ASM2_BY_SIZE(J_, R_, 1);                                              \
    else VF_ASM2_BY_SIZE(J_, R_, 0);                                                              \
  } while (0)
    // Default for the fan-ordered path: rows are stored straight to HBM from phase 2 (every
    // 16-byte entry once; L2 merges the sectors), so no CSR slice is kept in shared memory and
    // 4 CTAs of 256 threads fit per SM (measured 0.4015 -> 0.3751 ms with 80-node tiles;
    // VF_TILE2_DIRECT=0 restores the staged write-out)
    static const char* env_direct = getenv("VF_TILE2_DIRECT");
    if (!(env_direct && atoi(env_direct) == 0) && jac && v_row == 2 && nt <= 320) {
      const size_t smem_d = tile2_smem_bytes(d, true);
#define VF_LAUNCH_ASM2D(R_, MT_, MB_)                                                              \
  do {                                                                                            \
    VF_CUDA(cudaFuncSetAttribute(asm_tile2_kernel<true, R_, 2, MT_, MB_, true>,                   \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d));      \
    asm_tile2_kernel<true, R_, 2, MT_, MB_, true><<<grid, MT_, smem_d, st>>>(                     \
        e->dev, member, newmark_coef(dt), is_static, mix, e->tile_desc_dev, e->te_quad_dev,       \
        e->pair_info_dev, e->tile_halo_dev, d.max_tile_elems, d.tile_max_values,                  \
        d.max_tile_pairs, d.tile_threads, d.max_tile_verts, pf_dist, dbg_skip);                   \
  } while (0)
      if (nt <= 256) {
        if (res) VF_LAUNCH_ASM2D(true, 256, 4); else VF_LAUNCH_ASM2D(false, 256, 4);
      } else {
        if (res) VF_LAUNCH_ASM2D(true, 320, 3); else VF_LAUNCH_ASM2D(false, 320, 3);
      }
#undef VF_LAUNCH_ASM2D
    } else
    if (jac && res) VF_ASM2_BY_MODE(true, true);
    else if (jac) VF_ASM2_BY_MODE(true, false);
    else VF_ASM2_BY_SIZE(false, true, 1);
#undef VF_ASM2_BY_MODE
#undef VF_ASM2_BY_SIZE
#undef VF_LAUNCH_ASM2
    e->launches += 1;
    VF_CUDA(cudaGetLastError());
    if (e->n_touch > 0) {
      const int fb = 128, fg = (e->n_touch + fb - 1) / fb;
      if (jac && res)
        facet_bc_kernel<2, true, true><<<fg, fb, 0, st>>>(e->dev, member, dt, is_static, mix, e->touch_dev, e->n_touch);
      else if (jac)
        facet_bc_kernel<2, true, false><<<fg, fb, 0, st>>>(e->dev, member, dt, is_static, mix, e->touch_dev, e->n_touch);
      else
        facet_bc_kernel<2, false, true><<<fg, fb, 0, st>>>(e->dev, member, dt, is_static, mix, e->touch_dev, e->n_touch);
      e->launches += 1;
      VF_CUDA(cudaGetLastError());
    }
    return 0;
  }
  if (e->desc.dim == 3 && !(getenv("VF_TET_SMEM") && atoi(getenv("VF_TET_SMEM")))) {
    const int nb = 128, ng = (e->desc.nn + nb - 1) / nb;
    if (jac && res) asm_node_global_kernel<3, true, true><<<ng, nb, 0, st>>>(e->dev, member, dt, is_static, mix);
    else if (jac) asm_node_global_kernel<3, true, false><<<ng, nb, 0, st>>>(e->dev, member, dt, is_static, mix);
    else asm_node_global_kernel<3, false, true><<<ng, nb, 0, st>>>(e->dev, member, dt, is_static, mix);
    e->launches += 1;
    VF_CUDA(cudaGetLastError());
    return 0;
  }
  const size_t smem = jac ? (size_t)e->desc.tile_max_values * sizeof(double) : 0;
#define VF_LAUNCH_ASM(D)                                                                          \
  if (jac && res)                                                                                 \
    asm_tile_kernel<D, true, true><<<grid, block, smem, st>>>(e->dev, member, dt, is_static, mix, \
                                                              e->tile_start_dev);                 \
  else if (jac)                                                                                   \
    asm_tile_kernel<D, true, false><<<grid, block, smem, st>>>(e->dev, member, dt, is_static, mix, \
                                                               e->tile_start_dev);                \
  else                                                                                            \
    asm_tile_kernel<D, false, true><<<grid, block, 0, st>>>(e->dev, member, dt, is_static, mix, \
                                                            e->tile_start_dev);
  if (e->desc.dim == 2) {
    VF_LAUNCH_ASM(2)
  } else {
    VF_LAUNCH_ASM(3)
  }
#undef VF_LAUNCH_ASM
  e->launches += 1;
  VF_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace

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

int vf_pressure_control_blocks(vf_engine* e, int member, double* out_dev, int32_t* rows_host,
                               int32_t* cols_host, void* stream) {
  if (!e) return fail("null engine");
  if (member < 0 || member >= e->desc.n_members) return fail("member out of range");
  const int nfp = e->desc.nfp, d = e->desc.dim;
  if (rows_host || cols_host) {
    if (!rows_host || !cols_host) return fail("rows_host and cols_host go together");
    if (e->pf_nodes.size() != (size_t)nfp * d) return fail("facet vertex table missing");
    for (int f = 0; f < nfp; ++f)
      for (int a = 0; a < d; ++a)
        for (int b = 0; b < d; ++b) {
          rows_host[((size_t)f * d + a) * d + b] = e->pf_nodes[(size_t)f * d + a];
          cols_host[((size_t)f * d + a) * d + b] = e->pf_nodes[(size_t)f * d + b];
        }
  }
  if (nfp == 0) return 0;
  if (!out_dev) return fail("null output");
  const int block = 128, grid = (nfp + block - 1) / block;
  if (d == 2) pressure_control_kernel<2><<<grid, block, 0, as_stream(stream)>>>(e->dev, member, out_dev);
  else pressure_control_kernel<3><<<grid, block, 0, as_stream(stream)>>>(e->dev, member, out_dev);
  e->launches += 1;
  VF_CUDA(cudaGetLastError());
  return 0;
}

int vf_glottal_width_series(vf_engine* e, int member, int nt, const double* u_hist_dev, size_t ldu,
                            double* out_dev, void* stream) {
  if (!e) return fail("null engine");
  if (member < 0 || member >= e->desc.n_members) return fail("member out of range");
  if (nt <= 0) return 0;
  if (!u_hist_dev || !out_dev) return fail("null argument");
  const int na = e->desc.n_fluid * e->desc.ns;
  if (na <= 0) return fail("vf_glottal_width_series: the engine has no fluid");
  const size_t smem = sizeof(double) * na;
  if (smem > 48 * 1024) return fail("vf_glottal_width_series: fluid mesh too large");
  if (ldu < (size_t)e->desc.dim * e->desc.nn) return fail("vf_glottal_width_series: ldu < N");
  cudaStream_t st = as_stream(stream);
  if (e->desc.dim == 2)
    glottal_width_series_kernel<2><<<nt, 128, smem, st>>>(e->dev, member, u_hist_dev, ldu, out_dev);
  else
    glottal_width_series_kernel<3><<<nt, 128, smem, st>>>(e->dev, member, u_hist_dev, ldu, out_dev);
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

int vf_newmark_residual(vf_engine* e, int member, double dt, double* fv_dev, double* fa_dev,
                        void* stream) {
  if (!e) return fail("null engine");
  if (member < 0 || member >= e->desc.n_members) return fail("member out of range");
  if (!fv_dev || !fa_dev) return fail("null output");
  if (!(dt > 0.0)) return fail("dt must be positive");
  const size_t n = (size_t)e->desc.dim * e->desc.nn;
  const int block = 256;
  const int grid = (int)std::min<size_t>(148 * 8, (n + block - 1) / block);
  newmark_res_kernel<<<grid, block, 0, as_stream(stream)>>>(e->dev, member, newmark_coef(dt),
                                                             fv_dev, fa_dev);
  e->launches += 1;
  VF_CUDA(cudaGetLastError());
  return 0;
}

int vf_linear_solve(vf_engine* e, int member, const double* b_dev, double* x_dev,
                    const vf_solver_opts* opts, double* info_host, void* stream) {
  if (!e) return fail("null engine");
  cudaStream_t st = as_stream(stream);
  SolverOpts so = to_opts(opts);
  if (launch_member_any(e, member, 1, MODE_LINEAR_SOLVE, 0, nullptr, 0, nullptr, so, 0.0, nullptr,
                        nullptr, b_dev, x_dev, st))
    return 1;
  if (info_host) {
    double info[kInfoCount];
    VF_CUDA(cudaMemcpyAsync(info, member_array(e, VF_INFO, member), sizeof(info),
                            cudaMemcpyDeviceToHost, st));
    VF_CUDA(cudaStreamSynchronize(st));
    info_host[0] = info[INFO_GMRES_ITERS];
    info_host[1] = info[INFO_GMRES_RESID];
    info_host[2] = info[INFO_BNORM];
  }
  return 0;
}

int vf_solve_state1(vf_engine* e, int member0, int count, double dt, const vf_solver_opts* opts,
                    void* stream) {
  if (!e) return fail("null engine");
  SolverOpts so = to_opts(opts);
  return launch_member_any(e, member0, count, MODE_SOLVE_SOLID, 0, nullptr, 0, nullptr, so, dt,
                           nullptr, nullptr, nullptr, nullptr, as_stream(stream));
}

int vf_fluid_solve(vf_engine* e, int member0, int count, void* stream) {
  if (!e) return fail("null engine");
  if (member0 < 0 || count <= 0 || member0 + count > e->desc.n_members)
    return fail("member range out of bounds");
  if (e->desc.n_fluid <= 0) return fail("model has no fluid");
  const int warps = std::min(e->desc.n_fluid, 8);
  fluid_kernel<<<count, 32 * warps, 0, as_stream(stream)>>>(e->dev, member0);
  e->launches += 1;
  VF_CUDA(cudaGetLastError());
  return 0;
}

int vf_integrate(vf_engine* e, int nsteps, const double* dts_host, int ncontrols,
                 const double* controls_host, const vf_solver_opts* opts, double* hist_state_dev,
                 double* hist_info_dev, void* stream) {
  if (!e) return fail("null engine");
  if (nsteps <= 0) return fail("nsteps must be positive");
  if (ncontrols <= 0 || !controls_host || !dts_host) return fail("missing dts/controls");
  if (e->desc.n_fluid <= 0) return fail("vf_integrate needs a coupled fluid");
  cudaStream_t st = as_stream(stream);
  SolverOpts so = to_opts(opts);
  so.is_static = 0;
  const size_t nctl = (size_t)ncontrols * 2 * e->desc.n_fluid;
  double* scratch = nullptr;
  VF_CUDA(cudaMallocAsync(&scratch, sizeof(double) * (nsteps + nctl), st));
  cudaError_t err = cudaMemcpyAsync(scratch, dts_host, sizeof(double) * nsteps,
                                    cudaMemcpyHostToDevice, st);
  if (err == cudaSuccess)
    err = cudaMemcpyAsync(scratch + nsteps, controls_host, sizeof(double) * nctl,
                          cudaMemcpyHostToDevice, st);
  int rc = 0;
  if (err != cudaSuccess) {
    rc = fail(std::string("vf_integrate upload: ") + cudaGetErrorString(err));
  } else {
    rc = launch_member_any(e, 0, e->desc.n_members, MODE_INTEGRATE, nsteps, scratch, ncontrols,
                           scratch + nsteps, so, 0.0, hist_state_dev, hist_info_dev, nullptr,
                           nullptr, st);
  }
  cudaFreeAsync(scratch, st);
  return rc;
}

int vf_integrate_host(vf_engine* e, int nsteps, const double* dts_host, int ncontrols,
                      const double* controls_host, const vf_solver_opts* opts,
                      const double* ini_state_host, const double* emod_host,
                      const double* eta_host, double* fin_state_host, double* info_series_host,
                      void* stream) {
  if (!e) return fail("null engine");
  if (!ini_state_host || !fin_state_host) return fail("null state buffers");
  cudaStream_t st = as_stream(stream);
  const int B = e->desc.n_members;
  const size_t N = e->dev.N, nq = e->desc.n_fluid, np = (size_t)e->desc.n_fluid * e->desc.ns;
  const size_t SS = 3 * N + nq + np, ne = e->desc.ne;
  const size_t hcount = info_series_host ? (size_t)B * (nsteps + 1) * 4 : 0;
  // one staging allocation: [state B*SS][emod B*ne][eta B*ne][info series]
  double* stage = nullptr;
  VF_CUDA(cudaMallocAsync(&stage, sizeof(double) * (B * SS + 2 * B * ne + hcount), st));
  double* st_state = stage;
  double* st_emod = stage + B * SS;
  double* st_eta = st_emod + B * ne;
  double* hist_info = hcount ? st_eta + B * ne : nullptr;
  int rc = 0;
  auto cp = [&](void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
    if (rc) return;
    cudaError_t err = cudaMemcpyAsync(dst, src, bytes, kind, st);
    if (err != cudaSuccess) rc = fail(std::string("vf_integrate_host copy: ") + cudaGetErrorString(err));
  };
  cp(st_state, ini_state_host, sizeof(double) * B * SS, cudaMemcpyHostToDevice);
  if (emod_host) cp(st_emod, emod_host, sizeof(double) * B * ne, cudaMemcpyHostToDevice);
  if (eta_host) cp(st_eta, eta_host, sizeof(double) * B * ne, cudaMemcpyHostToDevice);
  if (rc == 0) {
    pack_state_kernel<<<B, 256, 0, st>>>(e->dev, st_state, 1);
    e->launches += 1;
    if (emod_host) {
      pack_array_kernel<<<B, 256, 0, st>>>(e->dev, st_emod, VF_EMOD, (int)ne);
      e->launches += 1;
    }
    if (eta_host) {
      pack_array_kernel<<<B, 256, 0, st>>>(e->dev, st_eta, VF_ETA, (int)ne);
      e->launches += 1;
    }
    rc = vf_integrate(e, nsteps, dts_host, ncontrols, controls_host, opts, nullptr, hist_info, st);
  }
  if (rc == 0) {
    pack_state_kernel<<<B, 256, 0, st>>>(e->dev, st_state, 0);
    e->launches += 1;
    cp(fin_state_host, st_state, sizeof(double) * B * SS, cudaMemcpyDeviceToHost);
    if (hcount) cp(info_series_host, hist_info, sizeof(double) * hcount, cudaMemcpyDeviceToHost);
  }
  cudaFreeAsync(stage, st);
  cudaError_t err = cudaStreamSynchronize(st);
  if (rc == 0 && err != cudaSuccess) rc = fail(std::string("sync: ") + cudaGetErrorString(err));
  return rc;
}

int64_t vf_launch_count(const vf_engine* e) { return e ? e->launches : 0; }

}  // extern "C"
