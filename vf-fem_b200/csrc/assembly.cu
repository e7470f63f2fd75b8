// Residual + Jacobian assembly kernels and their C-ABI entry points (vf_assemble,
// vf_assemble_mix, vf_set_fan_tables, vf_pressure_control_blocks, vf_newmark_residual).
// See DESIGN.md section 5 for the roofline of each kernel.
#include "engine_internal.h"
#include "tet_tables.h"
#include "dev_util.cuh"
#include "fan_assembly.cuh"

namespace vf {

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



// ---- node-centric fan assembly (triangles) --------------------------------------------------
// One CTA per tile of TN consecutive nodes, one thread per node (fan_assembly.cuh).  Apart from
// the gather of the nodal state, all global traffic of the kernel is bulk asynchronous copies
// issued by one thread (TMA engine, cp.async.bulk + mbarrier), which do not occupy the LSU data
// pipe -- the unit that bounded the round-1 kernel (79 % busy with shared-memory and scattered
// 16-byte global wavefronts, profiles/README.md):
//   stage  thread 0 arms an mbarrier and issues two bulk loads: the tile's ring table and the
//          DG0 properties of the tile's cells (a tile-ordered copy kept by fan_pack_kernel);
//          all threads stage the tile's own and halo vertices (coordinates, u1, v_nmk, a_nmk;
//          16-byte loads, SoA planes in shared memory);
//   walk   every thread walks the fan of its node reading shared memory only and puts the
//          finished 16-byte row entries into the tile's slice of the CSR array in shared memory;
//   store  ONE bulk store streams the slice (contiguous in the CSR array) to HBM; the residual
//          pairs are stored coalesced;
//   tail   the inputs of the tile pf_dist CTAs ahead are pulled into L2.
// Two CTA-wide barriers (after staging, before the bulk store); no per-cell records.
// Shared memory: [CSR slice][ring][properties][xy][u][v][a].

// Tile-ordered copy of the DG0 properties: per tile a block [emod | eta | rho] over its (padded)
// cell list, so that one bulk copy brings a tile's material into shared memory.
__global__ void fan_pack_kernel(EngineDev E, int member, FanTablesDev T, double* mat) {
  const int4 d1 = __ldg(T.desc + 3 * blockIdx.x + 1);
  const int tc0 = d1.y, ncp = d1.z;
  const double* mb = E.members + (size_t)member * E.L.stride;
  const double* emod = mb + E.L.off[VF_EMOD];
  const double* eta = mb + E.L.off[VF_ETA];
  const double* rho = mb + E.L.off[VF_RHO];
  double* dst = mat + (size_t)3 * tc0;
  for (int k = threadIdx.x; k < ncp; k += blockDim.x) {
    const int e = __ldg(T.tcell + tc0 + k);
    dst[k] = emod[e];
    dst[ncp + k] = eta[e];
    dst[2 * ncp + k] = rho[e];
  }
}

template <bool JAC, bool RES, int TN, int MINB>
__global__ void __launch_bounds__(TN, MINB) asm_fan_kernel(
    EngineDev E, int member, NewmarkCoef nc_arg, int is_static, JacMix mix, FanTablesDev T,
    const double* __restrict__ mat_m, int pf_dist) {
  extern __shared__ __align__(128) unsigned char fan_smem[];
  double* s_J = reinterpret_cast<double*>(fan_smem);
  unsigned* s_ring = reinterpret_cast<unsigned*>(s_J + (JAC ? 4 * (size_t)T.max_blocks : 0));
  double* s_mat = reinterpret_cast<double*>(s_ring + (size_t)T.max_rows * TN);
  D2* s_xy = reinterpret_cast<D2*>(s_mat + 3 * (size_t)T.max_cells);
  D2* s_u = s_xy + T.max_verts;
  D2* s_v = s_u + T.max_verts;
  D2* s_a = s_v + T.max_verts;
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ FanCoef s_fc;
  const int tid = threadIdx.x;
  double* mb = E.members + (size_t)member * E.L.stride;
  const Layout& L = E.L;
  const MeshView& m = E.mesh;

  const int4* dsc = T.desc + 3 * blockIdx.x;
  const int4 d0 = __ldg(dsc), d1 = __ldg(dsc + 1);
  const int i0 = d0.x, nT = d0.y & 0xffff, nH = (int)((unsigned)d0.y >> 16);
  const int h0 = d0.z, ring0 = d0.w, rows = d1.x, tc0 = d1.y, ncp = d1.z, bbase = d1.w;
  const int nV = nT + nH;
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    const unsigned b_ring = (unsigned)rows * TN * (unsigned)sizeof(unsigned);
    const unsigned b_mat = (unsigned)ncp * 3u * (unsigned)sizeof(double);
    mbar_expect_tx(&s_bar, b_ring + b_mat);
    bulk_g2s(s_ring, T.ring + ring0, b_ring, &s_bar);
    bulk_g2s(s_mat, mat_m + (size_t)3 * tc0, b_mat, &s_bar);
  }
  const PropView pv = member_props<2>(E, mb);
  if (tid == TN - 1) s_fc = fan_coef(lame_fac(pv.scal[SC_NU]), prop_damping(pv), mix);
  const double* u1 = mb + L.off[VF_U1];
  const double* u0 = is_static ? u1 : mb + L.off[VF_U0];
  const double* v0 = mb + L.off[VF_V0];
  const double* a0 = mb + L.off[VF_A0];

  // ---- stage own + halo vertices: ids first, then every global load, then the stores --------
  constexpr int KV = 3;
  for (int base = 0; base < nV; base += KV * TN) {
    int vt[KV];
#pragma unroll
    for (int k = 0; k < KV; ++k) {
      const int t = base + k * TN + tid;
      vt[k] = t < nT ? i0 + t : (t < nV ? __ldg(T.halo + h0 + t - nT) : -1);
    }
    D2 c2[KV];
    NodeUVA s3[KV];
#pragma unroll
    for (int k = 0; k < KV; ++k) {
      if (vt[k] >= 0) {
        c2[k] = reinterpret_cast<const D2*>(m.xy)[vt[k]];
        if (RES) s3[k] = gather_node_uva(nc_arg, is_static != 0, vt[k], u1, u0, v0, a0);
      }
    }
#pragma unroll
    for (int k = 0; k < KV; ++k) {
      const int t = base + k * TN + tid;
      if (vt[k] >= 0) {
        s_xy[t] = c2[k];
        if (RES) {
          s_u[t] = s3[k].u;
          s_v[t] = s3[k].v;
          s_a[t] = s3[k].a;
        }
      }
    }
  }
  __syncthreads();
  mbar_wait(&s_bar, 0);

  // ---- walk: shared memory only ------------------------------------------------------------------
  if (tid < nT) {
    const FanCoef fc = s_fc;
    auto ring = [&](int r) { return s_ring[r * TN + tid]; };
    auto vtx_xy = [&](int s) { return s_xy[s]; };
    auto vtx_uva = [&](int s, D2& u, D2& v, D2& a) {
      u = s_u[s];
      v = s_v[s];
      a = s_a[s];
    };
    auto mat = [&](int c, double& emod, double& eta, double& rho) {
      emod = s_mat[c];
      eta = s_mat[ncp + c];
      rho = s_mat[2 * ncp + c];
    };
    double res[2];
    fan_walk_node<JAC, RES>(tid, ring, vtx_xy, vtx_uva, mat, fc, s_J, res);
    if (RES) reinterpret_cast<D2*>(mb + L.off[VF_F])[i0 + tid] = D2{res[0], res[1]};
  }
  if (JAC) {
    // the slice written through the generic proxy must be visible to the bulk-copy engine
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      bulk_s2g(mb + L.off[VF_J] + 4 * (size_t)bbase, s_J, (unsigned)dsc[2].x * 32u);
      bulk_commit();
    }
  }

  // ---- tail: pull the inputs of a later tile into L2 ---------------------------------------------
  const int far = blockIdx.x + pf_dist;
  if (pf_dist > 0 && far < (int)gridDim.x) {
    const int4 f0 = __ldg(T.desc + 3 * far), f1 = __ldg(T.desc + 3 * far + 1);
    const int f_i0 = f0.x, f_nT = f0.y & 0xffff, f_nH = (int)((unsigned)f0.y >> 16);
    auto pull = [&](const void* p, int nbytes) {
      const char* c = reinterpret_cast<const char*>(p);
      for (int off = tid * 128; off < nbytes; off += TN * 128) prefetch_l2(c + off);
    };
    pull(T.ring + f0.w, f1.x * TN * (int)sizeof(unsigned));
    pull(mat_m + (size_t)3 * f1.y, f1.z * 24);
    pull(m.xy + 2 * f_i0, 16 * f_nT);
    if (RES) {
      pull(u1 + 2 * f_i0, 16 * f_nT);
      if (!is_static) {
        pull(u0 + 2 * f_i0, 16 * f_nT);
        pull(v0 + 2 * f_i0, 16 * f_nT);
        pull(a0 + 2 * f_i0, 16 * f_nT);
      }
    }
    for (int h = tid; h < f_nH; h += TN) {
      const int v = __ldg(T.halo + f0.z + h);
      prefetch_l2(m.xy + 2 * v);
      if (RES) {
        prefetch_l2(u1 + 2 * v);
        if (!is_static) {
          prefetch_l2(u0 + 2 * v);
          prefetch_l2(v0 + 2 * v);
          prefetch_l2(a0 + 2 * v);
        }
      }
    }
  }
  // the CTA's shared memory must stay allocated until the bulk store has read it
  if (JAC && tid == 0) bulk_wait_read();
}

// ---- the same fan walk as a persistent, warp-specialised pipeline ---------------------------
// One CTA per SM, alive for the whole launch; tile k of CTA b is tile b + k * gridDim.x.
// Warpgroup 0 (keeps 56 registers per thread, the rest goes to the consumers: setmaxnreg):
//   warps 0-2   producer team.  Warp 0 owns a circular byte pool in shared memory: for every tile
//               it carves a stage of exactly the tile's size out of it (waiting, oldest first, for
//               consumed stages to be released: mbarrier `empty`).  The team then brings the
//               tile's inputs in WITHOUT touching registers: bulk asynchronous copies (TMA engine,
//               complete_tx on the stage's `full` mbarrier) for everything contiguous -- ring
//               table, tile-ordered cell properties, the own-vertex ranges of xy, u1, u0, v0, a0
//               -- and 16-byte cp.async gathers for the halo vertices
//               (cp.async.mbarrier.arrive.noinc ties them to the same mbarrier).  While those are
//               in flight it turns the raw nodal state of the PREVIOUS tile's stage into
//               (u1, v_nmk, a_nmk) in place and arrives on that stage's `ready` mbarrier.
//   warp 3      pulls the inputs of the tiles pf_dist rounds ahead into L2.
// Warpgroups 1 .. NG (152 registers per thread when NG = 3): consumer groups, one thread per node
// of a 128-node tile; group g takes tiles k = g, g + NG, ...  Every consumer WARP is on its own:
// it waits on `ready`, walks the fans of its 32 nodes out of shared memory into its private slice
// buffer, hands the slice (contiguous in the CSR array) to one bulk store and arrives on `empty`.
// Consumers never wait on global memory and never on each other.
// Stage: [ring | cell properties | xy | u1 | u0 | v0 -> v_nmk | a0 -> a_nmk], sized per tile.
constexpr int kPipeSlots = 8;  // tiles in flight per CTA (barrier slots)
constexpr int kRecSlots = 8;   // ring of producer records ...
constexpr int kRecAhead = 7;   // ... fetched this many tiles ahead of their use (DRAM latency)

template <bool JAC, bool RES, int NG, bool RAY>
__global__ void __launch_bounds__(128 + 128 * NG, 1) asm_fan_pipe_kernel(
    EngineDev E, int member, NewmarkCoef nc_arg, int is_static, JacMix mix, FanTablesDev T,
    const double* __restrict__ mat_m, int wj_bytes, int pool_bytes, int pf_dist, int dbg) {
  constexpr int TN = 128;
  extern __shared__ __align__(128) unsigned char pipe_smem[];
  __shared__ __align__(8) unsigned long long s_full[kPipeSlots], s_ready[kPipeSlots],
      s_empty[kPipeSlots], s_rec[kRecSlots];
  __shared__ int s_off[kPipeSlots];
  __shared__ int4 s_desc[kPipeSlots][3];  // the tile's descriptor, for the consumers
  __shared__ FanCoef s_fc;
  unsigned char* const s_recbuf = pipe_smem + (JAC ? (size_t)4 * NG * wj_bytes : 0);
  unsigned char* const pool = s_recbuf + (size_t)kRecSlots * T.prec_stride;
  double* mb = E.members + (size_t)member * E.L.stride;
  const Layout& L = E.L;
  const MeshView& m = E.mesh;
  const int ntiles = T.ntiles;
  const double* u1 = mb + L.off[VF_U1];
  const double* u0 = mb + L.off[VF_U0];
  const double* v0 = mb + L.off[VF_V0];
  const double* a0 = mb + L.off[VF_A0];
  const bool dyn = RES && !is_static;
  constexpr int kPlanes = RES ? 5 : 1;
  // producer team: warps 0-2 (warp 3 converts) or, team4, all of warpgroup 0 (consumers convert)
  const bool team4 = (dbg & 256) != 0;
  // VF_PIPE_DBG 512 (A/B variant, off by default): own vertices' u1, u0, v0, a0 loaded by the
  // consumer threads themselves (coalesced LDG issued BEFORE the wait for the stage, converted on
  // the fly) instead of four bulk copies
  const bool own_lsu = team4 && RES && (dbg & 512) != 0;
  const int kTeam = team4 ? 128 : 96;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kPipeSlots; ++s) {
      mbar_init_only(&s_full[s], kTeam / 32 + kTeam);  // one expect_tx arrival per team warp + one
                                                       // cp.async arrival per team thread
      mbar_init_only(&s_ready[s], 32);         // every converter thread
      mbar_init_only(&s_empty[s], 4);   // lane 0 of each consumer warp of the group
    }
    for (int s = 0; s < kRecSlots; ++s) mbar_init_only(&s_rec[s], 1);
    mbar_init_fence();
    const PropView pv = member_props<2>(E, mb);
    s_fc = fan_coef(lame_fac(pv.scal[SC_NU]), prop_damping(pv), mix);
  }
  __syncthreads();

  if (threadIdx.x < 128) {
    if (NG == 3) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if (threadIdx.x >= kTeam) {
      // ============================ converter: warp 3 ==============================================
      if (!RES) return;
      const int ct = (int)threadIdx.x - kTeam;
      for (int k = 0; k < my_tiles; ++k) {
        const int i = k & (kPipeSlots - 1);
        mbar_wait(&s_full[i], (unsigned)(k / kPipeSlots) & 1u);
        const int4 d0 = s_desc[i][0], d1 = s_desc[i][1];
        const int nV = (d0.y & 0xffff) + (int)((unsigned)d0.y >> 16);
        D2* s_u = reinterpret_cast<D2*>(pool + s_off[i] + d1.x * TN * (int)sizeof(unsigned) + d1.z * 24) + nV;
        D2* s_u0 = s_u + nV;
        D2* s_v = s_u0 + nV;
        D2* s_a = s_v + nV;
        if (is_static) {
          for (int t = ct; t < nV; t += 32) {
            s_v[t] = D2{0.0, 0.0};
            s_a[t] = D2{0.0, 0.0};
          }
        } else {
          // two vertices per pass: independent chains hide the fp64 latency of a lone warp
          for (int t = ct; t < nV; t += 64) {
            const int t2 = t + 32 < nV ? t + 32 : t;
            const D2 a_u = s_u[t], a_u0 = s_u0[t], a_v = s_v[t], a_a = s_a[t];
            const D2 b_u = s_u[t2], b_u0 = s_u0[t2], b_v = s_v[t2], b_a = s_a[t2];
            const NodeUVA ra = node_uva(nc_arg, false, a_u, a_u0, a_v, a_a);
            const NodeUVA rb = node_uva(nc_arg, false, b_u, b_u0, b_v, b_a);
            s_v[t] = ra.v;
            s_a[t] = ra.a;
            if (t2 != t) {
              s_v[t2] = rb.v;
              s_a[t2] = rb.a;
            }
          }
        }
        mbar_arrive(&s_ready[i]);
      }
      return;
    }
    // ================= producer team: warps 0-2 -- 96 threads in lockstep per tile ================
    const int tt = threadIdx.x, lane = tt & 31, role = tt >> 5;
    const int rstride = T.prec_stride;  // bytes
    // producer records (descriptor + halo vertex ids of a tile) arrive by bulk copies kRecAhead
    // tiles ahead of their use: no index load is ever on the critical path of the team
    auto fetch_record = [&](int k) {
      if (tt == 0 && k < my_tiles) {
        const size_t tile = (size_t)blockIdx.x + (size_t)k * gridDim.x;
        mbar_expect_tx(&s_rec[k & (kRecSlots - 1)], (unsigned)rstride);
        bulk_g2s(s_recbuf + (size_t)(k & (kRecSlots - 1)) * rstride, T.prec + tile * rstride,
                 (unsigned)rstride, &s_rec[k & (kRecSlots - 1)]);
      }
    };
    for (int k = 0; k < kRecAhead; ++k) fetch_record(k);
    int head = 0, tail_off = 0, tail = 0, inflight = 0;  // allocator state (warp 0 only)
    bool wrapped = false;
    for (int k = 0; k < my_tiles; ++k) {
      const int i = k & (kPipeSlots - 1);
      mbar_wait(&s_rec[k & (kRecSlots - 1)], (unsigned)(k / kRecSlots) & 1u);
      const int* rec = reinterpret_cast<const int*>(s_recbuf + (size_t)(k & (kRecSlots - 1)) * rstride);
      const int4 d0 = reinterpret_cast<const int4*>(rec)[0], d1 = reinterpret_cast<const int4*>(rec)[1];
      const int i0 = d0.x, nT = d0.y & 0xffff, nH = (int)((unsigned)d0.y >> 16);
      const int ring0 = d0.w, rows = d1.x, tc0 = d1.y, ncp = d1.z;
      const int nV = nT + nH;
      const int n_ring = rows * TN * (int)sizeof(unsigned), n_mat = ncp * 24, n_vec = nV * 16;
      if (role == 0) {
        // ---- carve the stage out of the circular pool; stages are released oldest first ----------
        const int sz = (n_ring + n_mat + kPlanes * n_vec + 127) & ~127;
        int off;
        for (;;) {
          if (inflight == 0) {
            off = 0;
            wrapped = false;
            break;
          }
          if (inflight < kPipeSlots) {
            if (!wrapped) {
              if (head + sz <= pool_bytes) {
                off = head;
                break;
              }
              if (sz <= tail_off) {
                off = 0;
                wrapped = true;
                break;
              }
            } else if (head + sz <= tail_off) {
              off = head;
              break;
            }
          }
          mbar_wait(&s_empty[tail & (kPipeSlots - 1)], (unsigned)(tail / kPipeSlots) & 1u);
          ++tail;
          --inflight;
          if (inflight) {
            const int t_new = s_off[tail & (kPipeSlots - 1)];
            if (t_new < tail_off) wrapped = false;  // the oldest stage is past the wrap point too
            tail_off = t_new;
          }
        }
        if (inflight == 0) tail_off = off;
        head = off + sz;
        ++inflight;
        if (lane == 0) {
          // published to the consumers by this thread's arrive.expect_tx (release) on `full`
          s_off[i] = off;
          s_desc[i][0] = d0;
          s_desc[i][1] = d1;
          s_desc[i][2] = reinterpret_cast<const int4*>(rec)[2];
        }
      }
      named_barrier(1, kTeam);
      // the record slot of tile k - 1 is free now (every thread is past its reads)
      fetch_record(k - 1 + kRecSlots);
      unsigned char* st = pool + s_off[i];
      D2* s_xy = reinterpret_cast<D2*>(st + n_ring + n_mat);
      D2* s_u1 = s_xy + nV;
      D2* s_u0 = s_u1 + nV;
      D2* s_v0 = s_u0 + nV;
      D2* s_a0 = s_v0 + nV;
      if (!(dbg & 128)) {
        // contiguous inputs as coalesced 16-byte cp.async (LSU path, 512 bytes per warp
        // instruction): measured, the bulk-copy engine moves ~16 B/clk per SM in both directions
        // together, and the CSR slices going out already keep it busy (profiles/README.md)
        auto copy16 = [&](void* dst, const void* src, int nbytes) {
          for (int o = tt * 16; o < nbytes; o += kTeam * 16)
            cp_async16(reinterpret_cast<unsigned char*>(dst) + o,
                       reinterpret_cast<const unsigned char*>(src) + o);
        };
        const int n_own = nT * 16;
        copy16(st, T.ring + ring0, n_ring);
        copy16(st + n_ring, mat_m + (size_t)3 * tc0, n_mat);
        copy16(s_xy, m.xy + 2 * (size_t)i0, n_own);
        if (RES) {
          copy16(s_u1, u1 + 2 * (size_t)i0, n_own);
          if (dyn) {
            copy16(s_u0, u0 + 2 * (size_t)i0, n_own);
            copy16(s_v0, v0 + 2 * (size_t)i0, n_own);
            copy16(s_a0, a0 + 2 * (size_t)i0, n_own);
          }
        }
        if (lane == 0) mbar_arrive(&s_full[i]);  // stands in for the expect_tx arrival
      } else
      if (lane == 0 && !(dbg & 6)) {
        // VF_PIPE_DBG 128: the seven bulk copies of the tile (bulk-copy engine).  Issuing one
        // costs the issuing lane ~150-200 cycles, so they are spread over the team's warps
        // (four warps: at most three each, counting the record fetch of warp 0)
        const unsigned n_own = (unsigned)nT * 16u;
        const bool st8 = RES && !own_lsu;       // own-vertex state by bulk copies
        const bool dy8 = st8 && dyn;
        if (role == 0) {
          mbar_expect_tx(&s_full[i], (unsigned)(n_ring + n_mat));
          bulk_g2s(st, T.ring + ring0, (unsigned)n_ring, &s_full[i]);
          bulk_g2s(st + n_ring, mat_m + (size_t)3 * tc0, (unsigned)n_mat, &s_full[i]);
        } else if (role == 1) {
          const bool with_u0 = dy8 && !team4;
          mbar_expect_tx(&s_full[i], n_own * (1u + (st8 ? 1u : 0u) + (with_u0 ? 1u : 0u)));
          bulk_g2s(s_xy, m.xy + 2 * (size_t)i0, n_own, &s_full[i]);
          if (st8) bulk_g2s(s_u1, u1 + 2 * (size_t)i0, n_own, &s_full[i]);
          if (with_u0) bulk_g2s(s_u0, u0 + 2 * (size_t)i0, n_own, &s_full[i]);
        } else if (role == 2) {
          // four warps: u0, v0 here and a0 on warp 3; three warps: v0, a0 here
          mbar_expect_tx(&s_full[i], dy8 ? 2u * n_own : 0u);
          if (dy8) {
            if (team4) bulk_g2s(s_u0, u0 + 2 * (size_t)i0, n_own, &s_full[i]);
            bulk_g2s(s_v0, v0 + 2 * (size_t)i0, n_own, &s_full[i]);
            if (!team4) bulk_g2s(s_a0, a0 + 2 * (size_t)i0, n_own, &s_full[i]);
          }
        } else {
          mbar_expect_tx(&s_full[i], dy8 ? n_own : 0u);
          if (dy8) bulk_g2s(s_a0, a0 + 2 * (size_t)i0, n_own, &s_full[i]);
        }
      } else if (lane == 0) {
        mbar_expect_tx(&s_full[i], 0u);  // VF_PIPE_DBG 2 / 4 (measurement aid): no bulk copies
      }
      __syncwarp();
      // one halo vertex per thread: 16-byte asynchronous gathers
      for (int h = tt; h < nH && !(dbg & 1); h += kTeam) {
        const size_t v = (size_t)rec[16 + h];
        cp_async16(s_xy + nT + h, m.xy + 2 * v);
        if (RES) {
          cp_async16(s_u1 + nT + h, u1 + 2 * v);
          if (dyn) {
            cp_async16(s_u0 + nT + h, u0 + 2 * v);
            cp_async16(s_v0 + nT + h, v0 + 2 * v);
            cp_async16(s_a0 + nT + h, a0 + 2 * v);
          }
        }
      }
      cp_async_mbar_arrive_noinc(&s_full[i]);
      // pull the inputs of the tile pf_dist rounds ahead into L2 (its record is already here)
      if (pf_dist > 0 && k + pf_dist < my_tiles) {
        const int kf = k + pf_dist;
        const bool pf_halo = !(dbg & 32);  // VF_PIPE_DBG 32: no prefetch of the halo vertex lines
        mbar_wait(&s_rec[kf & (kRecSlots - 1)], (unsigned)(kf / kRecSlots) & 1u);
        const int* fr = reinterpret_cast<const int*>(s_recbuf + (size_t)(kf & (kRecSlots - 1)) * rstride);
        const int4 f0 = reinterpret_cast<const int4*>(fr)[0], f1 = reinterpret_cast<const int4*>(fr)[1];
        const int f_i0 = f0.x, f_nT = f0.y & 0xffff, f_nH = (int)((unsigned)f0.y >> 16);
        if (lane == 0) {
          // contiguous inputs: one bulk prefetch each, split like the copies
          const unsigned f_own = (unsigned)f_nT * 16u;
          if (role == 0) {
            bulk_prefetch_l2(T.ring + f0.w, (unsigned)f1.x * TN * (unsigned)sizeof(unsigned));
            if (!team4) bulk_prefetch_l2(mat_m + (size_t)3 * f1.y, (unsigned)f1.z * 24u);
          } else if (role == 1) {
            bulk_prefetch_l2(m.xy + 2 * (size_t)f_i0, f_own);
            if (RES) bulk_prefetch_l2(u1 + 2 * (size_t)f_i0, f_own);
            if (dyn && !team4) bulk_prefetch_l2(u0 + 2 * (size_t)f_i0, f_own);
          } else if (role == 2) {
            if (dyn && team4) bulk_prefetch_l2(u0 + 2 * (size_t)f_i0, f_own);
            if (dyn) bulk_prefetch_l2(v0 + 2 * (size_t)f_i0, f_own);
            if (dyn && !team4) bulk_prefetch_l2(a0 + 2 * (size_t)f_i0, f_own);
          } else {
            bulk_prefetch_l2(mat_m + (size_t)3 * f1.y, (unsigned)f1.z * 24u);
            if (dyn) bulk_prefetch_l2(a0 + 2 * (size_t)f_i0, f_own);
          }
        }
        __syncwarp();
        if (pf_halo)
        for (int h = tt; h < f_nH; h += kTeam) {
          const size_t v = (size_t)fr[16 + h];
          prefetch_l2(m.xy + 2 * v);
          if (RES) {
            prefetch_l2(u1 + 2 * v);
            if (dyn) {
              prefetch_l2(u0 + 2 * v);
              prefetch_l2(v0 + 2 * v);
              prefetch_l2(a0 + 2 * v);
            }
          }
        }
      }
    }
    return;
  }

  // ================================== consumer warps ===============================================
  if (NG == 3) asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
  const int g = ((int)threadIdx.x - 128) / TN;
  const int tid = ((int)threadIdx.x - 128) % TN;
  const int lane = tid & 31, w0 = tid & ~31;
  double* s_Jw = reinterpret_cast<double*>(pipe_smem + (size_t)(4 * g + (tid >> 5)) * wj_bytes);
#ifdef VF_PIPE_PROF
  long long ck_wait = 0, ck_jwait = 0, ck_walk = 0, ck_t0 = clock64();
#endif
  for (int k = g;; k += NG) {
    const long long tile_ll = (long long)blockIdx.x + (long long)k * gridDim.x;
    if (tile_ll >= ntiles) break;
    const int i = k & (kPipeSlots - 1);
    const unsigned ph = (unsigned)(k / kPipeSlots) & 1u;
#ifdef VF_PIPE_PROF
    const long long c0 = clock64();
#endif
    // own vertex of this thread: its nodal state is requested before waiting for the stage
    D2 o_u1 = D2{0.0, 0.0}, o_u0 = o_u1, o_v0 = o_u1, o_a0 = o_u1;
    if (own_lsu) {
      const long long vtx = tile_ll * TN + tid;
      if (vtx < m.nn) {
        o_u1 = reinterpret_cast<const D2*>(u1)[vtx];
        if (!is_static) {
          o_u0 = reinterpret_cast<const D2*>(u0)[vtx];
          o_v0 = reinterpret_cast<const D2*>(v0)[vtx];
          o_a0 = reinterpret_cast<const D2*>(a0)[vtx];
        }
      }
    }
    mbar_wait(RES && !team4 ? &s_ready[i] : &s_full[i], ph);
#ifdef VF_PIPE_PROF
    const long long c1 = clock64();
    ck_wait += c1 - c0;
#endif
    const int4 d0 = s_desc[i][0], d1 = s_desc[i][1];
    const int i0 = d0.x, nT = d0.y & 0xffff, nH = (int)((unsigned)d0.y >> 16);
    const int ncp = d1.z, bbase = d1.w;
    const int nV = nT + nH;
    const int nblk = JAC ? s_desc[i][2].x : 0;
    const unsigned char* st = pool + s_off[i];
    const unsigned* s_ring = reinterpret_cast<const unsigned*>(st);
    const double* s_mat = reinterpret_cast<const double*>(st + d1.x * TN * (int)sizeof(unsigned));
    const D2* s_xy = reinterpret_cast<const D2*>(s_mat + 3 * ncp);
    D2* s_u = const_cast<D2*>(s_xy) + nV;
    D2* s_v = s_u + 2 * nV;
    D2* s_a = s_v + nV;
    if (RES && team4) {
      // raw nodal state -> (u1, v_nmk, a_nmk), once per staged vertex, in place, by the group
      const D2* s_u0 = s_u + nV;
      if (own_lsu && tid < nT) {
        const NodeUVA r = node_uva(nc_arg, is_static != 0, o_u1, o_u0, o_v0, o_a0);
        s_u[tid] = r.u;
        s_v[tid] = r.v;
        s_a[tid] = r.a;
      }
      for (int t = own_lsu ? nT + tid : tid; t < nV; t += TN) {
        if (is_static) {
          s_v[t] = D2{0.0, 0.0};
          s_a[t] = D2{0.0, 0.0};
        } else {
          const NodeUVA r = node_uva(nc_arg, false, s_u[t], s_u0[t], s_v[t], s_a[t]);
          s_v[t] = r.v;
          s_a[t] = r.a;
        }
      }
      named_barrier(2 + g, TN);
    }
    // this warp's slice of the tile's CSR values: blocks [b_lo, b_hi) relative to the tile
    int b_lo = 0, b_hi = 0;
    if (JAC && w0 < nT) {
      b_lo = fan_hdr_b0(s_ring[w0]);
      b_hi = w0 + 32 < nT ? fan_hdr_b0(s_ring[w0 + 32]) : nblk;
      // the slice buffer is free once the bulk store of the previous tile has read it
      if (lane == 0 && !(dbg & 64)) bulk_wait_read();
      __syncwarp();
    }
#ifdef VF_PIPE_PROF
    const long long c2 = clock64();
    ck_jwait += c2 - c1;
#endif
    if (tid < nT && !(dbg & 8)) {
      const FanCoef& fc = s_fc;
      auto ring = [&](int r) { return s_ring[r * TN + tid]; };
      auto vtx_xy = [&](int v) { return s_xy[v]; };
      auto vtx_uva = [&](int v, D2& pu, D2& pv, D2& pa) {
        pu = s_u[v];
        pv = s_v[v];
        pa = s_a[v];
      };
      auto mat = [&](int c, double& emod, double& eta, double& rho) {
        emod = s_mat[c];
        eta = s_mat[ncp + c];
        rho = s_mat[2 * ncp + c];
      };
      double res[2];
      fan_walk_node<JAC, RES, RAY>(tid, ring, vtx_xy, vtx_uva, mat, fc, s_Jw - 4 * (ptrdiff_t)b_lo,
                                   res);
      if (RES) reinterpret_cast<D2*>(mb + L.off[VF_F])[i0 + tid] = D2{res[0], res[1]};
    }
    // generic-proxy writes (slice, converted state) are read / overwritten by the bulk-copy engine
    fence_proxy_async_smem();
    __syncwarp();
#ifdef VF_PIPE_PROF
    ck_walk += clock64() - c2;
#endif
    if (JAC && (dbg & 64) && b_hi > b_lo) {
      // VF_PIPE_DBG 64: the warp streams its slice out itself (coalesced 16-byte stores)
      const double2* src = reinterpret_cast<const double2*>(s_Jw);
      double2* dst = reinterpret_cast<double2*>(mb + L.off[VF_J] + 4 * ((size_t)bbase + b_lo));
      const int n2 = 2 * (b_hi - b_lo);
      for (int o = lane; o < n2; o += 32) __stcs(dst + o, src[o]);
      __syncwarp();
    }
    if (lane == 0) {
      if (JAC && b_hi > b_lo && !(dbg & (16 | 64))) {
        bulk_s2g(mb + L.off[VF_J] + 4 * ((size_t)bbase + b_lo), s_Jw, (unsigned)(b_hi - b_lo) * 32u);
        bulk_commit();
      }
      mbar_arrive(&s_empty[i]);
    }
  }
  // shared memory must stay allocated until the last bulk store has read it
  if (JAC && lane == 0) bulk_wait_read();
#ifdef VF_PIPE_PROF
  if (lane == 0) {
    double* info = mb + L.off[VF_INFO];
    atomicAdd(info + 12, (double)(clock64() - ck_t0));  // consumer warp: loop total
    atomicAdd(info + 13, (double)ck_wait);              //   waiting for a ready stage
    atomicAdd(info + 14, (double)ck_jwait);             //   waiting for the slice buffer
    atomicAdd(info + 15, (double)ck_walk);              //   walking
  }
#endif
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


// The same terms from the boundary-node records (engine_internal.h FacetRec2D): triangles with
// follower pressure and Dirichlet rows only.  facet_bc_kernel walks brptr -> n2f -> pf_cell ->
// cells -> coordinates before it can touch the state: six dependent loads from tables that the
// 0.8 GB stream of the assembly has just evicted (19 us for 10 k nodes).  Here one record load is
// followed by all nodal loads at once, then the row blocks.
template <bool JAC, bool RES>
__global__ void facet_bc_fast_kernel(EngineDev E, int member, JacMix mix,
                                     const FacetRec2D* __restrict__ recs, int n_touch) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_touch) return;
  const FacetRec2D r = recs[t];
  double* mb = E.members + (size_t)member * E.L.stride;
  const double* u1 = mb + E.L.off[VF_U1];
  const double* p1 = mb + E.L.off[VF_P1];
  double* F = mb + E.L.off[VF_F];
  double* rowblk = mb + E.L.off[VF_J] + (size_t)4 * r.b0;
  const int ld = 2 * r.deg;
  // every nodal value of both facets first
  D2 U[2][3];
  double P[2][3];
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const bool on = k < r.nfac;
      U[k][b] = on ? reinterpret_cast<const D2*>(u1)[r.f[k].nd[b]] : D2{0.0, 0.0};
      P[k][b] = on ? p1[r.f[k].nd[b]] : 0.0;
    }
  // launched with programmatic stream serialisation: everything above (record, nodal values:
  // inputs of the assembly) may run while the assembly kernel drains; its outputs (F, J) are
  // touched only after this point
  asm volatile("griddepcontrol.wait;" ::: "memory");
  double res[2] = {0.0, 0.0};
  if (RES) {
    const D2 f2 = reinterpret_cast<const D2*>(F)[r.node];
    res[0] = f2.x;
    res[1] = f2.y;
  }
  for (int k = 0; k < r.nfac; ++k) {
    const FacetRec2D::Facet& q = r.f[k];
    if (q.a == q.o) continue;
    double gu[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        double sacc = 0.0;
#pragma unroll
        for (int b = 0; b < 3; ++b) sacc += (i == 0 ? U[k][b].x : U[k][b].y) * q.G[b][j];
        gu[i][j] = sacc;
      }
    const double mw = q.meas / 6.0;  // facet mass weight: mw (1 + delta_ab)
    double pw = 0.0;
#pragma unroll
    for (int b = 0; b < 3; ++b)
      if (b != q.o) pw += (q.a == b ? 2.0 : 1.0) * P[k][b];
    pw *= mw;
    const double N[2] = {q.N[0], q.N[1]};
    if (RES) {
      double c[2];
      cof_normal(gu, N, c);
      res[0] += pw * c[0];
      res[1] += pw * c[1];
    }
    if (JAC) {
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        double dc[2][2];
        dcof_normal(gu, N, q.G[b], dc);
        add_block<2>(rowblk, ld, q.slot[b], dc, pw * mix.p);
      }
    }
  }
  // Dirichlet rows: zero row, unit diagonal, zero residual (App. A.4)
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    if (r.bc & (1 << a)) {
      if (JAC && mix.bc) {
        for (int c = 0; c < ld; ++c) rowblk[a * ld + c] = 0.0;
        rowblk[a * ld + r.self * 2 + a] = 1.0;
      }
      if (RES) res[a] = 0.0;
    }
  }
  if (RES) reinterpret_cast<D2*>(F)[r.node] = D2{res[0], res[1]};
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
  assemble_node_auto<D, JAC, RES>(
      i, E.mesh, pv, sv, JAC ? mb + L.off[VF_J] + (size_t)D * D * E.mesh.brptr[i] : nullptr, res);
  if (RES) {
    double* F = mb + L.off[VF_F];
#pragma unroll
    for (int c = 0; c < D; ++c) F[D * i + c] = res[c];
  }
}


// The same thread-per-node gather with the block rows of one warp's 32 consecutive nodes
// accumulated in shared memory, packed exactly like the CSR array (rows of consecutive nodes are
// contiguous there), and streamed out by ONE coalesced copy.  The ~800 read-modify-writes a
// tetrahedral node makes on its row (24 cells x 4 blocks x 9 entries) then stay in shared memory
// instead of going through to L2 one partial sector at a time.  One warp per CTA (no block
// barrier; the CTA's shared memory is the largest group of 32 rows, e->node_warp_blocks), several
// CTAs per SM.  Same visiting order per entry as the global variant, so the two produce identical
// bits.  Measured (0.99 M tets, profiles/r2_variants_ab.json): the rows alone gain 4 %, the gather
// tables of tet_tables.h alone 20 %, the two together 2.6x (1.98 -> 0.758 ms): each removes one of
// two limits that cap the kernel at about the same time.
template <int D, bool RES>
__global__ void __launch_bounds__(32)
asm_node_warp_kernel(EngineDev E, int member, double dt, int is_static, JacMix mix) {
  extern __shared__ double s_rows[];
  const int nn = E.mesh.nn;
  const int i0 = blockIdx.x * 32;
  const int i = i0 + threadIdx.x;
  const int bfirst = E.mesh.brptr[i0];
  const int blast = E.mesh.brptr[min(i0 + 32, nn)];
  double* mb = E.members + (size_t)member * E.L.stride;
  const Layout& L = E.L;
  if (i < nn) {
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
    assemble_node_auto<D, true, RES>(i, E.mesh, pv, sv,
                                     s_rows + (size_t)D * D * (E.mesh.brptr[i] - bfirst), res);
    if (RES) {
      double* F = mb + L.off[VF_F];
#pragma unroll
      for (int c = 0; c < D; ++c) F[D * i + c] = res[c];
    }
  }
  __syncwarp();
  double* __restrict__ out = mb + L.off[VF_J] + (size_t)D * D * bfirst;
  const int total = D * D * (blast - bfirst);
  for (int t = threadIdx.x; t < total; t += 32) out[t] = s_rows[t];
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


}  // namespace vf

using namespace vf;

namespace {


// dynamic shared memory of asm_tile2_kernel: records, region A (CSR slice aliasing the nodal
// staging: coordinates + u/v/a = 64 bytes per own or halo vertex), F, index slices
size_t tile2_smem_bytes(const vf_problem_desc& d, bool direct = false) {
  const size_t idx_words = (size_t)d.max_tile_pairs + 2 * ((size_t)d.tile_threads + 1);
  const size_t region_a = direct ? (size_t)8 * d.max_tile_verts
                                 : std::max((size_t)d.tile_max_values, (size_t)8 * d.max_tile_verts);
  return sizeof(double) * ((size_t)d.max_tile_elems * kRec2D + region_a +
                           2 * (size_t)d.tile_threads + ((idx_words + 3) / 4) * 2);
}


}  // namespace

int vf::assembly_configure(vf_engine* e, const vf_problem_desc& d, bool two_phase) {
  // tetrahedra: gather-friendly copies of the mesh tables + precomputed CSR slots (tet_tables.h)
  if (d.dim == 3 && d.nn > 0 && d.ne > 0) {
    std::vector<int32_t> cells4;
    std::vector<double> xyz4;
    std::vector<uint32_t> slots;
    if (build_tet_gather_tables(d.nn, d.ne, d.xyz_host, d.cells_host, e->brptr.data(),
                                e->bcol.data(), d.n2e_ptr_host, d.n2e_host, cells4, xyz4, slots)) {
      const size_t b_xyz = align_up(xyz4.size() * sizeof(double), 256);
      const size_t b_cells = align_up(cells4.size() * sizeof(int32_t), 256);
      const size_t b_slots = align_up(std::max<size_t>(slots.size(), 1) * sizeof(uint32_t), 256);
      char* mem = nullptr;
      VF_CUDA(cudaMalloc(&mem, b_xyz + b_cells + b_slots));
      e->node_mem = mem;
      VF_CUDA(cudaMemcpy(mem, xyz4.data(), xyz4.size() * sizeof(double), cudaMemcpyHostToDevice));
      VF_CUDA(cudaMemcpy(mem + b_xyz, cells4.data(), cells4.size() * sizeof(int32_t),
                         cudaMemcpyHostToDevice));
      VF_CUDA(cudaMemcpy(mem + b_xyz + b_cells, slots.data(), slots.size() * sizeof(uint32_t),
                         cudaMemcpyHostToDevice));
      e->dev.mesh.xyz4 = reinterpret_cast<const double*>(mem);
      e->dev.mesh.cells4 = reinterpret_cast<const int*>(mem + b_xyz);
      e->dev.mesh.n2e_slots = reinterpret_cast<const unsigned*>(mem + b_xyz + b_cells);
    }
  }
  // asm_node_warp_kernel: most CSR blocks owned by 32 consecutive nodes = its shared memory
  e->node_warp_blocks = 0;
  for (int n = 0; n < d.nn; n += 32)
    e->node_warp_blocks =
        std::max(e->node_warp_blocks, e->brptr[std::min(n + 32, d.nn)] - e->brptr[n]);
  {
    const size_t smem = sizeof(double) * d.dim * d.dim * (size_t)e->node_warp_blocks;
    if (smem <= (size_t)kNodeWarpMaxSmem) {
      const int sm = (int)smem;
      if (d.dim == 3) {
        VF_CUDA(cudaFuncSetAttribute(asm_node_warp_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        VF_CUDA(cudaFuncSetAttribute(asm_node_warp_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
      } else {
        VF_CUDA(cudaFuncSetAttribute(asm_node_warp_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        VF_CUDA(cudaFuncSetAttribute(asm_node_warp_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
      }
    }
  }
  if (two_phase) {
    const int smem2 = (int)tile2_smem_bytes(d);
    if (smem2 > 227 * 1024) {
      return fail("two-phase tile exceeds the 227 KB shared memory of an SM");
    }
#define VF_SMEM2(K) VF_CUDA(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2))
#define VF_SMEM2_ALL(J_, R_, ROW_)                          \
  VF_SMEM2((asm_tile2_kernel<J_, R_, ROW_, 128, 8>));       \
  VF_SMEM2((asm_tile2_kernel<J_, R_, ROW_, 192, 5>));       \
  VF_SMEM2((asm_tile2_kernel<J_, R_, ROW_, 256, 4>));       \
  VF_SMEM2((asm_tile2_kernel<J_, R_, ROW_, 320, 3>))
    VF_SMEM2_ALL(true, true, 1);
    VF_SMEM2_ALL(true, false, 1);
    VF_SMEM2_ALL(false, true, 1);
#undef VF_SMEM2_ALL
#undef VF_SMEM2
  }
  return 0;
}

extern "C" {

int vf_set_fan_tables(vf_engine* e, int tile_nodes, int ntiles, const int32_t* desc_host,
                      const uint32_t* ring_host, size_t n_ring, const int32_t* halo_host,
                      size_t n_halo, const int32_t* tcell_host, size_t n_tcell, int max_verts,
                      int max_rows, int max_cells, int max_blocks, void* stream) {
  if (!e) return fail("null engine");
  if (e->desc.dim != 2 || !e->fan_ok) return fail("fan tables need triangles with ordered fans");
  if (tile_nodes != 64 && tile_nodes != 96 && tile_nodes != 128)
    return fail("fan tile_nodes must be 64, 96 or 128");
  if (ntiles <= 0 || !desc_host || !ring_host || !halo_host || !tcell_host || n_tcell == 0)
    return fail("missing fan tables");
  if ((size_t)ntiles * tile_nodes < (size_t)e->desc.nn) return fail("fan tiles do not cover the mesh");
  if (max_cells % 2) return fail("fan tile cell lists must be padded to an even count");
  cudaStream_t st = as_stream(stream);
  const size_t b_desc = align_up(sizeof(int32_t) * 12 * (size_t)ntiles, 256);
  const size_t b_ring = align_up(sizeof(uint32_t) * n_ring, 256);
  const size_t b_halo = align_up(sizeof(int32_t) * std::max<size_t>(n_halo, 1), 256);
  const size_t b_tcell = align_up(sizeof(int32_t) * n_tcell, 256);
  const size_t b_mat = sizeof(double) * 3 * n_tcell * (size_t)e->desc.n_members;
  // producer records of the pipelined kernel: [descriptor (12 int32) | 4 pad | halo vertex ids],
  // one fixed stride per tile, so that one bulk copy brings a tile's control data on chip
  int max_nh = 0;
  for (int t = 0; t < ntiles; ++t)
    max_nh = std::max(max_nh, (int)((uint32_t)desc_host[12 * (size_t)t + 1] >> 16));
  const int prec_stride = 64 + 4 * ((max_nh + 3) / 4 * 4);
  std::vector<int32_t> prec((size_t)ntiles * (prec_stride / 4), 0);
  for (int t = 0; t < ntiles; ++t) {
    int32_t* r = prec.data() + (size_t)t * (prec_stride / 4);
    const int32_t* d = desc_host + 12 * (size_t)t;
    std::copy(d, d + 12, r);
    const int nh = (int)((uint32_t)d[1] >> 16);
    if ((size_t)d[2] + nh > std::max<size_t>(n_halo, 1)) return fail("fan halo list out of range");
    std::copy(halo_host + d[2], halo_host + d[2] + nh, r + 16);
  }
  const size_t b_prec = align_up(prec.size() * sizeof(int32_t), 256);
  if (e->fan_mem) {
    cudaFree(e->fan_mem);
    e->fan_mem = nullptr;
    e->fan.ring = nullptr;
  }
  char* mem = nullptr;
  VF_CUDA(cudaMalloc(&mem, b_desc + b_ring + b_halo + b_tcell + b_prec + b_mat));
  e->fan_mem = mem;
  VF_CUDA(cudaMemcpyAsync(mem, desc_host, sizeof(int32_t) * 12 * (size_t)ntiles,
                          cudaMemcpyHostToDevice, st));
  VF_CUDA(cudaMemcpyAsync(mem + b_desc, ring_host, sizeof(uint32_t) * n_ring,
                          cudaMemcpyHostToDevice, st));
  if (n_halo)
    VF_CUDA(cudaMemcpyAsync(mem + b_desc + b_ring, halo_host, sizeof(int32_t) * n_halo,
                            cudaMemcpyHostToDevice, st));
  VF_CUDA(cudaMemcpyAsync(mem + b_desc + b_ring + b_halo, tcell_host, sizeof(int32_t) * n_tcell,
                          cudaMemcpyHostToDevice, st));
  VF_CUDA(cudaMemcpyAsync(mem + b_desc + b_ring + b_halo + b_tcell, prec.data(),
                          prec.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  VF_CUDA(cudaStreamSynchronize(st));
  FanTablesDev& T = e->fan;
  T.desc = reinterpret_cast<const int4*>(mem);
  T.ring = reinterpret_cast<const unsigned*>(mem + b_desc);
  T.halo = reinterpret_cast<const int*>(mem + b_desc + b_ring);
  T.tcell = reinterpret_cast<const int*>(mem + b_desc + b_ring + b_halo);
  T.prec = reinterpret_cast<const unsigned char*>(mem + b_desc + b_ring + b_halo + b_tcell);
  T.prec_stride = prec_stride;
  T.mat = reinterpret_cast<double*>(mem + b_desc + b_ring + b_halo + b_tcell + b_prec);
  T.n_tcell = n_tcell;
  T.tile_nodes = tile_nodes;
  T.ntiles = ntiles;
  T.max_verts = max_verts;
  T.max_rows = max_rows;
  T.max_cells = max_cells;
  T.max_blocks = max_blocks;
  e->fan_dirty.assign((size_t)e->desc.n_members, 1);
  // largest number of CSR blocks owned by 32 consecutive nodes (a consumer warp of the pipeline)
  e->fan_max_wblocks = 0;
  if (e->brptr.size() == (size_t)e->desc.nn + 1) {
    const int nn = e->desc.nn;
    for (int n = 0; n < nn; n += 32)
      e->fan_max_wblocks = std::max(e->fan_max_wblocks, e->brptr[std::min(n + 32, nn)] - e->brptr[n]);
  }
  return 0;
}

int vf_props_changed(vf_engine* e, int member) {
  if (!e) return fail("null engine");
  if (member >= e->desc.n_members) return fail("member out of range");
  if (member < 0) std::fill(e->fan_dirty.begin(), e->fan_dirty.end(), 1);
  else if (!e->fan_dirty.empty()) e->fan_dirty[member] = 1;
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
// Node-centric fan kernel (triangles, ordered fans): the default 2D path once its tables are set.
size_t fan_smem_bytes(const FanTablesDev& T, bool jac, bool res) {
  return (jac ? 32 * (size_t)T.max_blocks : 0) + sizeof(unsigned) * (size_t)T.max_rows * T.tile_nodes +
         24 * (size_t)T.max_cells + sizeof(D2) * (size_t)T.max_verts * (res ? 4 : 1);
}

int launch_fan(vf_engine* e, int member, bool res, bool jac, double dt, int is_static,
               const JacMix& mix, cudaStream_t st) {
  const FanTablesDev& T = e->fan;
  double* mat_m = T.mat + (size_t)member * 3 * T.n_tcell;
  if (e->fan_dirty[member]) {
    // the tile-ordered copy of emod / eta / rho follows the member's properties
    fan_pack_kernel<<<T.ntiles, 128, 0, st>>>(e->dev, member, T, mat_m);
    e->fan_dirty[member] = 0;
    e->launches += 1;
    VF_CUDA(cudaGetLastError());
  }
  const char* env_pf = getenv("VF_PF_DIST");
  const char* env_mb = getenv("VF_FAN_MINB");
  const int minb_env = env_mb ? atoi(env_mb) : 0;
  const NewmarkCoef nc = newmark_coef(dt);
  const size_t smem = fan_smem_bytes(T, jac, res);
  if (smem > 227 * 1024) return fail("fan tile exceeds the 227 KB shared memory of an SM");
#define VF_FAN_GO(J_, R_, TN_, MB_)                                                               \
  do {                                                                                            \
    const int pf_dist = env_pf ? atoi(env_pf) : 148 * (MB_);                                      \
    VF_CUDA(cudaFuncSetAttribute(asm_fan_kernel<J_, R_, TN_, MB_>,                                \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
    asm_fan_kernel<J_, R_, TN_, MB_><<<T.ntiles, TN_, smem, st>>>(e->dev, member, nc, is_static,  \
                                                                   mix, T, mat_m, pf_dist);       \
  } while (0)
#define VF_FAN_BY_MODE(TN_, MB_)                                                                  \
  do {                                                                                            \
    if (jac && res) VF_FAN_GO(true, true, TN_, MB_);                                              \
    else if (jac) VF_FAN_GO(true, false, TN_, MB_);                                               \
    else VF_FAN_GO(false, true, TN_, MB_);                                                        \
  } while (0)
  if (T.tile_nodes == 64) VF_FAN_BY_MODE(64, 7);
  else if (T.tile_nodes == 96) VF_FAN_BY_MODE(96, 5);
  else if (minb_env == 4) VF_FAN_BY_MODE(128, 4);
  else VF_FAN_BY_MODE(128, 3);
#undef VF_FAN_BY_MODE
#undef VF_FAN_GO
  e->launches += 1;
  VF_CUDA(cudaGetLastError());
  return 0;
}

// Persistent producer / consumer pipeline (asm_fan_pipe_kernel): the default when the tile is
// 128 nodes and the slice buffers of the consumer warps leave a pool of at least two of the
// largest stages in the 227 KB of an SM.
size_t fan_pipe_stage_bytes(const FanTablesDev& T, bool res) {
  return sizeof(unsigned) * (size_t)T.max_rows * 128 + 24 * (size_t)T.max_cells +
         sizeof(D2) * (size_t)T.max_verts * (res ? 5 : 1) + 128;
}

// returns 0 when launched, -1 when the configuration does not fit (caller falls back), 1 on error
int launch_fan_pipe(vf_engine* e, int member, bool res, bool jac, double dt, int is_static,
                    const JacMix& mix, cudaStream_t st) {
  const FanTablesDev& T = e->fan;
  if (T.tile_nodes != 128 || e->fan_max_wblocks <= 0) return -1;
  const char* env_ng = getenv("VF_PIPE_GROUPS");
  const char* env_pool = getenv("VF_PIPE_POOL_KB");
  const char* env_pf = getenv("VF_PIPE_PF");
  const char* env_grid = getenv("VF_PIPE_GRID");
  const size_t b_stage = fan_pipe_stage_bytes(T, res);
  const size_t wj = jac ? align_up(32 * (size_t)e->fan_max_wblocks, 128) : 0;
  // static shared memory of the kernel (barriers, FanCoef) and its ring of producer records
  const size_t cap = 227 * 1024 - 1024 - (size_t)kRecSlots * T.prec_stride;
  int ng = env_ng ? atoi(env_ng) : 3;
  if (ng != 2 && ng != 3) ng = 3;
  // three of the largest stages: with fewer, the circular pool could have room for the next tile
  // only where the previous, not yet converted one sits (the producer team would wait forever)
  while (ng > 2 && 4 * ng * wj + 3 * b_stage > cap) --ng;
  if (4 * ng * wj + 3 * b_stage > cap) return -1;
  size_t pool = cap - 4 * ng * wj;
  if (env_pool) pool = std::min(pool, std::max(3 * b_stage, (size_t)atoi(env_pool) * 1024));
  pool &= ~(size_t)127;
  const size_t smem = 4 * ng * wj + (size_t)kRecSlots * T.prec_stride + pool;
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  int grid = env_grid ? atoi(env_grid) : n_sm;
  grid = std::max(1, std::min(grid, T.ntiles));
  const int pf_dist = env_pf ? atoi(env_pf) : 1;
  // VF_PIPE_DBG: variants kept for A/B timing (profiles/README.md) and measurement aids.
  //   32 also prefetch the halo vertex lines, 64 consumer warps store their slices themselves,
  //   128 contiguous inputs by cp.async instead of the bulk-copy engine, 256 three producer warps
  //   + one converter warp (default: four producer warps, consumer groups convert), 512 own-vertex
  //   state loaded by the consumer threads instead of four bulk copies (measured 2.5 % slower);
  //   results are WRONG with 1 (no halo gathers), 2 / 4 (no bulk copies), 8 (no walk),
  //   16 (no slice store).  The kernel's own bits 32, 128, 256 have the opposite sense.
  const int dbg = (getenv("VF_PIPE_DBG") ? atoi(getenv("VF_PIPE_DBG")) : 0) ^ (32 | 128 | 256);
  double* mat_m = T.mat + (size_t)member * 3 * T.n_tcell;
  if (e->fan_dirty[member]) {
    fan_pack_kernel<<<T.ntiles, 128, 0, st>>>(e->dev, member, T, mat_m);
    e->fan_dirty[member] = 0;
    e->launches += 1;
    VF_CUDA(cudaGetLastError());
  }
  const NewmarkCoef nc = newmark_coef(dt);
#define VF_PIPE_GO(J_, R_, NG_, RAY_)                                                             \
  do {                                                                                            \
    VF_CUDA(cudaFuncSetAttribute(asm_fan_pipe_kernel<J_, R_, NG_, RAY_>,                          \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
    asm_fan_pipe_kernel<J_, R_, NG_, RAY_><<<grid, 128 + 128 * NG_, smem, st>>>(                  \
        e->dev, member, nc, is_static, mix, T, mat_m, (int)wj, (int)pool, pf_dist, dbg);          \
  } while (0)
#define VF_PIPE_BY_MODE(NG_, RAY_)                                                                \
  do {                                                                                            \
    if (jac && res) VF_PIPE_GO(true, true, NG_, RAY_);                                            \
    else if (jac) VF_PIPE_GO(true, false, NG_, RAY_);                                             \
    else VF_PIPE_GO(false, true, NG_, RAY_);                                                      \
  } while (0)
  // the Rayleigh-only terms of the residual are compiled out for the Kelvin-Voigt model
  const bool ray = e->desc.damping != 0;
  if (ng == 2) {
    if (ray) VF_PIPE_BY_MODE(2, true); else VF_PIPE_BY_MODE(2, false);
  } else {
    if (ray) VF_PIPE_BY_MODE(3, true); else VF_PIPE_BY_MODE(3, false);
  }
#undef VF_PIPE_BY_MODE
#undef VF_PIPE_GO
  e->launches += 1;
  VF_CUDA(cudaGetLastError());
  return 0;
}

int launch_facet_bc(vf_engine* e, int member, bool res, bool jac, double dt, int is_static,
                    const JacMix& mix, cudaStream_t st) {
  if (e->n_touch <= 0) return 0;
  static const char* env_fast = getenv("VF_FACET_FAST");
  if (e->facet_rec_dev && !(env_fast && atoi(env_fast) == 0)) {
    const int fb = 64, fg = (e->n_touch + fb - 1) / fb;
    // programmatic dependent launch: the kernel may start while the assembly kernel before it
    // in the stream is still draining (it synchronises on that grid itself, griddepcontrol.wait)
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(fg);
    cfg.blockDim = dim3(fb);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const FacetRec2D* recs = e->facet_rec_dev;
    const int n_touch = e->n_touch;
    if (jac && res)
      VF_CUDA(cudaLaunchKernelEx(&cfg, facet_bc_fast_kernel<true, true>, e->dev, member, mix, recs, n_touch));
    else if (jac)
      VF_CUDA(cudaLaunchKernelEx(&cfg, facet_bc_fast_kernel<true, false>, e->dev, member, mix, recs, n_touch));
    else
      VF_CUDA(cudaLaunchKernelEx(&cfg, facet_bc_fast_kernel<false, true>, e->dev, member, mix, recs, n_touch));
    e->launches += 1;
    VF_CUDA(cudaGetLastError());
    return 0;
  }
  const int fb = 128, fg = (e->n_touch + fb - 1) / fb;
  if (jac && res)
    facet_bc_kernel<2, true, true><<<fg, fb, 0, st>>>(e->dev, member, dt, is_static, mix, e->touch_dev, e->n_touch);
  else if (jac)
    facet_bc_kernel<2, true, false><<<fg, fb, 0, st>>>(e->dev, member, dt, is_static, mix, e->touch_dev, e->n_touch);
  else
    facet_bc_kernel<2, false, true><<<fg, fb, 0, st>>>(e->dev, member, dt, is_static, mix, e->touch_dev, e->n_touch);
  e->launches += 1;
  VF_CUDA(cudaGetLastError());
  return 0;
}

int assemble_impl(vf_engine* e, int member, int flags, double dt, int is_static, const JacMix& mix,
                  void* stream) {
  if (member < 0 || member >= e->desc.n_members) return fail("member out of range");
  const bool res = flags & 1, jac = flags & 2;
  if (!res && !jac) return 0;
  cudaStream_t st = as_stream(stream);
  const char* env_fan = getenv("VF_FAN");
  if (e->fan.ring && !(env_fan && atoi(env_fan) == 0)) {
    const char* env_pipe = getenv("VF_FAN_PIPE");
    int rc = -1;
    if (!(env_pipe && atoi(env_pipe) == 0))
      rc = launch_fan_pipe(e, member, res, jac, dt, is_static, mix, st);
    if (rc > 0) return 1;
    if (rc < 0 && launch_fan(e, member, res, jac, dt, is_static, mix, st)) return 1;
    return launch_facet_bc(e, member, res, jac, dt, is_static, mix, st);
  }
  const int grid = e->desc.ntiles;
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
    // fan-ordered meshes: one thread per scalar row completes every off-diagonal block in
    // registers and stores it straight to HBM (ROW 2, direct: 4 CTAs of 256 threads per SM);
    // meshes whose vertex fans cannot be ordered, and residual-only launches, accumulate the row
    // slice in shared memory by read-modify-write (ROW 1)
    const int nt = d.tile2_threads;
#define VF_ASM2_BY_SIZE(J_, R_, ROW_)                                                              \
  do {                                                                                            \
    if (nt <= 128) VF_LAUNCH_ASM2(J_, R_, ROW_, 128, 8);                                          \
    else if (nt <= 192) VF_LAUNCH_ASM2(J_, R_, ROW_, 192, 5);                                     \
    else if (nt <= 256) VF_LAUNCH_ASM2(J_, R_, ROW_, 256, 4);                                     \
    else VF_LAUNCH_ASM2(J_, R_, ROW_, 320, 3);                                                    \
  } while (0)
    if (jac && e->fan_ok && nt <= 320) {
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
    } else if (jac && res) VF_ASM2_BY_SIZE(true, true, 1);
    else if (jac) VF_ASM2_BY_SIZE(true, false, 1);
    else VF_ASM2_BY_SIZE(false, true, 1);
#undef VF_ASM2_BY_SIZE
#undef VF_LAUNCH_ASM2
    e->launches += 1;
    VF_CUDA(cudaGetLastError());
    return launch_facet_bc(e, member, res, jac, dt, is_static, mix, st);
  }
  // tetrahedra, and triangle meshes whose tiles do not fit the two-phase kernel: thread-per-node
  // gather.
  // VF_TET_TABLES=0: generic gathers even when the tetrahedral gather tables exist (A/B, tests)
  EngineDev dev = e->dev;
  {
    const char* env_tt = getenv("VF_TET_TABLES");
    if (env_tt ? atoi(env_tt) == 0 : !kTetTablesDefault) dev.mesh.n2e_slots = nullptr;
  }
  // Jacobian launches: rows of 32 consecutive nodes accumulated in shared memory and written by one
  // coalesced copy (VF_NODE_WARP=0: the global read-modify-write variant below)
  {
    const char* env_nw = getenv("VF_NODE_WARP");  // read per call: tests switch it
    const size_t smem =
        sizeof(double) * e->desc.dim * e->desc.dim * (size_t)e->node_warp_blocks;
    const bool want = env_nw ? atoi(env_nw) != 0 : kNodeWarpDefault;
    if (jac && want && e->node_warp_blocks > 0 && smem <= (size_t)kNodeWarpMaxSmem) {
      const int ng = (e->desc.nn + 31) / 32;
      if (e->desc.dim == 3) {
        if (res) asm_node_warp_kernel<3, true><<<ng, 32, smem, st>>>(dev, member, dt, is_static, mix);
        else asm_node_warp_kernel<3, false><<<ng, 32, smem, st>>>(dev, member, dt, is_static, mix);
      } else {
        if (res) asm_node_warp_kernel<2, true><<<ng, 32, smem, st>>>(dev, member, dt, is_static, mix);
        else asm_node_warp_kernel<2, false><<<ng, 32, smem, st>>>(dev, member, dt, is_static, mix);
      }
      e->launches += 1;
      VF_CUDA(cudaGetLastError());
      return 0;
    }
  }
  {
    const int nb = 128, ng = (e->desc.nn + nb - 1) / nb;
#define VF_LAUNCH_NODE(D)                                                                         \
  do {                                                                                            \
    if (jac && res) asm_node_global_kernel<D, true, true><<<ng, nb, 0, st>>>(dev, member, dt, is_static, mix); \
    else if (jac) asm_node_global_kernel<D, true, false><<<ng, nb, 0, st>>>(dev, member, dt, is_static, mix); \
    else asm_node_global_kernel<D, false, true><<<ng, nb, 0, st>>>(dev, member, dt, is_static, mix); \
  } while (0)
    if (e->desc.dim == 3) VF_LAUNCH_NODE(3);
    else VF_LAUNCH_NODE(2);
#undef VF_LAUNCH_NODE
    e->launches += 1;
    VF_CUDA(cudaGetLastError());
    return 0;
  }
}
}  // namespace

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


}  // extern "C"
