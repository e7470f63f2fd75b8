// vffem_b200: CUDA kernels (sm_100a) and the C ABI declared in include/vffem_b200.h.
//
// This translation unit: engine lifecycle (arena layout, vf_create / vf_destroy, named-array
// access, CSR pattern).  Kernels live in
//   assembly.cu   residual + Jacobian assembly (asm_fan_kernel, facet_bc_kernel, ...)
//   krylov.cu     CSR SpMV, block-Jacobi, fused dots / updates (grid-wide Krylov pieces)
//   member.cu     persistent Newton/GMRES/Newmark/FSI time loop, one CTA per member
// See DESIGN.md for the layout and the roofline of each.

#include "engine_internal.h"

#include <mutex>

namespace vf {

static thread_local std::string g_err;

int fail(const std::string& msg) {
  g_err = msg;
  return 1;
}

}  // namespace vf

using namespace vf;

// layout pinned for the ctypes binding (femvf_b200/_cabi.py, tests/test_cabi.py)
static_assert(sizeof(vf_solver_opts) == 56, "vf_solver_opts layout changed");
static_assert(VF_ARRAY_COUNT == 26, "vf_array_id changed: update _cabi.ARRAY_IDS");


namespace {

// engines that raised the release threshold of a device's default memory pool (vf_create), and
// the value to put back when the last of them is destroyed
std::mutex g_pool_mutex;
int g_pool_users = 0, g_pool_dev = -1;
unsigned long long g_pool_prev = 0;

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

  bool pool_user = false;
  {
    // vf_integrate_host stages through cudaMallocAsync: let the device's default pool keep up to
    // 1 GiB of freed blocks instead of returning them to the driver at every synchronisation
    // point.  The previous threshold is restored by vf_destroy.
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long prev = 0, keep = 1ull << 30;
      std::lock_guard<std::mutex> lock(g_pool_mutex);
      if (g_pool_users > 0 && g_pool_dev == dev) {
        ++g_pool_users;
        pool_user = true;
      } else if (g_pool_users == 0 &&
                 cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &prev) == cudaSuccess &&
                 prev < keep &&
                 cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) == cudaSuccess) {
        g_pool_prev = prev;
        g_pool_dev = dev;
        g_pool_users = 1;
        pool_user = true;
      }
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
  e->fan = FanTablesDev{};
  e->fan_mem = nullptr;
  e->fan_max_wblocks = 0;
  e->pool_user = pool_user;
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
  // boundary-node records (triangles without contact / membrane terms, at most two pressure
  // facets per node): state-independent facet data gathered once on the host
  e->facet_rec_dev = nullptr;
  if (d.dim == 2 && !d.contact && !d.membrane && !touch.empty()) {
    std::vector<FacetRec2D> recs(touch.size());
    bool ok = true;
    for (size_t t = 0; t < touch.size() && ok; ++t) {
      const int i = touch[t];
      FacetRec2D& r = recs[t];
      memset(&r, 0, sizeof(r));
      r.node = i;
      r.b0 = d.brptr_host[i];
      r.deg = d.brptr_host[i + 1] - r.b0;
      const int* bcol_i = d.bcol_host + r.b0;
      r.self = find_slot(bcol_i, r.deg, i);
      r.bc = (d.bc_host[2 * i] ? 1 : 0) | (d.bc_host[2 * i + 1] ? 2 : 0);
      r.nfac = d.n2f_ptr_host[i + 1] - d.n2f_ptr_host[i];
      if (r.nfac > 2) {
        ok = false;
        break;
      }
      for (int k = 0; k < r.nfac; ++k) {
        const int ref = d.n2f_host[d.n2f_ptr_host[i] + k];
        const int f = ref >> 2;
        FacetRec2D::Facet& q = r.f[k];
        q.a = ref & 3;
        q.o = d.pf_opp_host[f];
        const int cell = d.pf_cell_host[f];
        double x[3][2];
        for (int b = 0; b < 3; ++b) {
          q.nd[b] = d.cells_host[(size_t)b * d.ne + cell];
          x[b][0] = d.xyz_host[q.nd[b]];
          x[b][1] = d.xyz_host[(size_t)d.nn + q.nd[b]];
          q.slot[b] = find_slot(bcol_i, r.deg, q.nd[b]);
        }
        CellGeo<2> g;
        p1_geometry(x, g);
        facet_geometry<2>(g, q.o, q.N, q.meas);
        for (int b = 0; b < 3; ++b) {
          q.G[b][0] = g.G[b][0];
          q.G[b][1] = g.G[b][1];
        }
      }
    }
    if (ok) {
      VF_CUDA(cudaMalloc(&e->facet_rec_dev, sizeof(FacetRec2D) * recs.size()));
      VF_CUDA(cudaMemcpyAsync(e->facet_rec_dev, recs.data(), sizeof(FacetRec2D) * recs.size(),
                              cudaMemcpyHostToDevice, st));
      VF_CUDA(cudaStreamSynchronize(st));
    }
  }
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

  if (assembly_configure(e, d, two_phase)) {
    if (e->node_mem) cudaFree(e->node_mem);
    delete e;
    return 1;
  }
  *out = e;
  return 0;
}

void vf_destroy(vf_engine* e) {
  if (!e) return;
  if (e->fan_mem) cudaFree(e->fan_mem);
  if (e->node_mem) cudaFree(e->node_mem);
  if (e->facet_rec_dev) cudaFree(e->facet_rec_dev);
  ilu_release(e);
  band_release(e);
  if (e->pool_user) {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (--g_pool_users == 0) {
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, g_pool_dev) == cudaSuccess)
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &g_pool_prev);
      (void)cudaGetLastError();
    }
  }
  delete e;
}

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
  if ((array_id == VF_RHO || array_id == VF_ETA || array_id == VF_EMOD) && !e->fan_dirty.empty())
    e->fan_dirty[member] = 1;
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

int64_t vf_launch_count(const vf_engine* e) { return e ? e->launches : 0; }

}  // extern "C"
