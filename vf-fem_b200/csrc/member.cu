// Persistent per-member kernel (Newton / GMRES / Newmark / FSI time loop, one CTA per ensemble
// member) and the entry points built on it: vf_linear_solve, vf_solve_state1, vf_integrate,
// vf_integrate_host, vf_fluid_solve, vf_glottal_width_series.
#include "engine_internal.h"
#include "member_solver.cuh"

namespace vf {

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

using namespace vf;

namespace {

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
    if (emod_host || eta_host) std::fill(e->fan_dirty.begin(), e->fan_dirty.end(), 1);
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


}  // extern "C"
