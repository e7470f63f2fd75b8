// 1D quasi-steady Bernoulli glottal flow, one warp per fluid channel.
//
// Stands in for the jax.jit-compiled closures of
// /root/reference/src/femvf/residuals/fluid.py:17-34 (Bernoulli q, p),
// :229-311 (_BernoulliAreaRatioSep), :64-128 (_BernoulliFixedSep),
// :137-220 (_BernoulliSmoothMinSep) and equations/smoothapproximation.py:10-30,
// evaluated by JaxModel.solve_state1 (models/transient.py:667-672):
//   state1 - res(state1, control, prop) = (q, p)(area, psub, psup).
// The min / first-argmin searches over the surface line are warp shuffle reductions.
#pragma once

#include <math_constants.h>

namespace vf {

enum FluidKind { FLUID_AREA_RATIO_SEP = 0, FLUID_FIXED_SEP = 1, FLUID_SMOOTH_MIN_SEP = 2 };

// per-channel fluid property block
enum FluidProp { FP_RHO_AIR = 0, FP_R_SEP = 1, FP_AREA_LB = 2, FP_ZETA_MIN = 3, FP_ZETA_SEP = 4, FP_COUNT = 5 };

// (value, index) minimum with ties resolved to the lower index: first occurrence
__device__ __forceinline__ void warp_argmin(double& v, int& idx) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, v, off);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, off);
    if (ov < v || (ov == v && oi < idx)) {
      v = ov;
      idx = oi;
    }
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}

__device__ __forceinline__ double bernoulli_q(double psub, double psep, double area_sep,
                                              double rho) {
  // fluid.py:17-27 with area_sub = inf
  const double dp = psub - psep;
  const double sgn = (dp > 0.0) ? 1.0 : ((dp < 0.0) ? -1.0 : 0.0);
  const double inv2 = 1.0 / (area_sep * area_sep);
  return sgn * sqrt(2.0 / rho * fabs(dp) / inv2);
}

__device__ __forceinline__ double bernoulli_p(double q, double psep, double area_sep,
                                              double area, double rho) {
  // fluid.py:30-34
  return psep + 0.5 * rho * q * q * (1.0 / (area_sep * area_sep) - 1.0 / (area * area));
}

// One warp evaluates one channel.  area_in / p_out have ns entries, q_out one entry.
// All 32 lanes of the warp must call this.
__device__ inline void bernoulli_channel(int kind, int idx_sep_fixed, int ns, const double* s,
                                         const double* area_in, double psub, double psup,
                                         const double* fprop, double* q_out, double* p_out) {
  const int lane = threadIdx.x & 31;
  const double rho = fprop[FP_RHO_AIR];
  if (kind == FLUID_AREA_RATIO_SEP) {
    const double r_sep = fprop[FP_R_SEP], area_lb = fprop[FP_AREA_LB];
    // amin, first index of the minimum
    double v = CUDART_INF;
    int idx = 0x7fffffff;
    for (int k = lane; k < ns; k += 32) {
      const double a = fmax(area_in[k], area_lb);
      if (a < v) {
        v = a;
        idx = k;
      }
    }
    warp_argmin(v, idx);
    const double amin = v;
    const double smin = s[idx];
    const double asep = r_sep * amin;
    // separation point: first argmin of |area - asep| over s >= smin (nanargmin)
    double dv = CUDART_INF;
    int didx = 0x7fffffff;
    for (int k = lane; k < ns; k += 32) {
      if (s[k] >= smin) {
        const double a = fmax(area_in[k], area_lb);
        const double dd = fabs(a - asep);
        if (dd < dv) {
          dv = dd;
          didx = k;
        }
      }
    }
    warp_argmin(dv, didx);
    const double ssep = s[didx];
    const double q = bernoulli_q(psub, psup, asep, rho);
    for (int k = lane; k < ns; k += 32) {
      const double a = fmax(area_in[k], area_lb);
      const double p = bernoulli_p(q, psup, asep, a, rho);
      const double f = (s[k] < ssep) ? 1.0 : 0.0;
      p_out[k] = f * p + (1.0 - f) * psup;
    }
    if (lane == 0) *q_out = q;
  } else if (kind == FLUID_FIXED_SEP) {
    const double asep = area_in[idx_sep_fixed];
    const double q = bernoulli_q(psub, psup, asep, rho);
    for (int k = lane; k < ns; k += 32) {
      const double p = bernoulli_p(q, psup, asep, area_in[k], rho);
      const double f = (k <= idx_sep_fixed) ? 1.0 : 0.0;
      p_out[k] = f * p + (1.0 - f) * psup;
    }
    if (lane == 0) *q_out = q;
  } else {
    // smooth-min separation; reshape_args sets zeta_sep := zeta_min (fluid.py:153-154)
    const double zeta_min = fprop[FP_ZETA_MIN];
    const double zeta_sep = zeta_min;
    double mx = -CUDART_INF;
    for (int k = lane; k < ns; k += 32) mx = fmax(mx, -area_in[k] / zeta_min);
    mx = warp_max(mx);
    double se = 0.0;
    for (int k = lane; k < ns; k += 32) se += exp(-area_in[k] / zeta_min - mx);
    se = warp_sum(se);
    // trapezoid integrals of w, a*w, s*w over s
    double iw = 0.0, iaw = 0.0, isw = 0.0;
    for (int k = lane; k < ns - 1; k += 32) {
      const double w0 = exp(-area_in[k] / zeta_min - mx) / se;
      const double w1 = exp(-area_in[k + 1] / zeta_min - mx) / se;
      const double ds = s[k + 1] - s[k];
      iw += 0.5 * (w0 + w1) * ds;
      iaw += 0.5 * (area_in[k] * w0 + area_in[k + 1] * w1) * ds;
      isw += 0.5 * (s[k] * w0 + s[k + 1] * w1) * ds;
    }
    iw = warp_sum(iw);
    iaw = warp_sum(iaw);
    isw = warp_sum(isw);
    const double asep = iaw / iw, ssep = isw / iw;
    const double q = bernoulli_q(psub, psup, asep, rho);
    for (int k = lane; k < ns; k += 32) {
      const double p = bernoulli_p(q, psup, asep, area_in[k], rho);
      const double xarg = -(s[k] - ssep) / zeta_sep;
      const double f = (xarg >= 0.0) ? 1.0 / (1.0 + exp(-xarg)) : exp(xarg) / (1.0 + exp(xarg));
      p_out[k] = f * p;
    }
    if (lane == 0) *q_out = q;
  }
}

}  // namespace vf
