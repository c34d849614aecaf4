// Per-row tail of one populate turn, shared by the generic and the tcgen05 kernels:
// float64 rescale + prior bounds + log-weights, exactly the numpy-side arithmetic of
// /root/reference/src/nessai/proposal/flowproposal/flowproposal.py:345-389 and
// base.py:1069-1098 for a diagonal (z-score / null) reparameterisation and a
// uniform box prior.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace nb200 {

struct PopulateArgs {
  int64_t n;
  uint64_t seed, row_offset;
  float r_max, sqrt_t;
  const double *scale, *shift, *lo, *hi;
  double log_prior_const;  // NaN: prior added by the caller
  double min_log_q;        // rows with log_q <= min_log_q are dropped (-inf: keep all)
  float* xp;      // x' (flow output, before the rescale), fp32 [n, D]
  double* logq;
  double* logw;
  float* z;
  double* stats;  // {max log_w, n_valid}
};

__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if (!(v > __longlong_as_double(assumed))) break;
    old = atomicCAS(a, assumed, __double_as_longlong(v));
  } while (assumed != old);
}

// xp(d): the row's x' (fp32) for feature d.  Updates the thread-local running
// max / count; returns nothing (outputs written to global memory).
template <int MAXD = 0, typename XP>
__device__ __forceinline__ void populate_row(const PopulateArgs& A, int D, XP xp, int64_t row,
                                             bool alive, float base_lp, float logj, double& vmax,
                                             double& vcount, const double* __restrict__ scale,
                                             const double* __restrict__ shift,
                                             const double* __restrict__ lo,
                                             const double* __restrict__ hi, double log_const) {
  if (row >= A.n) return;
  // cst: scale | shift | lo | hi, each TC_DP-strided doubles (shared or global memory)
  bool inb = true;
  if (MAXD == 16 && D == 16) {
    // full rows: 64 contiguous, 64-byte aligned bytes out; no per-feature predicates, so the
    // sixteen float64 rescale / bounds chains interleave
    float v[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) v[d] = xp(d);
    float4* o4 = reinterpret_cast<float4*>(A.xp + row * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) o4[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    int bad = 0;
#pragma unroll
    for (int d = 0; d < 16; ++d) {
      const double xv = (double)v[d] * scale[d] + shift[d];
      bad |= (xv < lo[d]) | (xv > hi[d]);
    }
    inb = !bad;
  } else if (MAXD > 0) {  // compile-time trip count: xp(d) may index registers
#pragma unroll
    for (int d = 0; d < (MAXD > 0 ? MAXD : 1); ++d) {
      if (d < D) {
        const float v = xp(d);
        A.xp[row * D + d] = v;
        const double xv = (double)v * scale[d] + shift[d];
        inb = inb && !(xv < lo[d]) && !(xv > hi[d]);
      }
    }
  } else {
    for (int d = 0; d < D; ++d) {
      const float v = xp(d);
      A.xp[row * D + d] = v;
      const double xv = (double)v * scale[d] + shift[d];
      inb = inb && !(xv < lo[d]) && !(xv > hi[d]);
    }
  }
  double logq = NAN, logw = NAN;
  bool ok = alive;
  if (ok) {
    logq = (double)base_lp - (double)logj - log_const;
    ok = isfinite(logq) && inb && (logq > A.min_log_q);
  }
  if (ok) {
    logw = (isnan(A.log_prior_const) ? 0.0 : A.log_prior_const) - logq;
    vmax = fmax(vmax, logw);
    vcount += 1.0;
  }
  A.logq[row] = ok ? logq : NAN;
  A.logw[row] = ok ? logw : NAN;
}

// row constant of log_q: D log sqrt(T) (latent temperature, base.py:401-414; a N(0, var I) base
// distribution, flows/distributions.py:17-73, is the same thing with T var in place of T) plus the
// log-Jacobian of the diagonal rescale, sum log|scale| (rescale.py:263-291)
__device__ __forceinline__ double populate_log_const(const PopulateArgs& A, int D) {
  double s = (double)D * log((double)A.sqrt_t);
  for (int d = 0; d < D; ++d) s += log(fabs(A.scale[d]));
  return s;
}

// warp-reduce the thread-local (max, count) and publish with one atomic per warp
__device__ __forceinline__ void populate_publish(const PopulateArgs& A, double vmax,
                                                 double vcount) {
  for (int o = 16; o > 0; o >>= 1) {
    vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    vcount += __shfl_xor_sync(0xffffffffu, vcount, o);
  }
  if ((threadIdx.x & 31) == 0 && vcount > 0) {
    atomic_max_double(A.stats, vmax);
    atomicAdd(A.stats + 1, vcount);
  }
}

}  // namespace nb200
