// Populate tail for reparameterisations that are NOT a diagonal affine:
// per feature  x = h(a x' + b) * scale + shift,  float64, the inverse direction of
// /root/reference/src/nessai/reparameterisations/rescale.py:263-291 (ScaleAndShift) and :635-660
// (RescaleToBounds.inverse_reparameterise) with the rescaling functions of
// utils/rescaling.py:290-417:
//   * post_rescaling "logit"        -> h = sigmoid,  log|J| += log h + log1p(-h)      (:310-330)
//   * post_rescaling "log"          -> h = exp,      log|J| += u                       (:385-393)
//   * post_rescaling "exp"          -> h = log,      log|J| -= log u                   (:369-383)
//   * post_rescaling "gaussian_cdf" -> h = -sqrt2 erfcinv(2u), log|J| += log sqrt(2 pi) + h^2/2  (:403-407)
//   * post_rescaling "inv_gaussian_cdf" -> h = erfc(-u/sqrt2)/2, log|J| -= log sqrt(2 pi) + u^2/2 (:396-400)
//   * boundary inversion            -> h = |.|  (rescale.py:570-590: value[value < 0] *= -1; an
//                                      "upper" edge, 1 - value, is folded into a negative scale)
//   * the same functions as PRE-rescaling ("z-score-logit", "log-z-score", ...): the affine
//     map comes first, u = a x' + b with log|J| += log|a|, then h, then scale = 1, shift = 0
// followed by the affine map back to the prior bounds (rescale.py:544-553), whose log|J| is
// log|scale|, then the prior-bounds check (model.py:497-518) and log_w = log_prior - log_q
// (flowproposal/base.py:1069-1098).
//
// The per-row arithmetic is a __host__ __device__ function so that the very same source is
// compiled by g++ into a host harness in tests/ (tests/_hostcheck) and checked against the numpy
// oracle without a GPU; the library only ever calls it from the kernel below.
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define NB200_HD __host__ __device__ __forceinline__
#else
#define NB200_HD inline
#endif

namespace nb200 {

enum TailKind : int32_t {
  TAIL_IDENTITY = 0, TAIL_SIGMOID = 1, TAIL_ABS = 2, TAIL_EXP = 3, TAIL_LOG = 4,
  TAIL_NORMAL_CDF = 5, TAIL_NORMAL_QUANTILE = 6,
  // pair kinds: the output is a function of TWO flow features (u0, u1) = (x'[src0], x'[src1]),
  // the Cartesian pair of reparameterisations/angle.py:17-186 (Angle.inverse_reparameterise)
  TAIL_PAIR_FIRST = 7,
  TAIL_ANGLE = 7,       // atan2(u1, u0) * scale + shift           (scale = 1 / Angle.scale)
  TAIL_ANGLE_MOD = 8,   // (atan2(u1, u0) mod 2 pi) * scale + shift (prior starting at zero)
  TAIL_RADIUS = 9,      // sqrt(u0^2 + u1^2), log|J| -= log r
  TAIL_RADIUS_CHI = 10, // the same for an AUXILIARY radius: its chi(2) prior log r - r^2 / 2 is
                        // added to the log prior (angle.py:183-185)
  // a single-feature kind again: Dequantise (reparameterisations/discrete.py:66-78), whose
  // pre-rescaling inverse is floor(.) with no log-Jacobian
  TAIL_FLOOR = 11,
  // ToCartesian (angle.py:189-232): |atan2(u1, u0) * a| * scale + shift with a = 1 / ToCartesian.scale,
  // scale = hi - lo, shift = lo of the prior bounds; log|J| += log(hi - lo) (inverse_rescale_zero_to_one)
  TAIL_ANGLE_ABS = 12,
  // AnglePair (angle.py:235-538): THREE flow features (u0, u1, u2) = Cartesian (x, y, z); the
  // horizontal angle is TAIL_ANGLE / TAIL_ANGLE_MOD of (u0, u1)
  TAIL_ZENITH = 13,       // az-zen: atan2(sqrt(u0^2 + u1^2), u2), log|J| -= log sin(.)   (:418-452)
  TAIL_DECLINATION = 14,  // ra-dec: atan2(u2, sqrt(u0^2 + u1^2)), log|J| -= log cos(.)   (:454-489)
  TAIL_RADIUS3 = 15,      // sqrt(u0^2 + u1^2 + u2^2), log|J| -= 2 log r
  TAIL_RADIUS3_CHI = 16,  // the same for an AUXILIARY radius with its chi(3) prior (:529-537)
  // a single-feature kind: an augment parameter of AugmentedFlowProposal (proposal/augmented.py:150-178),
  // passed through unchanged, whose standard-normal prior is added to the log prior
  TAIL_GAUSS_AUX = 17,
  TAIL_N_KINDS = 18,
  // flag on a single-feature kind: floor(.) of the FINAL value, x = floor(h(a x' + b) * scale + shift) --
  // "dequantise-logit": Dequantise with a post-rescaling (discrete.py + rescale.py:635-660); no log-Jacobian
  TAIL_FLOOR_AFTER = 0x100,
  TAIL_KIND_MASK = 0xff
};
// kinds that read two or three flow features
NB200_HD bool tail_is_multi(int32_t kind) {
  kind &= TAIL_KIND_MASK;
  return kind >= TAIL_PAIR_FIRST && kind != TAIL_FLOOR && kind != TAIL_GAUSS_AUX;
}

#ifdef __CUDACC__
#define NB200_ERFCINV erfcinv
#else
// <cmath> has no erfcinv: the host harness supplies one (tests/_hostcheck/reparam_host.cpp)
extern "C" double nb200_host_erfcinv(double);
#define NB200_ERFCINV nb200_host_erfcinv
#endif

// One feature: u = a v + b, returns h(u) * scale + shift and adds the log-Jacobian of h (NOT of
// the two affine parts, which are row constants) to logj.
NB200_HD double tail_feature(int32_t kind, double a, double b, double scale, double shift, double v,
                             double& logj, double& logp_extra) {
  const bool floor_after = (kind & TAIL_FLOOR_AFTER) != 0;
  kind &= TAIL_KIND_MASK;
  const double u = a * v + b;
  double h = u;
  if (kind == TAIL_SIGMOID) {
    h = 1.0 / (1.0 + exp(-u));
    logj += log(h) + log1p(-h);
  } else if (kind == TAIL_ABS) {
    h = fabs(u);
  } else if (kind == TAIL_EXP) {
    h = exp(u);
    logj += u;
  } else if (kind == TAIL_LOG) {
    h = log(u);
    logj -= h;
  } else if (kind == TAIL_NORMAL_CDF) {
    h = 0.5 * erfc(-u / 1.4142135623730951);
    logj += -0.9189385332046727 - 0.5 * u * u;
  } else if (kind == TAIL_NORMAL_QUANTILE) {
    h = -1.4142135623730951 * NB200_ERFCINV(2.0 * u);
    logj += 0.9189385332046727 + 0.5 * h * h;
  } else if (kind == TAIL_FLOOR) {
    h = floor(u);
  } else if (kind == TAIL_GAUSS_AUX) {
    logp_extra += -0.5 * u * u - 0.9189385332046727;
  }
  const double r = h * scale + shift;
  return floor_after ? floor(r) : r;
}

// A pair / triple kind: x from (u0, u1[, u2]); the constant factor of the angle has NO log-Jacobian
// in the reference (angle.py:120-128,157-170), so scale / shift of these kinds stay out of the row
// constant -- except TAIL_ANGLE_ABS, whose map back to the prior bounds has one (tail_log_affine_sum).
NB200_HD double tail_pair(int32_t kind, double a, double scale, double shift, double u0, double u1, double u2,
                          double& logj, double& logp_extra) {
  if (kind == TAIL_ANGLE || kind == TAIL_ANGLE_MOD) {
    double th = atan2(u1, u0);
    // numpy's `% (2 pi)` of a value in [-pi, pi]: fmod, then + 2 pi when negative (a tiny
    // negative angle therefore gives exactly 2 pi, as it does in the reference)
    if (kind == TAIL_ANGLE_MOD && th < 0.0) th += 6.283185307179586;
    return th * scale + shift;
  }
  if (kind == TAIL_ANGLE_ABS) return fabs(atan2(u1, u0) * a) * scale + shift;
  if (kind == TAIL_ZENITH || kind == TAIL_DECLINATION) {
    const double rho = sqrt(u0 * u0 + u1 * u1);
    const double v = kind == TAIL_ZENITH ? atan2(rho, u2) : atan2(u2, rho);
    logj -= log(kind == TAIL_ZENITH ? sin(v) : cos(v));
    return v * scale + shift;
  }
  if (kind == TAIL_RADIUS3 || kind == TAIL_RADIUS3_CHI) {
    const double r = sqrt(u0 * u0 + u1 * u1 + u2 * u2);
    const double lr = log(r);
    logj -= 2.0 * lr;
    // scipy.stats.chi(3).logpdf(r) = log sqrt(2 / pi) + 2 log r - r^2 / 2
    if (kind == TAIL_RADIUS3_CHI) logp_extra += -0.22579135264472744 + 2.0 * lr - 0.5 * r * r;
    return r * scale + shift;
  }
  const double r = sqrt(u0 * u0 + u1 * u1);
  const double lr = log(r);
  logj -= lr;
  if (kind == TAIL_RADIUS_CHI) logp_extra += lr - 0.5 * r * r;
  return r * scale + shift;
}

// One row.  logq_flow: log q of the flow alone (NaN: the row was already dropped by the draw
// kernel).  log_affine_sum = tail_log_affine_sum.  pre_a / pre_b may be NULL (a = 1, b = 0); src
// (int32[3 D]: the one, two or three flow features output slot d reads) may be NULL (slot d
// reads feature d).  Writes x[D]; returns true when the row survives
// and then logq_out / logw_out are its log q / log weight.
NB200_HD bool tail_row(int D, const float* xp, const int32_t* kind, const int32_t* src,
                       const double* pre_a, const double* pre_b, const double* scale,
                       const double* shift, const double* lo, const double* hi,
                       double log_affine_sum, double log_prior_const, double min_log_q,
                       double logq_flow, double* x, double& logq_out, double& logw_out) {
  double logj = log_affine_sum, logp_extra = 0.0;
  bool inb = true;
  for (int d = 0; d < D; ++d) {
    const int i0 = src ? src[3 * d] : d;
    double xv;
    if (tail_is_multi(kind[d])) {
      xv = tail_pair(kind[d], pre_a ? pre_a[d] : 1.0, scale[d], shift[d], (double)xp[i0],
                     (double)xp[src ? src[3 * d + 1] : d], (double)xp[src ? src[3 * d + 2] : d], logj, logp_extra);
    } else {
      xv = tail_feature(kind[d], pre_a ? pre_a[d] : 1.0, pre_b ? pre_b[d] : 0.0, scale[d], shift[d],
                        (double)xp[i0], logj, logp_extra);
    }
    x[d] = xv;
    inb = inb && !(xv < lo[d]) && !(xv > hi[d]);
  }
  const double logq = logq_flow - logj;
  const double logw = log_prior_const + logp_extra - logq;
  // (v - v == 0) is isfinite(): it also rejects the NaN of an already-dropped row and of a
  // log / quantile evaluated outside its domain
  const bool ok = inb && (logq - logq == 0.0) && (logw - logw == 0.0) && (logq > min_log_q);
  logq_out = ok ? logq : NAN;
  logw_out = ok ? logw : NAN;
  return ok;
}

// The row constant of log|J|: the affine parts of the single-feature slots, and ToCartesian's map
// back to its prior bounds.
NB200_HD double tail_log_affine_sum(int D, const int32_t* kind, const double* pre_a, const double* scale) {
  double s = 0.0;
  for (int d = 0; d < D; ++d) {
    if (!tail_is_multi(kind[d])) s += log(fabs(scale[d])) + (pre_a ? log(fabs(pre_a[d])) : 0.0);
    else if ((kind[d] & TAIL_KIND_MASK) == TAIL_ANGLE_ABS) s += log(fabs(scale[d]));
  }
  return s;
}

#ifdef __CUDACC__
#define TAIL_THREADS 128
#define TAIL_MAXD 64

// Dynamic shared memory of the kernel: per warp one tile of 32 rows, x' (floats, row stride D + 1)
// and x (doubles, row stride D + 1) -- the odd strides keep a warp's per-row accesses conflict-free.
__host__ __device__ inline size_t tail_smem_bytes(int D) {
  return (size_t)(TAIL_THREADS / 32) * 32 * (D + 1) * (sizeof(double) + sizeof(float)) + 16;
}
#ifdef NB200_SIMT_SHIM
static unsigned char* const tail_smem_dyn = simt::dynamic_smem;
#else
extern __shared__ __align__(16) unsigned char tail_smem_dyn[];
#endif

// One thread per row; a warp moves its 32 rows between global and shared memory cooperatively, so
// that every global access is a run of consecutive addresses: 4 D + 8 bytes in, 8 D + 16 bytes out
// per row (HBM class; the float64 transcendental functions of the non-identity kinds share the
// time).  The per-feature constants sit in shared memory.
__global__ void __launch_bounds__(TAIL_THREADS)
reparam_tail_kernel(int64_t n, int D, const float* __restrict__ xp, const int32_t* __restrict__ kind,
                    const int32_t* __restrict__ src, const double* __restrict__ pre_a, const double* __restrict__ pre_b,
                    const double* __restrict__ scale, const double* __restrict__ shift,
                    const double* __restrict__ lo, const double* __restrict__ hi,
                    double log_prior_const, double min_log_q, double* __restrict__ logq,
                    double* __restrict__ logw, double* __restrict__ x64, double* __restrict__ stats) {
  __shared__ double c_s[6 * TAIL_MAXD];  // scale | shift | lo | hi | a | b
  __shared__ int32_t k_s[TAIL_MAXD];
  __shared__ int32_t src_s[3 * TAIL_MAXD];
  __shared__ double lss_s;
  for (int d = threadIdx.x; d < D; d += TAIL_THREADS) {
    c_s[d] = scale[d];
    c_s[TAIL_MAXD + d] = shift[d];
    c_s[2 * TAIL_MAXD + d] = lo[d];
    c_s[3 * TAIL_MAXD + d] = hi[d];
    c_s[4 * TAIL_MAXD + d] = pre_a ? pre_a[d] : 1.0;
    c_s[5 * TAIL_MAXD + d] = pre_b ? pre_b[d] : 0.0;
    k_s[d] = kind[d];
    for (int j = 0; j < 3; ++j) src_s[3 * d + j] = src ? src[3 * d + j] : d;
  }
  if (threadIdx.x == 0) lss_s = tail_log_affine_sum(D, kind, pre_a, scale);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ld = D + 1;
  double* xt = reinterpret_cast<double*>(tail_smem_dyn) + (size_t)warp * 32 * ld;
  float* pt = reinterpret_cast<float*>(reinterpret_cast<double*>(tail_smem_dyn) + (size_t)(TAIL_THREADS / 32) * 32 * ld) +
              (size_t)warp * 32 * ld;
  double vmax = -INFINITY, vcount = 0.0;
  const int64_t n_chunks = (n + 31) / 32;
  for (int64_t chunk = (int64_t)blockIdx.x * (TAIL_THREADS / 32) + warp; chunk < n_chunks;
       chunk += (int64_t)gridDim.x * (TAIL_THREADS / 32)) {
    const int64_t row0 = chunk * 32, row = row0 + lane;
    const int rows = (int)min((int64_t)32, n - row0);
    // x' of the chunk: consecutive lanes read consecutive floats
    const float* gsrc = xp + row0 * D;
    for (int i = lane; i < rows * D; i += 32) {
      const int r = i / D;
      pt[r * ld + (i - r * D)] = gsrc[i];
    }
    const double lq_in = row < n ? logq[row] : NAN;
    __syncwarp();
    bool ok = false;
    double lq = NAN, lw = NAN;
    if (row < n)
      ok = tail_row(D, pt + lane * ld, k_s, src_s, c_s + 4 * TAIL_MAXD, c_s + 5 * TAIL_MAXD, c_s, c_s + TAIL_MAXD,
                    c_s + 2 * TAIL_MAXD, c_s + 3 * TAIL_MAXD, lss_s, log_prior_const, min_log_q, lq_in,
                    xt + lane * ld, lq, lw);
    __syncwarp();
    // x of every row (the accept kernel only reads the rows it keeps, a device likelihood may
    // read them all): consecutive lanes write consecutive doubles
    double* gdst = x64 + row0 * D;
    for (int i = lane; i < rows * D; i += 32) {
      const int r = i / D;
      gdst[i] = xt[r * ld + (i - r * D)];
    }
    if (row < n) {
      logq[row] = lq;
      logw[row] = lw;
    }
    if (ok) {
      vmax = fmax(vmax, lw);
      vcount += 1.0;
    }
    __syncwarp();
  }
  // same publication as the draw kernels (populate_common.cuh: populate_publish)
  for (int o = 16; o > 0; o >>= 1) {
    vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    vcount += __shfl_xor_sync(0xffffffffu, vcount, o);
  }
  if ((threadIdx.x & 31) == 0 && vcount > 0) {
    atomic_max_double(stats, vmax);
    atomicAdd(stats + 1, vcount);
  }
}

// sum_i exp(log_w_i - max) over the non-NaN rows: the expected pool size of the
// accumulate_weights variant (flowproposal.py:474-475: logsumexp(log_weights - log_constant)).
// One partial per block, written unconditionally, so the total does not depend on atomics.
#define SUMEXP_THREADS 256
__global__ void __launch_bounds__(SUMEXP_THREADS)
sum_exp_kernel(const double* __restrict__ logw, int64_t n, const double* __restrict__ d_max,
               double* __restrict__ partials) {
  const double mx = *d_max;
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * SUMEXP_THREADS + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * SUMEXP_THREADS) {
    const double lw = logw[i];
    if (!isnan(lw) && lw > -INFINITY) s += exp(lw - mx);
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double ws[SUMEXP_THREADS / 32];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < SUMEXP_THREADS / 32; ++i) t += ws[i];
    partials[blockIdx.x] = t;
  }
}
#endif  // __CUDACC__

}  // namespace nb200
