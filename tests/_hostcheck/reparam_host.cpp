// TEST HARNESS ONLY (never linked into libnessai_b200.so): compiles the __host__ __device__
// row function of nessai_b200/csrc/reparam_tail.cuh with g++ so that the arithmetic the CUDA
// kernel runs per row can be checked against the numpy oracle on a machine without a GPU.
// The loop below mirrors reparam_tail_kernel (constants, log-scale sum, statistics).
#include "../../nessai_b200/csrc/reparam_tail.cuh"

// erfcinv for the host build (<cmath> has none; the device build uses CUDA's): Newton steps on
// erfc from a crude start, to double precision for y in (0, 2).
extern "C" double nb200_host_erfcinv(double y) {
  if (!(y > 0.0) || !(y < 2.0)) return y == 0.0 ? INFINITY : (y == 2.0 ? -INFINITY : NAN);
  const bool flip = y > 1.0;
  const double yy = flip ? 2.0 - y : y;  // (0, 1]: x >= 0
  double x = yy < 1e-300 ? 26.0 : sqrt(fmax(-log(yy * (1.0 + 0.5 * sqrt(-log(fmin(yy, 0.9))))), 0.0));
  for (int it = 0; it < 60; ++it) {
    const double f = erfc(x) - yy;
    const double df = -1.1283791670955126 * exp(-x * x);  // d erfc / dx = -2/sqrt(pi) e^{-x^2}
    double step = f / df;
    step /= 1.0 + x * step;  // Halley: f'' / f' = -2x
    x -= step;
    if (fabs(step) <= 4e-16 * fmax(1.0, fabs(x))) break;
  }
  return flip ? -x : x;
}

extern "C" void tail_rows_host(int64_t n, int D, const float* xp, const int32_t* kind,
                               const int32_t* src, const double* pre_a, const double* pre_b,
                               const double* scale, const double* shift, const double* lo,
                               const double* hi, double log_prior_const, double min_log_q,
                               double* logq, double* logw, double* x64, double* stats) {
  const double lss = nb200::tail_log_affine_sum(D, kind, pre_a, scale);
  for (int64_t row = 0; row < n; ++row) {
    double lq, lw;
    const bool ok = nb200::tail_row(D, xp + row * D, kind, src, pre_a, pre_b, scale, shift, lo, hi, lss,
                                    log_prior_const, min_log_q, logq[row], x64 + row * D, lq, lw);
    logq[row] = lq;
    logw[row] = lw;
    if (ok) {
      stats[0] = lw > stats[0] ? lw : stats[0];
      stats[1] += 1.0;
    }
  }
}
