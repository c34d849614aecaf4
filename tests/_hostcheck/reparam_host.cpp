// TEST HARNESS ONLY (never linked into libnessai_b200.so): compiles the __host__ __device__
// row function of nessai_b200/csrc/reparam_tail.cuh with g++ so that the arithmetic the CUDA
// kernel runs per row can be checked against the numpy oracle on a machine without a GPU.
// The loop below mirrors reparam_tail_kernel (constants, log-scale sum, statistics).
#include "../../nessai_b200/csrc/reparam_tail.cuh"

extern "C" void tail_rows_host(int64_t n, int D, const float* xp, const int32_t* kind,
                               const double* scale, const double* shift, const double* lo,
                               const double* hi, double log_prior_const, double min_log_q,
                               double* logq, double* logw, double* x64, double* stats) {
  double lss = 0.0;
  for (int d = 0; d < D; ++d) lss += log(fabs(scale[d]));
  for (int64_t row = 0; row < n; ++row) {
    double lq, lw;
    const bool ok = nb200::tail_row(D, xp + row * D, kind, scale, shift, lo, hi, lss, log_prior_const,
                                    min_log_q, logq[row], x64 + row * D, lq, lw);
    logq[row] = lq;
    logw[row] = lw;
    if (ok) {
      stats[0] = lw > stats[0] ? lw : stats[0];
      stats[1] += 1.0;
    }
  }
}
