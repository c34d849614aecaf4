// TEST HARNESS ONLY: runs reparam_tail_kernel and sum_exp_kernel -- the CUDA source, unchanged --
// under the CPU SIMT shim (simt_shim.h), so that what the row-function check cannot see (staging
// of the constants in shared memory, the grid-stride loop, the warp / block reductions and the
// statistics) is executed and compared with the numpy oracle on a machine without a GPU.
// (its own namespace: reparam_host.cpp, linked into the same library for nb200_host_erfcinv,
// compiles the header's host flavour under the name nb200)
#define NB200_SIMT_SHIM 1
#define nb200 nb200_simt
#include "simt_shim.h"

#include "../../nessai_b200/csrc/reparam_tail.cuh"

extern "C" void simt_reparam_tail(int grid, int64_t n, int D, const float* xp, const int32_t* kind,
                                  const int32_t* src, const double* pre_a, const double* pre_b,
                                  const double* scale, const double* shift, const double* lo,
                                  const double* hi, double log_prior_const, double min_log_q,
                                  double* logq, double* logw, double* x64, double* stats) {
  simt_launch(nb200::reparam_tail_kernel, (unsigned)grid, TAIL_THREADS, n, D, xp, kind, src, pre_a, pre_b, scale,
              shift, lo, hi, log_prior_const, min_log_q, logq, logw, x64, stats);
}

extern "C" void simt_sum_exp(int grid, const double* logw, int64_t n, const double* d_max, double* partials) {
  simt_launch(nb200::sum_exp_kernel, (unsigned)grid, SUMEXP_THREADS, logw, n, d_max, partials);
}
