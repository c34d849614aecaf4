// TEST HARNESS ONLY: coupling_vec_kernel / coupling_row_kernel (nessai_b200/csrc/coupling.cuh, the
// HBM-roofline kernel of the bench) -- the CUDA source, unchanged -- under the CPU SIMT shim.
// Compile with -I tests/_hostcheck/fake_cuda so that the header's <cuda_runtime.h> is the stand-in.
#define nb200 nb200_simt_coupling
#include "simt_shim.h"

#include "../../nessai_b200/csrc/coupling.cuh"

extern "C" int simt_coupling(int grid, const float* x, const float* params, float* y, float* logdet, int64_t n,
                             int D, const int32_t* transform_features, int d_tr, int additive, int inverse,
                             int vectorised) {
  nb200::CouplingMap map;
  for (int f = 0; f < nb200::CP_MAXD; ++f) map.rank[f] = -1;
  for (int i = 0; i < d_tr; ++i) map.rank[transform_features[i]] = (int8_t)i;
  if (!vectorised) {
    simt_launch(nb200::coupling_row_kernel, (unsigned)grid, 256u, x, params, y, logdet, n, D, d_tr, additive, inverse,
                map);
    return 0;
  }
  const int lpr = D / 4;
  if (D % 4 || (lpr & (lpr - 1)) || lpr > 16) return 1;
#define LAUNCH(L) \
  simt_launch(nb200::coupling_vec_kernel<L>, (unsigned)grid, 256u, x, params, y, logdet, n, d_tr, additive, inverse, map)
  if (lpr == 1) LAUNCH(1);
  else if (lpr == 2) LAUNCH(2);
  else if (lpr == 4) LAUNCH(4);
  else if (lpr == 8) LAUNCH(8);
  else LAUNCH(16);
#undef LAUNCH
  return 0;
}

extern "C" double nb200_host_erfcinv(double) { return 0.0; }  // the shim declares it; unused here
