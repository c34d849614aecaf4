// TEST HARNESS ONLY: compiles nessai_b200/csrc/philox.cuh -- the generator the draw kernels use --
// for the host, so that the CUDA source itself (not a restatement) is checked against the
// Random123 known-answer vectors and against oracle/philox_numpy.py without a GPU.
#include <cmath>
// the fast-math intrinsics of the device build (sin / cos of an angle in [-pi, pi)): glibc's
// <math.h> already declares these two names, so map them to the accurate functions by macro
#define __cosf(x) cosf(x)
#define __sinf(x) sinf(x)
#include "../../nessai_b200/csrc/philox.cuh"

extern "C" void philox_host(uint64_t seed, uint64_t row, uint32_t block, uint32_t stream, uint32_t* out) {
  const nb200::Philox4 r = nb200::philox4x32_10(seed, row, block, stream);
  out[0] = r.x, out[1] = r.y, out[2] = r.z, out[3] = r.w;
}

// the latent draw of one row as the kernels form it (sample_latent_kernel / populate_draw_kernel)
extern "C" void latent_row_host(uint64_t seed, uint64_t row, int D, float* z) {
  for (int d0 = 0; d0 < D; d0 += 4) {
    const nb200::Philox4 r = nb200::philox4x32_10(seed, row, d0 / 4, 0);
    float v[4];
    nb200::box_muller(r.x, r.y, v[0], v[1]);
    nb200::box_muller(r.z, r.w, v[2], v[3]);
    for (int j = 0; j < 4 && d0 + j < D; ++j) z[d0 + j] = v[j];
  }
}

// the uniform of the rejection step (accept_row in nessai_b200.cu: stream 1, block 0, first word)
extern "C" double accept_uniform_host(uint64_t seed, uint64_t row) {
  const nb200::Philox4 r = nb200::philox4x32_10(seed, row, 0, 1);
  return ((double)r.x + 0.5) * 2.3283064365386963e-10;
}
