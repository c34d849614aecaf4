// Counter-based RNG: Philox4x32-10 (Salmon et al. 2011), keyed by the seed and
// counted by the GLOBAL row index, so a pool drawn on G GPUs is identical to the
// pool drawn on one.  Replaces torch.randn in
// /root/reference/src/nessai/flowmodel/base.py:889-904 and numpy's
// Generator.random in flowproposal.py:493 -- RNG streams cannot match the
// reference (different generators); parity is distributional.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nb200 {

struct Philox4 {
  uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#ifdef __CUDA_ARCH__
  const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
#else
  const uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
  const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
  const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
  const uint32_t n0 = hi1 ^ c[1] ^ k[0];
  const uint32_t n1 = lo1;
  const uint32_t n2 = hi0 ^ c[3] ^ k[1];
  const uint32_t n3 = lo0;
  c[0] = n0;
  c[1] = n1;
  c[2] = n2;
  c[3] = n3;
  k[0] += 0x9E3779B9u;
  k[1] += 0xBB67AE85u;
}

// counter = (row_lo, row_hi, block, stream); key = seed
__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint64_t seed, uint64_t row,
                                                         uint32_t block, uint32_t stream) {
  uint32_t c[4] = {(uint32_t)row, (uint32_t)(row >> 32), block, stream};
  uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
  for (int i = 0; i < 10; ++i) philox_round(c, k);
  return Philox4{c[0], c[1], c[2], c[3]};
}

// (0, 1]-open uniform from 32 random bits, fp32
__device__ __forceinline__ float u01(uint32_t r) { return ((float)r + 0.5f) * 2.3283064365386963e-10f; }

// two standard normals from two 32-bit words: Box-Muller with an accurate log
// (small radii matter) and the fast sin/cos on an angle in [-pi, pi)
__device__ __forceinline__ void box_muller(uint32_t r0, uint32_t r1, float& n0, float& n1) {
  const float u1 = fminf(u01(r0), 1.0f);
  const float u2 = u01(r1);
  const float rad = sqrtf(-2.0f * logf(u1));
  const float th = fmaf(6.283185307179586f, u2, -3.141592653589793f);
  n0 = rad * __cosf(th);
  n1 = rad * __sinf(th);
}

}  // namespace nb200
