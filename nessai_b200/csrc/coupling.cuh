// The element-wise stage of a coupling layer on its own: nflows'
// AffineCouplingTransform._coupling_transform_forward / _inverse with the
// conditioner output SUPPLIED -- y = x * scale + shift on the transformed features,
// scale = sigmoid(u + 2) + 1e-3, log|det| = sum log scale; identity features pass through.
// This is the HBM-class part of the flow (4D + 4*2*d_tr read, 4D + 4 written per row: 196 B at
// D = 16) and the kernel the "coupling forward vs HBM roofline" figure is measured on.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace nb200 {

constexpr int CP_MAXD = 64;
struct CouplingMap {
  int8_t rank[CP_MAXD];  // rank of feature f among the transformed features, -1: identity
};

__device__ __forceinline__ void cp_transform(float& v, float shift, float u, bool additive, bool inverse,
                                             float& ld) {
  float s = 1.f, ls = 0.f;
  if (!additive) {
    s = __fdividef(1.f, 1.f + __expf(-(u + 2.f))) + 1e-3f;
    ls = __logf(s);
  }
  v = inverse ? __fdividef(v - shift, s) : fmaf(v, s, shift);
  ld += inverse ? -ls : ls;
}

// LPR = lanes per row (D = 4 * LPR): each lane owns 4 consecutive features of a row, so a warp
// reads / writes 32 consecutive float4 (512 contiguous bytes) per instruction.
template <int LPR>
__global__ void __launch_bounds__(256) coupling_vec_kernel(const float* __restrict__ x,
                                                           const float* __restrict__ params,
                                                           float* __restrict__ y,
                                                           float* __restrict__ logdet, int64_t n, int d_tr,
                                                           int additive, int inverse, CouplingMap map) {
  constexpr int D = 4 * LPR;
  const int np = additive ? d_tr : 2 * d_tr;
  const int64_t total = n * LPR;
  const int j = threadIdx.x % LPR;
  int r0 = map.rank[4 * j], r1 = map.rank[4 * j + 1], r2 = map.rank[4 * j + 2], r3 = map.rank[4 * j + 3];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ((total + 31) & ~31ll);
       i += (int64_t)gridDim.x * blockDim.x) {
    const bool ok = i < total;
    const int64_t row = i / LPR;
    float4 v = ok ? __ldcs(reinterpret_cast<const float4*>(x) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float* p = params + row * np;
    float ld = 0.f;
    if (ok) {
      if (r0 >= 0) cp_transform(v.x, __ldg(p + r0), additive ? 0.f : __ldg(p + d_tr + r0), additive, inverse, ld);
      if (r1 >= 0) cp_transform(v.y, __ldg(p + r1), additive ? 0.f : __ldg(p + d_tr + r1), additive, inverse, ld);
      if (r2 >= 0) cp_transform(v.z, __ldg(p + r2), additive ? 0.f : __ldg(p + d_tr + r2), additive, inverse, ld);
      if (r3 >= 0) cp_transform(v.w, __ldg(p + r3), additive ? 0.f : __ldg(p + d_tr + r3), additive, inverse, ld);
      __stcs(reinterpret_cast<float4*>(y) + i, v);
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) ld += __shfl_xor_sync(0xffffffffu, ld, o);
    if (ok && j == 0) logdet[row] = ld;
  }
  (void)D;
}

// any D <= CP_MAXD: one thread per row
__global__ void __launch_bounds__(256) coupling_row_kernel(const float* __restrict__ x,
                                                           const float* __restrict__ params,
                                                           float* __restrict__ y,
                                                           float* __restrict__ logdet, int64_t n, int D,
                                                           int d_tr, int additive, int inverse,
                                                           CouplingMap map) {
  const int np = additive ? d_tr : 2 * d_tr;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n;
       row += (int64_t)gridDim.x * blockDim.x) {
    const float* p = params + row * np;
    float ld = 0.f;
    for (int f = 0; f < D; ++f) {
      float v = x[row * D + f];
      const int r = map.rank[f];
      if (r >= 0) cp_transform(v, p[r], additive ? 0.f : p[d_tr + r], additive, inverse, ld);
      y[row * D + f] = v;
    }
    logdet[row] = ld;
  }
}

}  // namespace nb200
