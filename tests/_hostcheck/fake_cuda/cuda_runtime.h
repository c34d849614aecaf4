// TEST HARNESS ONLY: stands in for <cuda_runtime.h> when a kernel header is compiled by g++
// against simt_shim.h (the vector type the kernels use, nothing else).
#pragma once
struct alignas(16) float4 {
  float x, y, z, w;
};
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct alignas(8) float2 {
  float x, y;
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
