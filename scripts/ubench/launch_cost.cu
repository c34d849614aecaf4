// Micro-benchmark: GPU time per launch of tiny kernels as a function of parameter size,
// dynamic shared memory and CTA shape (the training step is a chain of such launches).
#include <cstdio>
#include <cuda_runtime.h>
struct Big { int v[2304]; };   // 9 KB, like TrPlan
struct Small { int v[16]; };
extern __shared__ float dyn[];
template <class P> __global__ void k(const __grid_constant__ P p, float* out, int work) {
  float s = 0.f;
  for (int i = 0; i < work; ++i) s += __ldg(out + ((threadIdx.x + i * 37) & 1023));
  if (s == 12345.f) out[0] = s + p.v[0] + dyn[0];
}
template <class P> float run(const char* name, int grid, int block, size_t smem, int work, int n = 200) {
  P p{}; float* out; cudaMalloc(&out, 4096 * 4); cudaMemset(out, 0, 4096 * 4);
  cudaFuncSetAttribute(k<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 10; ++i) k<P><<<grid, block, smem>>>(p, out, work);
  cudaEventRecord(a);
  for (int i = 0; i < n; ++i) k<P><<<grid, block, smem>>>(p, out, work);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  printf("%-44s grid=%3d block=%3d smem=%6zu work=%3d : %.2f us/launch %s\n", name, grid, block, smem, work, 1e3 * ms / n, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); return ms;
}
int main() {
  run<Small>("small params", 63, 128, 0, 0);
  run<Small>("small params", 63, 512, 0, 0);
  run<Big>("9 KB params", 63, 512, 0, 0);
  run<Small>("small params, 75 KB smem", 63, 512, 75 * 1024, 0);
  run<Big>("9 KB params, 75 KB smem", 63, 512, 75 * 1024, 0);
  run<Big>("9 KB params, 95 KB smem", 63, 512, 95 * 1024, 0);
  run<Small>("small, 16 dependent L2 loads", 63, 512, 0, 16);
  run<Small>("small, 64 dependent L2 loads", 63, 512, 0, 64);
  // alternating carve-outs
  {
    Small p{}; float* out; cudaMalloc(&out, 4096 * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int i = 0; i < 100; ++i) { k<Small><<<63, 512, 75 * 1024>>>(p, out, 0); k<Small><<<63, 512, 0>>>(p, out, 0); }
    cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b);
    printf("alternating 75 KB / 0 KB smem: %.2f us/launch\n", 1e3 * ms / 200);
    cudaEventRecord(a);
    for (int i = 0; i < 100; ++i) { k<Small><<<63, 512, 75 * 1024>>>(p, out, 0); k<Small><<<63, 512, 95 * 1024>>>(p, out, 0); }
    cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    printf("alternating 75 KB / 95 KB smem: %.2f us/launch\n", 1e3 * ms / 200);
  }
  return 0;
}
