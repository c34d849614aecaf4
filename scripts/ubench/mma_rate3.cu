// Micro-benchmark 3: per-thread tcgen05.mma issue cost, divergent single-lane issue (ptxas wraps
// every MMA in an ELECT / R2UR.BROADCAST waterfall loop) vs warp-convergent issue with the
// instruction predicated by elect.sync and warp-uniform operands.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mkdesc(uint32_t a, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;}" ::"r"(d), "r"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
// whole warp executes; one elected lane issues
__device__ __forceinline__ void mma_ts_elect(uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{.reg .pred p, pe; setp.ne.b32 p, %4, 0; elect.sync _|pe, 0xffffffff; @pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;}" ::"r"(d), "r"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void commit_elect(uint32_t bar) { asm volatile("{.reg .pred pe; elect.sync _|pe, 0xffffffff; @pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];}" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mwait(uint32_t bar, uint32_t ph) {
  uint32_t done; do { asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(done) : "r"(bar), "r"(ph) : "memory"); } while (!done);
}
template <int MODE> __global__ void k(int nmma, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar; __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = 0;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tbase)), "r"(512u)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("fence.proxy.async.shared::cta;"); asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t id = idesc(128, 64);
  if (MODE == 0) {
    if (threadIdx.x == 0) {
      const uint32_t t = tbase;
      const uint64_t bdesc = mkdesc(s32(sm) + 4096, 64 * 16, 128);
      long long t0 = clock64();
#pragma unroll 8
      for (int i = 0; i < nmma; ++i) mma_ts(t, t + 64 + 8 * (i & 3), bdesc + (uint64_t)((i & 3) * 128), id, i > 0);
      commit(s32(&bar)); mwait(s32(&bar), 0);
      out[0] = clock64() - t0;
    }
  } else if (threadIdx.x < 32) {
    const uint32_t t = __shfl_sync(0xffffffffu, tbase, 0);
    const uint64_t bdesc = mkdesc(s32(sm) + 4096, 64 * 16, 128);
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < nmma; ++i) mma_ts_elect(t, t + 64 + 8 * (i & 3), bdesc + (uint64_t)((i & 3) * 128), id, i > 0);
    commit_elect(s32(&bar)); mwait(s32(&bar), 0);
    if (threadIdx.x == 0) out[0] = clock64() - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u));
}
template <int MODE> void run(const char* name) {
  long long* out; cudaMalloc(&out, 64);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int n : {64, 512}) {
    k<MODE><<<1, 128, 65536>>>(n, out); cudaDeviceSynchronize();
    k<MODE><<<1, 128, 65536>>>(n, out); cudaError_t e = cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
    printf("%-40s nmma=%4d cycles=%8lld  cyc/mma=%.1f %s\n", name, n, c, (double)c / n, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
}
int main() {
  run<0>("divergent single lane (waterfall)");
  run<1>("convergent warp + elect.sync");
  return 0;
}
