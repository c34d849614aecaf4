// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, bf16) for A in TMEM vs shared memory,
// N = 64 and 16, issued back to back by one thread; and LDTM throughput.
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
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mwait(uint32_t bar, uint32_t ph) {
  uint32_t done; do { asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(done) : "r"(bar), "r"(ph) : "memory"); } while (!done);
}
// mode 0: TS N=64, 1: SS N=64, 2: TS N=16, 3: SS N=16, 4: TS N=64 independent accumulators (4 D regions)
__global__ void k(int mode, int nmma, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar; __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = 0;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tbase)), "r"(512u)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("fence.proxy.async.shared::cta;"); asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t t = tbase;
  if (threadIdx.x == 0) {
    int N = (mode >= 2 && mode < 4) ? 16 : 64;
    uint64_t adesc = mkdesc(s32(sm), 2048, 128), bdesc = mkdesc(s32(sm) + 4096, N * 16, 128);
    uint32_t id = idesc(128, N);
    long long t0 = clock64();
    for (int i = 0; i < nmma; ++i) {
      if (mode == 0 || mode == 2) mma_ts(t, t + 64 + 8 * (i & 7), bdesc, id, i > 0);
      else if (mode == 4) mma_ts(t + 128 * (i & 3), t + 64 + 128 * (i & 3), bdesc, id, i > 3);
      else mma_ss(t, adesc, bdesc, id, i > 0);
    }
    commit(s32(&bar)); mwait(s32(&bar), 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(t), "r"(512u));
}
// LDTM throughput: W warps each load `cols` columns x16 repeatedly
__global__ void kld(int iters, long long* out, float* sink) {
  __shared__ uint32_t tbase;
  if (threadIdx.x < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tbase)), "r"(512u)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t t = tbase + ((uint32_t)(((threadIdx.x >> 5) & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]) : "r"(t + 16 * (i & 7)) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    acc ^= r[0] ^ r[7] ^ r[15];
  }
  __syncthreads();
  long long t1 = clock64();
  sink[threadIdx.x] = __uint_as_float(acc);
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u));
}
int main() {
  long long* out; float* sink; cudaMalloc(&out, 64); cudaMalloc(&sink, 4096 * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  const char* names[] = {"TS N=64 (A in TMEM)", "SS N=64 (A in smem)", "TS N=16", "SS N=16", "TS N=64, 4 indep. accumulators"};
  for (int mode = 0; mode < 5; ++mode) for (int n : {64, 512}) {
    k<<<1, 128, 65536>>>(mode, n, out); cudaDeviceSynchronize();
    k<<<1, 128, 65536>>>(mode, n, out); cudaError_t e = cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
    printf("%-34s nmma=%4d cycles=%8lld  cyc/mma=%.1f %s\n", names[mode], n, c, (double)c / n, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  for (int warps : {4, 8, 16}) {
    kld<<<1, warps * 32>>>(1024, out, sink); cudaDeviceSynchronize();
    kld<<<1, warps * 32>>>(1024, out, sink); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
    double bytes = 1024.0 * warps * 32 * 16 * 4;
    printf("LDTM.x16 warps=%2d cycles=%lld  bytes/cycle/SM=%.1f  cyc per LDTM(serial)=%.1f\n", warps, c, bytes / c, (double)c / 1024);
  }
  return 0;
}
