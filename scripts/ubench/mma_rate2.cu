// Micro-benchmark 2: cost of one tcgen05.mma (M=128, K=16, bf16, A in TMEM) as a function of
// N and of the number of concurrently issuing warps; round-trip latency of MMA+commit+wait.
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
// nwarps issuing warps; warp w: D at column w*dstride, A at w*dstride + N (8 cols per k step, 4 steps cycled)
// per_commit: MMAs per commit+wait (0 = one commit at the end)
__global__ void k(int N, int nwarps, int nmma, int per_commit, int ss, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar[4]; __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 65536; i += blockDim.x) sm[i] = 0;
  if (threadIdx.x == 0) { for (int w = 0; w < 4; ++w) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[w]))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tbase)), "r"(512u)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("fence.proxy.async.shared::cta;"); asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t t = tbase;
  const int w = threadIdx.x >> 5;
  const int dstride = 512 / nwarps;
  long long t0 = 0, t1 = 0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && w < nwarps) {
    const uint32_t d = t + w * dstride, a = d + N;
    const uint64_t bdesc = mkdesc(s32(sm) + 8192 * w, N * 16, 128);
    const uint64_t adesc = mkdesc(s32(sm) + 32768 + 4096 * w, 2048, 128);
    const uint32_t id = idesc(128, N);
    uint32_t ph = 0;
    t0 = clock64();
    if (per_commit == 0) {
      for (int i = 0; i < nmma; ++i) {
        if (ss) mma_ss(d, adesc, bdesc, id, i > 0); else mma_ts(d, a + 8 * (i & 3), bdesc, id, i > 0);
      }
      commit(s32(&bar[w])); mwait(s32(&bar[w]), 0);
    } else {
      for (int i = 0; i < nmma; i += per_commit) {
        for (int j = 0; j < per_commit; ++j) {
          if (ss) mma_ss(d, adesc, bdesc, id, j > 0); else mma_ts(d, a + 8 * (j & 3), bdesc, id, j > 0);
        }
        commit(s32(&bar[w])); mwait(s32(&bar[w]), ph); ph ^= 1;
      }
    }
    t1 = clock64();
    out[w] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(t), "r"(512u));
}
int main() {
  long long* out; cudaMalloc(&out, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  const int nmma = 512;
  for (int ss : {0, 1}) for (int nw : {1, 2, 4}) for (int N : {16, 32, 64, 96, 128, 192, 256}) {
    if (N + 32 > 512 / nw) continue;
    for (int pc : {0, 1, 3, 8, 16}) {
      cudaMemset(out, 0, 64);
      k<<<1, 128, 65536>>>(N, nw, nmma, pc, ss, out); cudaDeviceSynchronize();
      k<<<1, 128, 65536>>>(N, nw, nmma, pc, ss, out); cudaError_t e = cudaDeviceSynchronize();
      long long c[4]; cudaMemcpy(c, out, 32, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int w = 0; w < nw; ++w) mx = c[w] > mx ? c[w] : mx;
      printf("%s N=%3d issuers=%d per_commit=%2d  cyc/mma(per issuer)=%7.1f  SM cyc per mma=%7.1f  %s\n", ss ? "SS" : "TS", N, nw, pc,
             (double)mx / nmma, (double)mx / (nmma * nw), e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  return 0;
}
