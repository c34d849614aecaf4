// Micro-benchmark: issue throughput (cycles per warp-instruction per SMSP) of the ops the
// tcgen05 epilogue is made of.  One CTA of W warps per SM sub-partition x 4.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N_IT 256
#define UNR 16
template <int OP> __device__ __forceinline__ void body(float (&f)[UNR], uint32_t (&u)[UNR]) {
#pragma unroll
  for (int i = 0; i < UNR; ++i) {
    if (OP == 0) f[i] = fmaf(f[i], 1.0001f, 0.5f);                       // FFMA
    if (OP == 1) asm volatile("cvt.rz.relu.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(f[i]), "f"(f[(i + 1) % UNR]));  // F2FP
    if (OP == 2) asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(u[i]) : "r"(u[i]), "r"(u[(i + 1) % UNR]));        // PRMT
    if (OP == 3) f[i] = fmaxf(f[i], 0.25f);                               // FMNMX
    if (OP == 4) u[i] = (u[i] & 0xffff0000u) ^ u[(i + 1) % UNR];           // LOP3
    if (OP == 5) u[i] = u[i] << 16 | u[i] >> 3;                           // SHF
    if (OP == 6) f[i] = f[i] + 1.5f;                                       // FADD
    if (OP == 7 && (i & 1) == 0) asm volatile("{.reg .b64 a,b,c; mov.b64 a,{%0,%1}; mov.b64 b,{%2,%3}; fma.rn.f32x2 c,a,b,a; mov.b64 {%0,%1},c;}" : "+f"(f[i]), "+f"(f[i + 1]) : "f"(1.0001f), "f"(0.9999f));  // FFMA2 (UNR/2 instrs)
    if (OP == 8) f[i] = __expf(f[i]);                                      // MUFU.EX2 + FMUL
    if (OP == 9) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(f[i]), "f"(f[(i + 1) % UNR]));
  }
}
template <int OP> __global__ void k(float* out, long long* cyc) {
  float f[UNR]; uint32_t u[UNR];
  for (int i = 0; i < UNR; ++i) { f[i] = threadIdx.x * 0.001f + i; u[i] = threadIdx.x * 77 + i; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < N_IT; ++it) body<OP>(f, u);
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < UNR; ++i) s += f[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char* name, int per_it) {
  float* out; long long* cyc; cudaMalloc(&out, 4 * 2048 * 148); cudaMalloc(&cyc, 8);
  for (int warps : {4, 16, 32}) {   // 1, 4, 8 warps per SMSP
    k<OP><<<148, warps * 32>>>(out, cyc); cudaDeviceSynchronize();
    k<OP><<<148, warps * 32>>>(out, cyc); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double instr_per_smsp = (double)N_IT * per_it * (warps / 4);
    printf("%-12s warps/SMSP=%d  cycles=%lld  cyc/warp-instr/SMSP=%.2f\n", name, warps / 4, c, c / instr_per_smsp);
  }
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("FFMA", UNR); run<1>("F2FP.rz.relu", UNR); run<9>("F2FP.rn", UNR); run<2>("PRMT", UNR); run<3>("FMNMX", UNR);
  run<4>("LOP3", UNR); run<5>("SHF(2op)", UNR); run<6>("FADD", UNR); run<7>("FFMA2", UNR / 2); run<8>("EX2+FMUL", UNR);
  return 0;
}
