// tcgen05 / TMEM kernel for the headline flow shape: RealNVP with an MLP
// conditioner  d_id -> 64 -> 64 -> 2*d_tr  (ReLU), D <= 16, up to 8 coupling layers.
//
// Per 128-row tile one thread owns one row (TMEM lane == row).  The three
// conditioner GEMMs of every coupling layer run on the 5th-gen tensor cores:
//   A (activations)  : TENSOR MEMORY (tcgen05.mma with the A operand in TMEM), written
//                      by the row's own thread with tcgen05.st as bf16 hi/lo halves
//                      (split-bf16: hi*Whi + lo*Whi + hi*Wlo recovers ~16 mantissa
//                      bits -> fp32-grade parity with the reference's fp32 CPU flow)
//   B (weights)      : shared memory (K-major, no-swizzle canonical layout), resident
//                      for the whole kernel, all layers
//   D (accumulator)  : TMEM, fp32, read back with tcgen05.ld by the row's thread
// Activations never touch shared memory, so TC_NG tiles are in flight per SM (one
// 128-thread epilogue group + one single-thread MMA issuer warp each, 128 TMEM
// columns per group): one group's tensor-core and barrier latency hides behind the
// other groups' CUDA-core epilogues (ReLU + split, coupling, 16x16 affine).
//
// Replaces, for this shape, the generic interpreter in flow_interp.cuh -- i.e. the
// reference's NFlow.inverse / forward (flows/base.py:209-221) over nflows'
// AffineCouplingTransform + nessai.flows.nets.MLP (nets.py:83-126).
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "flow_program.h"
#include "philox.cuh"
#include "populate_common.cuh"

namespace nb200 {

// ------------------------------------------------------------------ image layout
constexpr int TC_H = 64;            // conditioner width
constexpr int TC_DP = 16;           // padded feature count (registers per row)
constexpr int TC_N3 = 16;           // padded 2*d_tr
constexpr int TC_TR0 = 8;           // register slot of the first transformed feature
constexpr int TC_MAXL = 8;
// NB200_TC_AFFMMA: the D x D affine in front of every coupling rides GEMM1 as 16 extra output
// columns (B operand [16 x 80] = W1' | A^T, same split-fp16 passes): the row's new state comes
// back with one tcgen05.ld instead of 128 FFMA2 + 64 broadcast LDS.128 per row and layer.
// (measured on B200, C2: 0.293 -> 0.247 ms per 1e6 rows; -DNB200_TC_NO_AFFMMA keeps the fp32 FFMA2 affine)
#ifdef NB200_TC_NO_AFFMMA
constexpr bool TC_AFFMMA = false;
#else
constexpr bool TC_AFFMMA = true;
#endif
constexpr int TC_N1 = TC_H + (TC_AFFMMA ? TC_DP : 0);  // GEMM1 output columns
constexpr int TC_W1_BYTES = TC_N1 * 16 * 2;           // 2 K-chunks x (N1 rows x 16 B): K = 16 state slots
constexpr int TC_W2_BYTES = TC_H * 16 * (TC_H / 8);   // 8 chunks x 1 KB
constexpr int TC_W3_BYTES = TC_N3 * 16 * (TC_H / 8);  // 8 chunks x 256 B
// Biases enter as ONE extra K=16 step against a constant "ones" A operand whose elements
// k = 0, 1 are 1: the B operand carries bias_hi at k = 0 and bias_lo at k = 1 (K-chunk 0 only;
// its K-chunk 1 is a shared all-zero chunk reached through the descriptor's LBO).
constexpr int TC_B1_BYTES = TC_N1 * 16;
constexpr int TC_B2_BYTES = TC_H * 16;
constexpr int TC_B3_BYTES = TC_N3 * 16;
constexpr int TC_LAYER_BYTES =
    2 * (TC_W1_BYTES + TC_W2_BYTES + TC_W3_BYTES) + TC_B1_BYTES + TC_B2_BYTES + TC_B3_BYTES;
constexpr int TC_ONES_BYTES = 2 * 2048;               // A operand [128 x 16] bf16, shared by all groups
constexpr int TC_ZERO_BYTES = TC_N1 * 16;             // all-zero K-chunk 1 of the bias operands
constexpr int TC_OFF_W1HI = 0;
constexpr int TC_OFF_W1LO = TC_W1_BYTES;
constexpr int TC_OFF_W2HI = 2 * TC_W1_BYTES;
constexpr int TC_OFF_W2LO = TC_OFF_W2HI + TC_W2_BYTES;
constexpr int TC_OFF_W3HI = TC_OFF_W2LO + TC_W2_BYTES;
constexpr int TC_OFF_W3LO = TC_OFF_W3HI + TC_W3_BYTES;
constexpr int TC_OFF_B1 = TC_OFF_W3LO + TC_W3_BYTES;
constexpr int TC_OFF_B2 = TC_OFF_B1 + TC_B1_BYTES;
constexpr int TC_OFF_B3 = TC_OFF_B2 + TC_B2_BYTES;
constexpr int TC_AFF_BYTES = (TC_DP * TC_DP + TC_DP) * 4;
// TMEM columns of one epilogue group (one 128-row tile in flight)
constexpr int TC_COLS = 128;
constexpr int TC_COL_D = 0;     // accumulator, 64 fp32 columns
constexpr int TC_COL_AH = 64;   // activations hi: 64 bf16 = 32 columns (A1 hi aliases the first 8)
constexpr int TC_COL_AL = 96;   // activations lo
// GEMM1's A operand (the 16 state slots, 8 + 8 columns).  With the affine on the tensor core
// GEMM1's accumulator is 80 columns wide, so its A operand moves to the top of the group's columns.
constexpr int TC_COL_A1H = TC_AFFMMA ? 112 : TC_COL_AH;
constexpr int TC_COL_A1L = TC_AFFMMA ? 120 : TC_COL_AL;
constexpr int TC_COL_AFF = TC_H;  // (TC_AFFMMA) the affine's 16 output columns of GEMM1
#ifndef NB200_TC_NG
#define NB200_TC_NG 4
#endif
constexpr int TC_NG = NB200_TC_NG;
// NB200_TC_SELFISSUE: no issuer warps -- the first warp of every epilogue group issues the group's
// MMAs itself right after the group's named barrier (one mbarrier hop less per GEMM round trip).
#ifdef NB200_TC_SELFISSUE
constexpr bool TC_SELF = true;
#else
constexpr bool TC_SELF = false;
#endif
constexpr int TC_THREADS = TC_NG * 128 + (TC_SELF ? 0 : TC_NG * 32);
constexpr int TC_ALLOC_WARP = TC_SELF ? 0 : TC_NG * 4;  // the warp that owns the TMEM allocation
constexpr int TC_TMEM_COLS = (TC_NG * TC_COLS <= 128) ? 128 : (TC_NG * TC_COLS <= 256 ? 256 : 512);

struct TcProgram {
  bool valid = false;
  uint8_t* d_image = nullptr;
  int image_bytes = 0;
  int L = 0, D = 0;
  int d_id[TC_MAXL] = {0}, d_tr[TC_MAXL] = {0};
  int additive = 0, inverse = 0;
  int narrow = 0;  // conditioner width <= 32: the hidden GEMMs run two K-steps, the epilogues 32 columns
  int act = ACT_RELU;  // conditioner activation (a compile-time parameter of the kernels)
  float const_logdet = 0.f;
};

struct TcParams {
  const uint8_t* image;
  int image_bytes;
  int L, D;
  int d_id[TC_MAXL], d_tr[TC_MAXL];
  int additive, inverse;
  int narrow;
  float const_logdet;
};

inline void tc_free(TcProgram& t) {
  if (t.d_image) cudaFree(t.d_image);
  t = TcProgram();
}

inline int& tc_enabled_flag() {
  static int v = -1;
  return v;
}
inline bool tc_enabled() {
  int& v = tc_enabled_flag();
  if (v < 0) {
    const char* e = getenv("NB200_DISABLE_TC");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

// ---- 16-bit operand format of the split.  fp16 (default): hi + lo carry 11 + 11 significand
// bits, so hi*Whi + lo*Whi + hi*Wlo is accurate to ~2^-21 -- fp32 class -- at the cost of
// fp16's range: |value| >= 65504 cannot be represented (weights: the flow falls back to the
// generic kernel; activations: the row comes out NaN and is dropped like any other numerical
// failure).  -DNB200_TC_BF16 selects bf16 (8 + 8 bits, ~1e-5, fp32 range).
#ifdef NB200_TC_BF16
constexpr uint16_t TC_ONE16 = 0x3F80;
constexpr uint32_t TC_IDESC_FMT = (1u << 7) | (1u << 10);  // a_format = b_format = BF16
inline uint16_t tc_h16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;  // NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
inline float tc_h16_to_f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
inline bool tc_h16_representable(float f) { return std::isfinite(f); }
#else
constexpr uint16_t TC_ONE16 = 0x3C00;
constexpr uint32_t TC_IDESC_FMT = 0u;  // a_format = b_format = F16
inline uint16_t tc_h16_rn(float f) {  // float -> IEEE half, round to nearest even
  uint32_t x;
  memcpy(&x, &f, 4);
  const uint16_t sign = (uint16_t)((x >> 16) & 0x8000u);
  x &= 0x7fffffffu;
  if (x > 0x7f800000u) return sign | 0x7e00;
  if (x >= 0x477ff000u) return sign | 0x7c00;  // >= 65520 rounds to infinity
  if (x < 0x38800000u) {                       // below 2^-14: subnormal half = round(f * 2^24)
    float a;
    memcpy(&a, &x, 4);
    return sign | (uint16_t)lrintf(a * 16777216.0f);
  }
  uint32_t h = (((x >> 23) - 112u) << 10) | ((x & 0x7fffffu) >> 13);
  const uint32_t rem = x & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;
  return sign | (uint16_t)h;
}
inline float tc_h16_to_f(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
  float f;
  if (e == 0) {
    f = (float)m * 5.9604644775390625e-8f;  // 2^-24
    uint32_t u;
    memcpy(&u, &f, 4);
    u |= sign;
    memcpy(&f, &u, 4);
    return f;
  }
  const uint32_t u = sign | (e == 31 ? 0x7f800000u | (m << 13) : ((e + 112u) << 23) | (m << 13));
  memcpy(&f, &u, 4);
  return f;
}
inline bool tc_h16_representable(float f) { return std::fabs(f) < 65504.f; }
#endif
inline bool& tc_put_overflow() {  // set when a weight does not fit the operand format
  static thread_local bool v = false;
  return v;
}
// write element (n, k) of a K-major canonical operand with `rows` rows per chunk
inline void tc_put(uint8_t* base_hi, uint8_t* base_lo, int rows, int n, int k, float w) {
  if (!tc_h16_representable(w)) tc_put_overflow() = true;
  const uint16_t hi = tc_h16_rn(w);
  const uint16_t lo = tc_h16_rn(w - tc_h16_to_f(hi));
  const size_t off = (size_t)(k / 8) * rows * 16 + (size_t)n * 16 + (k % 8) * 2;
  memcpy(base_hi + off, &hi, 2);
  memcpy(base_lo + off, &lo, 2);
}

// Recognise  affine (coupling affine)*  with the supported conditioner shape and
// build the shared-memory image.  Returns 0 (t.valid says whether it applies).
inline int tc_build(TcProgram& t, const FlowOp* ops, int n_ops, const float* blob, int D, int H,
                    int activation, int final_buf) {
  t.valid = false;
  (void)final_buf;
  if (D > TC_DP || H < 1 || H > TC_H || activation < ACT_RELU || activation > ACT_SILU) return 0;
  if (n_ops < 5 || (n_ops - 1) % 4 != 0) return 0;
  const int L = (n_ops - 1) / 4;
  if (L > TC_MAXL) return 0;
  auto is_affine = [&](const FlowOp& o) {
    return o.type == OP_LINEAR && o.src <= BUF_X1 && o.dst <= BUF_X1 && o.K == D && o.N == D &&
           o.flags == 0 && o.src_off == 0;
  };
  if (!is_affine(ops[0])) return 0;
  int inverse = -1, additive = -1;
  for (int l = 0; l < L; ++l) {
    const FlowOp& a = ops[1 + 4 * l];
    const FlowOp& b = ops[2 + 4 * l];
    const FlowOp& c = ops[3 + 4 * l];
    const FlowOp& f = ops[4 + 4 * l];
    if (a.type != OP_LINEAR || a.src > BUF_X1 || a.dst < BUF_A0 || a.N != H ||
        a.flags != FLAG_OUT_ACT || a.src_off != 0 || a.K < 1 || a.K > TC_TR0)
      return 0;
    if (b.type != OP_LINEAR || b.src < BUF_A0 || b.dst < BUF_A0 || b.K != H || b.N != H ||
        b.flags != FLAG_OUT_ACT)
      return 0;
    if (c.type != OP_COUPLING_AFFINE || c.K != H || c.d_id != a.K || c.d_tr < 1 ||
        2 * c.d_tr > TC_N3 || c.d_id + c.d_tr != D || c.N != 2 * c.d_tr)
      return 0;
    const int inv = (c.flags & FLAG_INVERSE) ? 1 : 0, add = (c.flags & FLAG_ADDITIVE) ? 1 : 0;
    if ((c.flags & ~(FLAG_INVERSE | FLAG_ADDITIVE)) != 0 || c.x_buf != c.dst) return 0;
    if ((inverse >= 0 && inverse != inv) || (additive >= 0 && additive != add)) return 0;
    inverse = inv;
    additive = add;
    if (!is_affine(f)) return 0;
    t.d_id[l] = c.d_id;
    t.d_tr[l] = c.d_tr;
  }
  tc_put_overflow() = false;
  const int ones_off = L * TC_LAYER_BYTES + (L + 1) * TC_AFF_BYTES;
  const int bytes = ones_off + TC_ONES_BYTES + TC_ZERO_BYTES;
  if (((bytes + 1023) & ~1023) + 4096 > 227 * 1024) return 0;  // does not fit: generic kernel
  std::vector<uint8_t> img((size_t)bytes, 0);
  for (int m = 0; m < 128; ++m) {  // elements (m, k = 0) and (m, k = 1) are 1.0 (bf16 0x3F80)
    const uint16_t one[2] = {TC_ONE16, TC_ONE16};
    memcpy(img.data() + ones_off + (size_t)m * 16, one, 4);
  }
  auto slot = [&](int layer, int j) {  // register slot of natural feature j; layer < 0 or >= L: natural
    if (layer < 0 || layer >= L) return j;
    return j < t.d_id[layer] ? j : TC_TR0 + (j - t.d_id[layer]);
  };
  // bias operand: (n, k = 0) = hi, (n, k = 1) = lo
  auto put_bias = [&](uint8_t* base, int n, float b) {
    if (!tc_h16_representable(b)) tc_put_overflow() = true;
    const uint16_t hi = tc_h16_rn(b);
    const uint16_t lo = tc_h16_rn(b - tc_h16_to_f(hi));
    memcpy(base + (size_t)n * 16, &hi, 2);
    memcpy(base + (size_t)n * 16 + 2, &lo, 2);
  };
  for (int l = 0; l < L; ++l) {
    uint8_t* lb = img.data() + (size_t)l * TC_LAYER_BYTES;
    const FlowOp& f = ops[4 * l];  // the affine in front of coupling l
    const FlowOp& a = ops[1 + 4 * l];
    const FlowOp& b = ops[2 + 4 * l];
    const FlowOp& c = ops[3 + 4 * l];
    // GEMM1 consumes the state BEFORE the affine (so the fp32 affine runs on the CUDA cores
    // while the tensor core works): W1' = W1 A[:, identity], b1' = b1 + W1 b[identity], float64
    for (int n = 0; n < H; ++n) {
      for (int k = 0; k < D; ++k) {
        double acc = 0.0;
        for (int j = 0; j < a.K; ++j)
          acc += (double)blob[a.w_off + j * a.Npad + n] * (double)blob[f.w_off + k * f.Npad + j];
        tc_put(lb + TC_OFF_W1HI, lb + TC_OFF_W1LO, TC_N1, n, slot(l - 1, k), (float)acc);
      }
      double bacc = blob[a.b_off + n];
      for (int j = 0; j < a.K; ++j)
        bacc += (double)blob[a.w_off + j * a.Npad + n] * (double)blob[f.b_off + j];
      put_bias(lb + TC_OFF_B1, n, (float)bacc);
    }
    if (TC_AFFMMA) {
      // rows 64 .. 79 of GEMM1's B operand: the affine itself, output slot n_s = slot(l, n) from
      // input slot slot(l - 1, k); its bias joins b1'
      for (int n = 0; n < D; ++n) {
        for (int k = 0; k < D; ++k)
          tc_put(lb + TC_OFF_W1HI, lb + TC_OFF_W1LO, TC_N1, TC_H + slot(l, n), slot(l - 1, k),
                 blob[f.w_off + k * f.Npad + n]);
        put_bias(lb + TC_OFF_B1, TC_H + slot(l, n), blob[f.b_off + n]);
      }
    }
    for (int n = 0; n < H; ++n) {
      for (int k = 0; k < H; ++k)
        tc_put(lb + TC_OFF_W2HI, lb + TC_OFF_W2LO, TC_H, n, k, blob[b.w_off + k * b.Npad + n]);
    }
    for (int n = 0; n < c.N; ++n) {
      for (int k = 0; k < H; ++k)
        tc_put(lb + TC_OFF_W3HI, lb + TC_OFF_W3LO, TC_N3, n, k, blob[c.w_off + k * c.Npad + n]);
    }
    for (int n = 0; n < H; ++n) put_bias(lb + TC_OFF_B2, n, blob[b.b_off + n]);
    for (int n = 0; n < c.N; ++n) put_bias(lb + TC_OFF_B3, n, blob[c.b_off + n]);
  }
  // Affines, re-laid-out to the kernel's register slots: inside coupling layer l the
  // identity features live in slots [0, d_id) and the transformed ones in
  // [TC_TR0, TC_TR0 + d_tr); the flow's input / output use the natural order.
  for (int i = 0; i <= L; ++i) {
    const FlowOp& f = ops[4 * i];
    float* A = reinterpret_cast<float*>(img.data() + (size_t)L * TC_LAYER_BYTES +
                                        (size_t)i * TC_AFF_BYTES);
    for (int k = 0; k < D; ++k)
      for (int n = 0; n < D; ++n)
        A[slot(i - 1, k) * TC_DP + slot(i, n)] = blob[f.w_off + k * f.Npad + n];
    for (int n = 0; n < D; ++n) A[TC_DP * TC_DP + slot(i, n)] = blob[f.b_off + n];
  }
  if (tc_put_overflow()) return 0;  // a weight outside the operand format's range: generic kernel
  if (cudaMalloc(&t.d_image, bytes) != cudaSuccess) return 2;
  if (cudaMemcpy(t.d_image, img.data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess) return 2;
  t.image_bytes = bytes;
  t.L = L;
  t.D = D;
  t.inverse = inverse;
  t.additive = additive;
  t.narrow = H <= TC_H / 2;
  t.act = activation;
  t.valid = true;
  return 0;
}

// ------------------------------------------------------------------ device helpers
__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void tc_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void tc_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity) {
  // try_wait suspends the thread in hardware (up to the tick hint) instead of spinning
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "TC_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra TC_DONE_%=;\n\t"
      "bra TC_WAIT_%=;\n\t"
      "TC_DONE_%=:\n\t}"
      ::"r"(bar), "r"(parity), "r"(0x989680u)
      : "memory");
}
// The weight image into shared memory as bulk asynchronous copies (the TMA engine: cp.async.bulk,
// SASS UBLKCP) that complete on an mbarrier -- no register staging, the threads go on to set up
// barriers / tensor memory meanwhile.  ONE thread calls this; every thread that reads the image
// (or issues MMAs on it) waits on `bar` (phase 0) after the block barrier that publishes its init.
__device__ __forceinline__ void tc_image_load(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  tc_mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  const char* src = reinterpret_cast<const char*>(gsrc);
  for (uint32_t off = 0; off < bytes; off += 32768u) {
    const uint32_t n = bytes - off < 32768u ? bytes - off : 32768u;
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst + off),
        "l"(src + off), "r"(n), "r"(bar)
        : "memory");
  }
}

__device__ __forceinline__ void tc_fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// K-major, no swizzle: LBO = bytes between the two 8-element K chunks of one MMA,
// SBO = bytes between 8-row groups (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: bf16 x bf16 -> f32, both operands K-major
__host__ __device__ constexpr uint32_t tc_idesc(int M, int N) {
  return (1u << 4) | TC_IDESC_FMT | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void tc_mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]   (A operand in tensor memory)
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The issuer warp runs CONVERGED and every tcgen05.mma / commit is predicated by elect.sync
// inside the asm: with warp-uniform operands ptxas keeps descriptors in uniform registers and
// issues back to back (measured 32.5 cycles per 128x64x16 MMA = the tensor-pipe floor).  Issued
// from a divergent single-lane branch instead, every MMA is wrapped in an ELECT /
// R2UR.BROADCAST waterfall loop (47-122 cycles each, scripts/ubench/mma_rate3.cu).
__device__ __forceinline__ void tc_mma_ss_e(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_ts_e(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_e(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// pin the loaded registers behind the preceding tcgen05.wait::ld (no instructions)
__device__ __forceinline__ void tc_pin16(uint32_t (&r)[16]) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]),
               "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]),
               "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])::"memory");
}
// packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 issue two lanes per instruction)
__device__ __forceinline__ void tc_fma2(float& d0, float& d1, float a0, float a1, float b0,
                                        float b1) {
  asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%0, %1};\n\t"
      "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "+f"(d0), "+f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void tc_sub2(float& d0, float& d1, float a0, float a1, float b0,
                                        float b1) {
  asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "sub.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
// split two fp32 into packed bf16x2 hi (truncated) and lo (remainder, rounded);
// element `a` goes to the low half-word.  RELU clamps negatives of both parts.
template <bool RELU>
__device__ __forceinline__ void tc_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
#ifdef NB200_TC_BF16
  if (RELU)
    asm("cvt.rz.relu.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  else
    asm("cvt.rz.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  float ra, rb;
  tc_sub2(ra, rb, a, b, __uint_as_float(hi << 16), __uint_as_float(hi & 0xffff0000u));
  if (RELU)
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
  else
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
#else
  // hi truncated toward zero, so the remainder has the sign of the value and ReLU on both
  // parts is ReLU on the sum
  if (RELU)
    asm("cvt.rz.relu.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  else
    asm("cvt.rz.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  float ha, hb;
  asm("{\n\t.reg .f16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}"
      : "=f"(ha), "=f"(hb)
      : "r"(hi));
  float ra, rb;
  tc_sub2(ra, rb, a, b, ha, hb);
  if (RELU)
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
  else
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
#endif
}

// raw SFU ops (no denormal / range fix-up code around them: operands here are O(1))
__device__ __forceinline__ float tc_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tc_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tc_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// The conditioner's activation fused into the split: ReLU is free (cvt .relu); tanh and SiLU
// (flows/utils.py:24-32, 200-205) cost two SFU operations per value -- ex2 + rcp, absolute error
// ~2e-7 -- before the split.  ACT: ACT_RELU / ACT_TANH / ACT_SILU of flow_program.h, or TC_ACT_NONE.
constexpr int TC_ACT_NONE = -1;
template <int ACT>
__device__ __forceinline__ float tc_act(float v) {
  if (ACT == ACT_TANH) return 1.f - 2.f * tc_rcp(1.f + tc_ex2(v * 2.8853900817779268f));
  if (ACT == ACT_SILU) return v * tc_rcp(1.f + tc_ex2(v * -1.4426950408889634f));
  return v;
}
template <int ACT>
__device__ __forceinline__ void tc_split_act2(float a, float b, uint32_t& hi, uint32_t& lo) {
  if (ACT == ACT_TANH || ACT == ACT_SILU) tc_split2<false>(tc_act<ACT>(a), tc_act<ACT>(b), hi, lo);
  else tc_split2<ACT == ACT_RELU>(a, b, hi, lo);
}

// The affine coupling on the 8 transformed slots h[TC_TR0 ..], branch-free so that the eight
// SFU chains (ex2 -> rcp -> lg2 / rcp) interleave instead of running one after the other:
// r[2f] = shift, r[2f+1] = unconstrained scale, scale = sigmoid(u + 2) + 1e-3.  Slots >= d_tr
// carry zero weights (their h stays 0); only their log-scale has to be masked.
template <bool INVERSE>
__device__ __forceinline__ float tc_coupling8(const uint32_t (&r)[16], float (&h)[TC_DP], int d_tr) {
  float s[8];
#pragma unroll
  for (int f = 0; f < 8; ++f)
    s[f] = tc_ex2((__uint_as_float(r[2 * f + 1]) + 2.f) * -1.4426950408889634f);
#pragma unroll
  for (int f = 0; f < 8; ++f) s[f] = tc_rcp(1.f + s[f]) + 1e-3f;
  float ld = 0.f;
#pragma unroll
  for (int f = 0; f < 8; ++f) {
    const float tt = __uint_as_float(r[2 * f]);
    const float ls = tc_lg2(s[f]) * 0.6931471805599453f;
    if (INVERSE) h[TC_TR0 + f] = (h[TC_TR0 + f] - tt) * tc_rcp(s[f]);
    else h[TC_TR0 + f] = fmaf(h[TC_TR0 + f], s[f], tt);
    ld += f < d_tr ? ls : 0.f;
  }
  return INVERSE ? -ld : ld;
}
// volume-preserving (additive) coupling: scale == 1
template <bool INVERSE>
__device__ __forceinline__ void tc_coupling8_additive(const uint32_t (&r)[16], float (&h)[TC_DP]) {
#pragma unroll
  for (int f = 0; f < 8; ++f) {
    const float tt = __uint_as_float(r[2 * f]);
    h[TC_TR0 + f] = INVERSE ? h[TC_TR0 + f] - tt : h[TC_TR0 + f] + tt;
  }
}
__device__ __forceinline__ float tc_coupling(const uint32_t (&r)[16], float (&h)[TC_DP], int d_tr,
                                             int additive, int inverse) {
  if (additive) {
    if (inverse) tc_coupling8_additive<true>(r, h);
    else tc_coupling8_additive<false>(r, h);
    return 0.f;
  }
  return inverse ? tc_coupling8<true>(r, h, d_tr) : tc_coupling8<false>(r, h, d_tr);
}

// hidden-layer epilogue: 64 accumulator columns (bias already accumulated by the
// GEMM) -> ReLU -> split -> the row's A operand in TMEM (hi: 32 columns, lo: 32).
// NQ: groups of 16 columns that carry hidden units (4; 2 for a conditioner of width <= 32, whose
// other columns are the zero padding of the weight image and are not an operand of any MMA).
template <int NQ = 4, int ACT = ACT_RELU>
__device__ __forceinline__ void tc_hidden_epilogue(uint32_t tg) {
  uint32_t ra[16], rb[16];
  tc_ld16(tg + TC_COL_D, ra);
  tc_wait_ld();
  tc_pin16(ra);
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    uint32_t(&cur)[16] = (q & 1) ? rb : ra;
    uint32_t(&nxt)[16] = (q & 1) ? ra : rb;
    if (q < NQ - 1) tc_ld16(tg + TC_COL_D + 16 * (q + 1), nxt);  // in flight while `cur` is processed
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#ifdef NB200_ABL_NO_SPLIT
      hi[j] = cur[2 * j] & 0xffff0000u; lo[j] = cur[2 * j + 1];
#else
      tc_split_act2<ACT>(__uint_as_float(cur[2 * j]), __uint_as_float(cur[2 * j + 1]), hi[j], lo[j]);
#endif
    }
    tc_st8(tg + TC_COL_AH + 8 * q, hi);
    tc_st8(tg + TC_COL_AL + 8 * q, lo);
    if (q < NQ - 1) {
      tc_wait_ld();
      tc_pin16(nxt);
    }
  }
  tc_wait_st();
}

// h <- A h + b  with A k-major [16][16] fp32 in shared memory (warp-wide broadcasts)
__device__ __forceinline__ void tc_affine(const float* __restrict__ A, float (&h)[TC_DP]) {
  float o[TC_DP];
  const float4* b4 = reinterpret_cast<const float4*>(A + TC_DP * TC_DP);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 b = b4[j];
    o[4 * j] = b.x, o[4 * j + 1] = b.y, o[4 * j + 2] = b.z, o[4 * j + 3] = b.w;
  }
#pragma unroll
  for (int k = 0; k < TC_DP; ++k) {
    const float4* w4 = reinterpret_cast<const float4*>(A + k * TC_DP);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 w = w4[j];
      tc_fma2(o[4 * j + 0], o[4 * j + 1], w.x, w.y, h[k], h[k]);
      tc_fma2(o[4 * j + 2], o[4 * j + 3], w.z, w.w, h[k], h[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < TC_DP; ++k) h[k] = o[k];
}

#ifdef NB200_TC_TRACE
// Debug build only: clock64 stamps of CTA 0 (epilogue group 0 thread 0: even slots; issuer 0:
// odd slots) to find where a tile-layer's latency goes.  Read back with nb200_debug_trace().
__device__ long long tc_trace_buf[2][4096];
__device__ int tc_trace_n[2];
__device__ __forceinline__ void tc_stamp(int who, int tag) {
  if (blockIdx.x == 0) {
    const int i = tc_trace_n[who];
    if (i < 4096) {
      tc_trace_buf[who][i] = (clock64() << 8) | tag;
      tc_trace_n[who] = i + 1;
    }
  }
}
#define TC_STAMP_E(tag) do { if (threadIdx.x == 0) tc_stamp(0, tag); } while (0)
#define TC_STAMP_I(tag) do { if (threadIdx.x == TC_NG * 128) tc_stamp(1, tag); } while (0)
#else
#define TC_STAMP_E(tag) do { } while (0)
#define TC_STAMP_I(tag) do { } while (0)
#endif

struct TcIO {
  // apply mode
  const float* in;
  float* out;
  float* out_logj;
  float* out_lp;
  int lp_mode;  // 1: inverse (base(in) - logj), 2: forward (base(out) + logj)
  int64_t n;
  // base distribution N(0, var I) (flows/distributions.py:17-73; var = 1: StandardNormal):
  // log p(z) = -0.5 * base_inv_var * |z|^2 - base_log_z,  base_log_z = 0.5 D log(2 pi var)
  float base_inv_var = 1.f;
  float base_log_z = 0.f;
};

// The MMAs of one conditioner GEMM of layer l (which = 1, 2, 3) for the group whose TMEM base is
// tgu (lane field 0, warp-uniform); the WHOLE warp runs this converged, one elected lane fires.
template <int NKS>
__device__ __forceinline__ void tc_issue_gemm_n(const TcParams& P, uint32_t img_s, uint32_t tgu, int l,
                                                int which, uint32_t bar_out) {
  constexpr uint32_t ID64 = tc_idesc(128, TC_H), ID16 = tc_idesc(128, TC_N3);
  const uint32_t d = tgu + TC_COL_D, ah = tgu + TC_COL_AH, al = tgu + TC_COL_AL;
  const uint32_t ones_s = img_s + P.L * TC_LAYER_BYTES + (P.L + 1) * TC_AFF_BYTES;
  const uint32_t zero_s = ones_s + TC_ONES_BYTES;
  const uint64_t ones = tc_desc(ones_s, 2048, 128);
  auto adv = [](uint64_t desc, uint32_t off) { return desc + (uint64_t)(off >> 4); };
  const uint32_t lb = img_s + l * TC_LAYER_BYTES;
  constexpr int nks = NKS;  // K-steps of 16 hidden units (4; 2 for a conditioner of width <= 32)
  if (which == 1) {
    constexpr uint32_t ID1 = tc_idesc(128, TC_N1);
    const uint32_t a1h = tgu + TC_COL_A1H, a1l = tgu + TC_COL_A1L;
    const uint64_t d1 = tc_desc(lb, TC_N1 * 16, 128);
    const uint64_t b1 = tc_desc(lb + TC_OFF_B1, zero_s - (lb + TC_OFF_B1), 128);
    tc_mma_ss_e(d, ones, b1, ID1, 0);
    tc_mma_ts_e(d, a1h, adv(d1, TC_OFF_W1HI), ID1, 1);
    tc_mma_ts_e(d, a1l, adv(d1, TC_OFF_W1HI), ID1, 1);
    tc_mma_ts_e(d, a1h, adv(d1, TC_OFF_W1LO), ID1, 1);
  } else if (which == 2) {
    const uint64_t d64 = tc_desc(lb, TC_H * 16, 128);
    const uint64_t b2 = tc_desc(lb + TC_OFF_B2, zero_s - (lb + TC_OFF_B2), 128);
    tc_mma_ss_e(d, ones, b2, ID64, 0);
#pragma unroll
    for (int ks = 0; ks < nks; ++ks) {
      tc_mma_ts_e(d, ah + 8 * ks, adv(d64, TC_OFF_W2HI + ks * 2 * TC_H * 16), ID64, 1);
      tc_mma_ts_e(d, al + 8 * ks, adv(d64, TC_OFF_W2HI + ks * 2 * TC_H * 16), ID64, 1);
      tc_mma_ts_e(d, ah + 8 * ks, adv(d64, TC_OFF_W2LO + ks * 2 * TC_H * 16), ID64, 1);
    }
  } else {
    const uint64_t d16 = tc_desc(lb, TC_N3 * 16, 128);
    const uint64_t b3 = tc_desc(lb + TC_OFF_B3, zero_s - (lb + TC_OFF_B3), 128);
    tc_mma_ss_e(d, ones, b3, ID16, 0);
#pragma unroll
    for (int ks = 0; ks < nks; ++ks) {
      tc_mma_ts_e(d, ah + 8 * ks, adv(d16, TC_OFF_W3HI + ks * 2 * TC_N3 * 16), ID16, 1);
      tc_mma_ts_e(d, al + 8 * ks, adv(d16, TC_OFF_W3HI + ks * 2 * TC_N3 * 16), ID16, 1);
      tc_mma_ts_e(d, ah + 8 * ks, adv(d16, TC_OFF_W3LO + ks * 2 * TC_N3 * 16), ID16, 1);
    }
  }
  tc_commit_e(bar_out);
}

// Hand a GEMM's A operand over to the tensor core.  Default: arrive at the issuer warp's
// mbarrier.  Self-issue: the group's named barrier, then the group's first warp issues.
struct TcGroup {
  uint32_t bar_in, bar_out;
  uint32_t img_s;   // shared-memory address of the weight image
  uint32_t tgu;     // the group's TMEM base (lane field 0), warp-uniform
  int g;            // group index
  bool first_warp;  // warp-uniform: this warp issues the group's MMAs (self-issue only)
};
template <bool NARROW>
__device__ __forceinline__ void tc_submit(const TcParams& P, const TcGroup& G, int l, int which) {
  if (TC_SELF) {
    asm volatile("bar.sync %0, 128;" ::"r"(1 + G.g) : "memory");
    if (G.first_warp) {
      tc_fence_after();
      tc_issue_gemm_n<NARROW ? 2 : 4>(P, G.img_s, G.tgu, l, which, G.bar_out);
    }
  } else {
    tc_mbar_arrive(G.bar_in);
  }
}

// The epilogue-group body shared by the apply and populate kernels: runs the whole
// program for one row held in h[] and returns the row log|det J| (without const).
// tg: the group's TMEM base with this warp's lane quarter in the upper half-word.
template <bool NARROW, int ACT>
__device__ __forceinline__ float tc_run_row(const TcParams& P, const uint8_t* img, uint32_t tg,
                                            const TcGroup& G, uint32_t& ph_out, float (&h)[TC_DP]) {
  const uint32_t bar_out = G.bar_out;
  const float* aff = reinterpret_cast<const float*>(img + (size_t)P.L * TC_LAYER_BYTES);
  float ld = 0.f;
  for (int l = 0; l < P.L; ++l) {
    const int d_tr = P.d_tr[l];
    // ---- E0: the 16 state slots BEFORE this layer's affine -> A1 (16 bf16 = 8 columns,
    //      hi and lo); GEMM1 carries the affine folded into its weights
    {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) tc_split2<false>(h[2 * j], h[2 * j + 1], hi[j], lo[j]);
      TC_STAMP_E(9);
      tc_st8(tg + TC_COL_A1H, hi);
      tc_st8(tg + TC_COL_A1L, lo);
      tc_wait_st();
    }
    tc_fence_before();
    TC_STAMP_E(1);
    tc_submit<NARROW>(P, G, l, 1);
    // the fp32 affine in front of the coupling, in the shadow of GEMM1
#ifndef NB200_ABL_NO_AFFINE
    if (!TC_AFFMMA) tc_affine(aff + (size_t)l * (TC_AFF_BYTES / 4), h);
#endif
    // ---- E1: hidden layer 1
    tc_mbar_wait(bar_out, ph_out);
    TC_STAMP_E(2);
    ph_out ^= 1;
    tc_fence_after();
    if (TC_AFFMMA) {
      // the state after this layer's affine: GEMM1's columns 64 .. 79 (read before the hidden
      // epilogue overwrites them with the next A operand)
      uint32_t r[16];
      tc_ld16(tg + TC_COL_AFF, r);
      tc_wait_ld();
      tc_pin16(r);
#pragma unroll
      for (int d = 0; d < TC_DP; ++d) h[d] = __uint_as_float(r[d]);
    }
    tc_hidden_epilogue<NARROW ? 2 : 4, ACT>(tg);
    tc_fence_before();
    TC_STAMP_E(3);
    tc_submit<NARROW>(P, G, l, 2);
    // ---- E2: hidden layer 2
    tc_mbar_wait(bar_out, ph_out);
    TC_STAMP_E(4);
    ph_out ^= 1;
    tc_fence_after();
    tc_hidden_epilogue<NARROW ? 2 : 4, ACT>(tg);
    tc_fence_before();
    TC_STAMP_E(5);
    tc_submit<NARROW>(P, G, l, 3);
    // ---- E3: coupling on the transformed half, then the next affine
    tc_mbar_wait(bar_out, ph_out);
    TC_STAMP_E(6);
    ph_out ^= 1;
    tc_fence_after();
    uint32_t r[16];
    tc_ld16(tg + TC_COL_D, r);
    tc_wait_ld();
    tc_pin16(r);
    tc_fence_before();
    TC_STAMP_E(7);
    ld += tc_coupling(r, h, d_tr, P.additive, P.inverse);
    TC_STAMP_E(8);
  }
  tc_affine(aff + (size_t)P.L * (TC_AFF_BYTES / 4), h);
  return ld;
}

// MMA issuer for one epilogue group: the WHOLE warp runs this (converged); descriptors are
// uniform arithmetic on the shared-memory image base.
template <int NKS>
__device__ __forceinline__ void tc_issuer_n(const TcParams& P, uint32_t img_s, uint32_t tg,
                                            uint32_t bar_in, uint32_t bar_out, int64_t my_tiles) {
  constexpr uint32_t ID64 = tc_idesc(128, TC_H), ID16 = tc_idesc(128, TC_N3);
  const uint32_t d = tg + TC_COL_D, ah = tg + TC_COL_AH, al = tg + TC_COL_AL;
  const uint32_t ones_s = img_s + P.L * TC_LAYER_BYTES + (P.L + 1) * TC_AFF_BYTES;
  const uint32_t zero_s = ones_s + TC_ONES_BYTES;
  const uint64_t ones = tc_desc(ones_s, 2048, 128);
  // a descriptor `off` bytes further into the image: the start address field is bits [0, 14) >> 4
  auto adv = [](uint64_t desc, uint32_t off) { return desc + (uint64_t)(off >> 4); };
  constexpr int nks = NKS;  // K-steps of 16 hidden units (4; 2 for a conditioner of width <= 32)
  uint32_t ph_in = 0;
  for (int64_t it = 0; it < my_tiles; ++it) {
    for (int l = 0; l < P.L; ++l) {
      const uint32_t lb = img_s + l * TC_LAYER_BYTES;
      const uint64_t d64 = tc_desc(lb, TC_H * 16, 128);    // 64-row operands (W1, W2)
      const uint64_t d16 = tc_desc(lb, TC_N3 * 16, 128);   // 16-row operands (W3)
      // bias operands: K-chunk 1 is the shared zero chunk (LBO = distance to it)
      const uint64_t b1 = tc_desc(lb + TC_OFF_B1, zero_s - (lb + TC_OFF_B1), 128);
      const uint64_t b2 = tc_desc(lb + TC_OFF_B2, zero_s - (lb + TC_OFF_B2), 128);
      const uint64_t b3 = tc_desc(lb + TC_OFF_B3, zero_s - (lb + TC_OFF_B3), 128);
      // GEMM1: bias + [128 x 16] x [16 x 64]
      tc_mbar_wait(bar_in, ph_in);
      TC_STAMP_I(11);
      ph_in ^= 1;
      tc_fence_after();
      constexpr uint32_t ID1 = tc_idesc(128, TC_N1);
      const uint64_t d1 = tc_desc(lb, TC_N1 * 16, 128);
      tc_mma_ss_e(d, ones, b1, ID1, 0);
      tc_mma_ts_e(d, tg + TC_COL_A1H, adv(d1, TC_OFF_W1HI), ID1, 1);
      tc_mma_ts_e(d, tg + TC_COL_A1L, adv(d1, TC_OFF_W1HI), ID1, 1);
      tc_mma_ts_e(d, tg + TC_COL_A1H, adv(d1, TC_OFF_W1LO), ID1, 1);
      tc_commit_e(bar_out);
      TC_STAMP_I(12);
      // GEMM2: bias + [128 x 64] x [64 x 64]
      tc_mbar_wait(bar_in, ph_in);
      TC_STAMP_I(13);
      ph_in ^= 1;
      tc_fence_after();
      tc_mma_ss_e(d, ones, b2, ID64, 0);
#pragma unroll
      for (int ks = 0; ks < nks; ++ks) {
        tc_mma_ts_e(d, ah + 8 * ks, adv(d64, TC_OFF_W2HI + ks * 2 * TC_H * 16), ID64, 1);
        tc_mma_ts_e(d, al + 8 * ks, adv(d64, TC_OFF_W2HI + ks * 2 * TC_H * 16), ID64, 1);
        tc_mma_ts_e(d, ah + 8 * ks, adv(d64, TC_OFF_W2LO + ks * 2 * TC_H * 16), ID64, 1);
      }
      tc_commit_e(bar_out);
      TC_STAMP_I(14);
      // GEMM3: bias + [128 x 64] x [64 x 16]
      tc_mbar_wait(bar_in, ph_in);
      TC_STAMP_I(15);
      ph_in ^= 1;
      tc_fence_after();
      tc_mma_ss_e(d, ones, b3, ID16, 0);
#pragma unroll
      for (int ks = 0; ks < nks; ++ks) {
        tc_mma_ts_e(d, ah + 8 * ks, adv(d16, TC_OFF_W3HI + ks * 2 * TC_N3 * 16), ID16, 1);
        tc_mma_ts_e(d, al + 8 * ks, adv(d16, TC_OFF_W3HI + ks * 2 * TC_N3 * 16), ID16, 1);
        tc_mma_ts_e(d, ah + 8 * ks, adv(d16, TC_OFF_W3LO + ks * 2 * TC_N3 * 16), ID16, 1);
      }
      tc_commit_e(bar_out);
      TC_STAMP_I(16);
    }
  }
}

struct TcShared {
  uint64_t bar_in[TC_NG];
  uint64_t bar_out[TC_NG];
  uint64_t bar_img;  // completion of the weight image's bulk copy
  uint32_t tmem_base;
  uint32_t pad;
  double cst[4][TC_DP];  // populate: scale, shift, lo, hi
  double log_const;      // populate: D log sqrt(T) + sum log|scale|
};

__device__ __forceinline__ size_t tc_image_pad(int image_bytes) {
  return ((size_t)image_bytes + 1023) & ~(size_t)1023;
}

// common prologue: weights -> smem, barriers, TMEM
__device__ __forceinline__ void tc_prologue(const TcParams& P, uint8_t* smem, TcShared*& sh) {
  sh = reinterpret_cast<TcShared*>(smem + tc_image_pad(P.image_bytes));
  const int tid = threadIdx.x;
  if (tid == 0) {
    tc_image_load(tc_smem_u32(smem), P.image, (uint32_t)P.image_bytes, tc_smem_u32(&sh->bar_img));
    for (int g = 0; g < TC_NG; ++g) {
      tc_mbar_init(tc_smem_u32(&sh->bar_in[g]), 128);
      tc_mbar_init(tc_smem_u32(&sh->bar_out[g]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if ((tid >> 5) == TC_ALLOC_WARP) {  // this warp owns the TMEM allocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     tc_smem_u32(&sh->tmem_base)),
                 "r"((uint32_t)TC_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  tc_mbar_wait(tc_smem_u32(&sh->bar_img), 0);  // the weight image has landed
}

__device__ __forceinline__ void tc_epilogue_end(TcShared* sh) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == TC_ALLOC_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(sh->tmem_base),
                 "r"((uint32_t)TC_TMEM_COLS)
                 : "memory");
  }
}

inline size_t tc_smem_bytes(int image_bytes) {
  return (((size_t)image_bytes + 1023) & ~(size_t)1023) + sizeof(TcShared) + 64;
}

__device__ __forceinline__ int64_t tc_my_tiles(int64_t ntiles, int g) {
  const int64_t first = (int64_t)blockIdx.x * TC_NG + g;
  const int64_t stride = (int64_t)gridDim.x * TC_NG;
  return first < ntiles ? (ntiles - first + stride - 1) / stride : 0;
}

#define TC_LOG_2PI 1.8378770664093453f

// NARROW: conditioner width <= 32 (two K-steps in the hidden GEMMs, 32-column hidden epilogues); a
// compile-time switch so that each instantiation's hot loop holds one variant only (both variants
// inlined behind a run-time branch cost the wide kernels 6 - 27 %: instruction cache).
template <bool NARROW, int ACT>
__global__ void __launch_bounds__(TC_THREADS, 1) flow_tc_apply_kernel(TcParams P, TcIO io) {
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  TcShared* sh;
  tc_prologue(P, tc_smem, sh);
  const int warp = threadIdx.x >> 5;
  const int64_t ntiles = (io.n + 127) / 128;
  const uint32_t tmem = sh->tmem_base;
  if (warp < TC_NG * 4) {
    const int g = warp >> 2, t = threadIdx.x & 127;
    const uint32_t tg = tmem + g * TC_COLS + ((uint32_t)((warp & 3) * 32) << 16);
    const TcGroup G{tc_smem_u32(&sh->bar_in[g]), tc_smem_u32(&sh->bar_out[g]), tc_smem_u32(tc_smem),
                    __shfl_sync(0xffffffffu, tmem, 0) + g * TC_COLS, g, (warp & 3) == 0};
    uint32_t ph_out = 0;
    const int64_t stride = (int64_t)gridDim.x * TC_NG;
    for (int64_t tile = (int64_t)blockIdx.x * TC_NG + g; tile < ntiles; tile += stride) {
      const int64_t row = tile * 128 + t;
      const bool valid = row < io.n;
      float h[TC_DP];
      float ss_in = 0.f;
      if (P.D == TC_DP) {  // a row is 64 contiguous, 64-byte aligned bytes
        const float4* p4 = reinterpret_cast<const float4*>(io.in + row * TC_DP);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 v = valid ? __ldg(p4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
          h[4 * q] = v.x, h[4 * q + 1] = v.y, h[4 * q + 2] = v.z, h[4 * q + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int d = 0; d < TC_DP; ++d) h[d] = (valid && d < P.D) ? __ldg(io.in + row * P.D + d) : 0.f;
      }
#pragma unroll
      for (int d = 0; d < TC_DP; ++d) ss_in = fmaf(h[d], h[d], ss_in);
      const float ld = tc_run_row<NARROW, ACT>(P, tc_smem, tg, G, ph_out, h) + P.const_logdet;
      float ss_out = 0.f;
#pragma unroll
      for (int d = 0; d < TC_DP; ++d) ss_out = d < P.D ? fmaf(h[d], h[d], ss_out) : ss_out;
      if (valid && io.out) {
        if (P.D == TC_DP) {
          float4* o4 = reinterpret_cast<float4*>(io.out + row * TC_DP);
#pragma unroll
          for (int q = 0; q < 4; ++q) o4[q] = make_float4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
        } else {
#pragma unroll
          for (int d = 0; d < TC_DP; ++d)
            if (d < P.D) io.out[row * P.D + d] = h[d];
        }
      }
      if (valid) {
        if (io.out_logj) io.out_logj[row] = ld;
        if (io.out_lp) {
          const float c = io.base_log_z, hv = 0.5f * io.base_inv_var;
          io.out_lp[row] = (io.lp_mode == 1) ? (-hv * ss_in - c) - ld : (-hv * ss_out - c) + ld;
        }
      }
    }
  } else if (!TC_SELF) {
    // issuer warp g: all 32 lanes converged, one elected lane fires each MMA
    const int g = __shfl_sync(0xffffffffu, warp - TC_NG * 4, 0);
    tc_issuer_n<NARROW ? 2 : 4>(P, tc_smem_u32(tc_smem), __shfl_sync(0xffffffffu, tmem, 0) + g * TC_COLS,
              tc_smem_u32(&sh->bar_in[g]), tc_smem_u32(&sh->bar_out[g]), tc_my_tiles(ntiles, g));
    __syncwarp();
  }
  tc_epilogue_end(sh);
}

template <bool NARROW, int ACT>
__global__ void __launch_bounds__(TC_THREADS, 1) flow_tc_populate_kernel(TcParams P, PopulateArgs A) {
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  TcShared* sh = reinterpret_cast<TcShared*>(tc_smem + tc_image_pad(P.image_bytes));
  if (threadIdx.x < 4 * TC_DP) {
    const int which = threadIdx.x / TC_DP, d = threadIdx.x % TC_DP;
    const double* src = which == 0 ? A.scale : which == 1 ? A.shift : which == 2 ? A.lo : A.hi;
    sh->cst[which][d] = d < P.D ? src[d] : 0.0;
  }
  if (threadIdx.x == 4 * TC_DP) sh->log_const = populate_log_const(A, P.D);
  tc_prologue(P, tc_smem, sh);
  const int warp = threadIdx.x >> 5;
  const int64_t ntiles = (A.n + 127) / 128;
  const uint32_t tmem = sh->tmem_base;
  if (warp < TC_NG * 4) {
    const int g = warp >> 2, t = threadIdx.x & 127;
    const uint32_t tg = tmem + g * TC_COLS + ((uint32_t)((warp & 3) * 32) << 16);
    const TcGroup G{tc_smem_u32(&sh->bar_in[g]), tc_smem_u32(&sh->bar_out[g]), tc_smem_u32(tc_smem),
                    __shfl_sync(0xffffffffu, tmem, 0) + g * TC_COLS, g, (warp & 3) == 0};
    uint32_t ph_out = 0;
    double vmax = -INFINITY, vcount = 0.0;
    const double *c_scale = sh->cst[0], *c_shift = sh->cst[1], *c_lo = sh->cst[2], *c_hi = sh->cst[3];
    const double log_const = sh->log_const;
    const int64_t stride = (int64_t)gridDim.x * TC_NG;
    for (int64_t tile = (int64_t)blockIdx.x * TC_NG + g; tile < ntiles; tile += stride) {
      const int64_t row = tile * 128 + t;
      float h[TC_DP];
      float ss = 0.f;
#pragma unroll
      for (int d0 = 0; d0 < TC_DP; d0 += 4) {
        // unconditional: the four Philox chains interleave (unused features are masked below)
        float v[4];
        {
          const Philox4 r = philox4x32_10(A.seed, A.row_offset + row, d0 / 4, 0);
          box_muller(r.x, r.y, v[0], v[1]);
          box_muller(r.z, r.w, v[2], v[3]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool use = d0 + j < P.D;
          ss = use ? fmaf(v[j], v[j], ss) : ss;
          h[d0 + j] = use ? v[j] * A.sqrt_t : 0.f;
          if (A.z && use && row < A.n) A.z[row * P.D + d0 + j] = h[d0 + j];
        }
      }
      const float rad = sqrtf(ss) * A.sqrt_t;
      const bool alive = !(A.r_max > 0.f) || (rad <= A.r_max);
      const float logj = tc_run_row<NARROW, ACT>(P, tc_smem, tg, G, ph_out, h) + P.const_logdet;
      const float base_lp = -0.5f * ss - 0.5f * P.D * TC_LOG_2PI;
      populate_row<TC_DP>(A, P.D, [&](int d) { return h[d]; }, row, alive, base_lp, logj, vmax,
                          vcount, c_scale, c_shift, c_lo, c_hi, log_const);
    }
    populate_publish(A, vmax, vcount);
  } else if (!TC_SELF) {
    // issuer warp g: all 32 lanes converged, one elected lane fires each MMA
    const int g = __shfl_sync(0xffffffffu, warp - TC_NG * 4, 0);
    tc_issuer_n<NARROW ? 2 : 4>(P, tc_smem_u32(tc_smem), __shfl_sync(0xffffffffu, tmem, 0) + g * TC_COLS,
              tc_smem_u32(&sh->bar_in[g]), tc_smem_u32(&sh->bar_out[g]), tc_my_tiles(ntiles, g));
    __syncwarp();
  }
  tc_epilogue_end(sh);
}

inline int tc_prep(const void* kernel, size_t smem) {
  // opt in to > 48 KB dynamic shared memory; remembers the largest size set per kernel
  // for the current device (the attribute is per device)
  static thread_local const void* done[32];
  static thread_local int done_dev[32];
  static thread_local size_t done_smem[32];
  static thread_local int nd = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 1;
  int slot = -1;
  for (int i = 0; i < nd; ++i)
    if (done[i] == kernel && done_dev[i] == dev) slot = i;
  if (slot >= 0 && done_smem[slot] >= smem) return 0;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
      cudaSuccess)
    return 1;
  if (slot < 0 && nd < 32) slot = nd++;
  if (slot >= 0) {
    done[slot] = kernel;
    done_dev[slot] = dev;
    done_smem[slot] = smem;
  }
  return 0;
}

inline TcParams tc_params(const TcProgram& t, float const_logdet) {
  TcParams P;
  P.image = t.d_image;
  P.image_bytes = t.image_bytes;
  P.L = t.L;
  P.D = t.D;
  for (int i = 0; i < TC_MAXL; ++i) {
    P.d_id[i] = t.d_id[i];
    P.d_tr[i] = t.d_tr[i];
  }
  P.additive = t.additive;
  P.inverse = t.inverse;
  P.narrow = t.narrow;
  P.const_logdet = const_logdet;
  return P;
}

// run LAUNCH(NARROW, ACT) for the run-time (narrow, act): one kernel instantiation per combination
#define NB200_TC_DISPATCH(LAUNCH, narrow, act)                \
  if (narrow) {                                                \
    if ((act) == ACT_TANH) LAUNCH(true, ACT_TANH)              \
    else if ((act) == ACT_SILU) LAUNCH(true, ACT_SILU)         \
    else LAUNCH(true, ACT_RELU)                                \
  } else {                                                     \
    if ((act) == ACT_TANH) LAUNCH(false, ACT_TANH)             \
    else if ((act) == ACT_SILU) LAUNCH(false, ACT_SILU)        \
    else LAUNCH(false, ACT_RELU)                               \
  }

inline int tc_grid(int64_t n, int num_sms) {
  const int64_t ntiles = (n + 127) / 128;
  const int64_t want = (ntiles + TC_NG - 1) / TC_NG;
  return (int)(want < num_sms ? want : num_sms);
}

inline int tc_launch_apply(TcProgram& t, const TcIO& io, int num_sms, cudaStream_t st) {
  const size_t smem = tc_smem_bytes(t.image_bytes);
  const int64_t n = io.n;
#define NB200_TC_APPLY(NARROW, ACT)                                                                            \
  {                                                                                                              \
    if (tc_prep((const void*)flow_tc_apply_kernel<NARROW, ACT>, smem)) return 1;                                \
    flow_tc_apply_kernel<NARROW, ACT><<<tc_grid(n, num_sms), TC_THREADS, smem, st>>>(tc_params(t, t.const_logdet), io); \
  }
  NB200_TC_DISPATCH(NB200_TC_APPLY, t.narrow, t.act)
#undef NB200_TC_APPLY
  return cudaGetLastError() != cudaSuccess;
}

inline int tc_launch_populate(TcProgram& t, const PopulateArgs& A, int num_sms, cudaStream_t st) {
  const size_t smem = tc_smem_bytes(t.image_bytes);
#define NB200_TC_POPULATE(NARROW, ACT)                                                                            \
  {                                                                                                                 \
    if (tc_prep((const void*)flow_tc_populate_kernel<NARROW, ACT>, smem)) return 1;                                \
    flow_tc_populate_kernel<NARROW, ACT><<<tc_grid(A.n, num_sms), TC_THREADS, smem, st>>>(tc_params(t, t.const_logdet), A); \
  }
  NB200_TC_DISPATCH(NB200_TC_POPULATE, t.narrow, t.act)
#undef NB200_TC_POPULATE
  return cudaGetLastError() != cudaSuccess;
}

}  // namespace nb200
