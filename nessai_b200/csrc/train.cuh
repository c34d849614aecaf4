// Fused training step of FlowModel._train (/root/reference/src/nessai/flowmodel/base.py:365-452)
// for RealNVP flows: train-mode forward (batch-statistics BatchNorm, uncached LU), the
// hand-derived backward pass, a deterministic gradient reduction, global-norm clipping and
// Adam/AdamW -- the arithmetic the reference leaves to torch autograd + glasflow.nflows.
// The backward pass is the one restated (and pinned against autograd) in oracle/train_numpy.py.
//
// Decomposition.  Rows are independent except through BatchNorm's batch statistics, so a step is a
// chain of small kernels whose boundaries are exactly those grid-wide reductions:
//   FWD(0..L-1)  -> LOSS -> BWD(L-1..0) -> REDUCE -> ADAM            (2L + 3 launches)
// A batch is cut into tiles of 16 rows; a CTA (512 threads) owns tiles blockIdx.x, +gridDim.x, ...
// and keeps every per-row quantity of a tile in shared memory, FEATURE-MAJOR ([feature][16 rows]),
// so each thread's 8-row register tile is two 16-byte shared loads per k and weights are read
// conflict-free.  Saved activations go to an L2-resident workspace in the same layout.
// Parameter gradients are per-CTA partial sums (no atomics -> bit-reproducible), reduced by REDUCE.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace nb200 {

#ifndef NB200_TR_R
#define NB200_TR_R 16
#endif
constexpr int TR_R = NB200_TR_R;  // rows per tile
constexpr int TR_THREADS = 512;  // threads per CTA of the row kernels (latency-bound: 4 warps / SMSP)
constexpr int TR_RG = TR_THREADS / 64;   // row groups of the tile GEMM
constexpr int TR_RT = TR_R / TR_RG;      // rows per thread
constexpr int TR_CHUNK = 64;             // output columns per staged weight unit
constexpr int TR_MAXL = 16;      // coupling layers
constexpr int TR_MAXBUF = 12;    // conditioner buffers per layer
constexpr int TR_MAXLIN = 12;    // linears per conditioner
constexpr int TR_MAXD = 64;      // features
constexpr int TR_MAXG = 148;     // CTAs per launch (one per SM)

// ---- plan: ints only; packed by nessai_b200/train_plan.py in exactly this order ----------------
struct TrLinear {
  int w_off, b_off, n_in, n_out, in_buf, out_buf, res_buf, pre_act;
};
struct TrLayer {
  int perm_off, lu_bias, lu_lower, lu_upper, lu_diag;  // -1 when absent (perm_off: into itab)
  int bn_uw, bn_bias, bn_rm, bn_rv;                    // bn_rm / bn_rv: offsets into theta_b
  int id_off, tr_off, d_id, d_tr;                      // index lists in itab
  int n_lin, n_buf, rec_floats;                        // saved floats per row: h2 | bufs 1.. | y
  int ws_off;                                          // feature offset of this layer's record
  int lu_part_off;                                     // dense LU dW (D*D) inside a partial vector
  int pad0, pad1;
  int buf_dim[TR_MAXBUF];
  int buf_off[TR_MAXBUF];  // feature offset inside the tile value area (buf 0 = identity half)
  TrLinear lin[TR_MAXLIN];
};
struct TrPlan {
  int D, L, act, additive;  // additive: 0 affine coupling, 1 additive coupling, 2 masked affine autoregressive (MAF)
  int n_params, n_part, rec_total, max_dim;
  int vals_floats, wmax, n_itab, n_reduce;
  int num_bins;      // > 0: rational-quadratic spline coupling (NSF) with this many bins
  int tail_bound;    // float bits of the spline's linear-tail bound
  int hidden;        // conditioner width (nflows divides the unnormalised widths / heights by sqrt of it)
  int max_in;        // max(D, widest linear INPUT): rows of the staged-operand tiles
  int base_var;      // float bits of the base distribution's variance (N(0, var I); 1: StandardNormal)
  int pad0, pad1, pad2;
  TrLayer layer[TR_MAXL];
};
constexpr int TR_LAYER_INTS = 20 + 2 * TR_MAXBUF + 8 * TR_MAXLIN;
constexpr int TR_PLAN_INTS = 20 + TR_MAXL * TR_LAYER_INTS;
static_assert(sizeof(TrLayer) == 4 * TR_LAYER_INTS, "TrLayer layout");
static_assert(sizeof(TrPlan) == 4 * TR_PLAN_INTS, "TrPlan layout");

struct TrBuffers {
  const int* itab;
  const int* reduce_idx;
  float* theta_p;
  float* theta_b;
  float* ws;          // [tile][rec_total][16]
  float* dout[2];     // [tile][D][16]
  float* ldrow;       // [tile][16]
  float* crow;        // [tile][16] loss weight of each row (0 for padding rows)
  float* stat_part;   // [L][G][2][D] per-CTA (mean, M2)
  float* stat_n;      // [G] rows per CTA
  float* stats;       // [L][2][D] batch mean, unbiased variance
  float* s_part[2];   // [G][2][D] BatchNorm backward sums
  float* wsum_part;   // [G]
  float* loss_part;   // [G]
  float* part;        // [G][part_stride] gradient partials (part_stride = n_part rounded up to 4 floats)
  int part_stride;
  float* grad;        // [n_params rounded up to 4]
  float* gn_part;     // [n_reduce_blocks]
  int n_reduce_blocks;
  const float* pmask; // [n_params] or NULL: MADE masks (1 elsewhere); masked weights stay exactly 0
  int G;              // CTAs of the row kernels
  // NB200_TR_TRACE (debugging aid): [2 * trace_cap + 1] = (tag, globaltimer ns) marks of CTA 0, then the mark count
  long long* trace = nullptr;
  int trace_cap = 0;
};

struct TrBatch {
  const float* x;       // [n_rows][D] row-major
  const int64_t* perm;  // row order of the epoch (NULL: identity)
  const float* w;       // per-row weights (NULL: unweighted)
  int64_t i0;           // first row of the batch in `perm`
  int B;                // rows in the batch
  int n_tiles;
};

struct TrOptim {
  int kind;  // 0 AdamW (decoupled decay), 1 Adam (L2 decay), 2 SGD, -1 gradient only
  float lr, beta1, beta2, eps, weight_decay, clip;
  float bc1, bc2;  // 1 - beta^t
};

#ifndef NB200_SIMT_SHIM
__device__ __forceinline__ void tr_mark(const TrBuffers& Bf, int tag) {
  if (Bf.trace && blockIdx.x == 0 && threadIdx.x == 0) {
    const long long slot = Bf.trace[2 * Bf.trace_cap];
    if (slot < Bf.trace_cap) {
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      Bf.trace[2 * slot] = tag;
      Bf.trace[2 * slot + 1] = t;
      Bf.trace[2 * Bf.trace_cap] = slot + 1;
    }
  }
}
#else
__device__ __forceinline__ void tr_mark(const TrBuffers&, int) {}
#endif
// -DNB200_TR_FINE: marks inside the phases as well (tags 1000+)
#ifdef NB200_TR_FINE
#define TR_T(tag) tr_mark(Bf, tag)
#else
#define TR_T(tag)
#endif

constexpr float TR_LU_EPS = 1e-3f, TR_BN_EPS = 1e-5f, TR_BN_MOM = 0.1f;
constexpr float TR_HALF_LOG_2PI = 0.91893853320467274178f;

// ---------------------------------------------------------------------------- elementwise
__device__ __forceinline__ float tr_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float tr_softplus(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float tr_act(int act, float z) {
  if (act == 0) return fmaxf(z, 0.f);
  if (act == 1) return tanhf(z);
  return z * tr_sigmoid(z);
}
__device__ __forceinline__ float tr_dact(int act, float z) {
  if (act == 0) return z > 0.f ? 1.f : 0.f;
  if (act == 1) {
    const float t = tanhf(z);
    return 1.f - t * t;
  }
  const float s = tr_sigmoid(z);
  return s * (1.f + z * (1.f - s));
}

// Parameters are rewritten by the optimiser phase of the same (persistent) kernel: never read them
// through the non-coherent path.
__device__ __forceinline__ float tr_ldp(const float* p) { return *p; }

// ---------------------------------------------------------------------------- tile GEMMs
// out[c][r] (=|+=) bias[c] + sum_k W[k*ldw + c] * A[k][r] (+ res[c][r]);  c < N, r < 16.
// Callers synchronise.
// Thread (c = tid & 63, rows 2 (tid >> 6) .. + 2).
__device__ __forceinline__ void tr_gemm(float* __restrict__ out, const float* __restrict__ A,
                                        const float* __restrict__ W, int ldw,
                                        const float* __restrict__ bias,
                                        const float* __restrict__ res, int N, int K, bool accum) {
  static_assert(TR_THREADS == 512 && (TR_RT == 2 || TR_RT % 4 == 0), "tile GEMM register tiles");
  // (a 4-row x 256-thread mapping halves the shared-memory wavefronts but measured slower: with one
  // tile per CTA the GEMM is latency-bound and wants every warp)
  const int rg = threadIdx.x >> 6, cl = threadIdx.x & 63;
  for (int c0 = 0; c0 < N; c0 += 64) {
    const int c = c0 + cl;
    if (c < N) {
      const float b = bias ? bias[c] : 0.f;
      float acc[TR_RT];
#pragma unroll
      for (int i = 0; i < TR_RT; ++i) acc[i] = b;
      const float* Ar = A + rg * TR_RT;
      const float* w = W + c;
#pragma unroll 4
      for (int k = 0; k < K; ++k) {
        const float wk = w[k * ldw];
        if constexpr (TR_RT == 2) {
          const float2 a = *reinterpret_cast<const float2*>(Ar + k * TR_R);
          acc[0] = fmaf(wk, a.x, acc[0]);
          acc[1] = fmaf(wk, a.y, acc[1]);
        } else {
#pragma unroll
          for (int q = 0; q < TR_RT / 4; ++q) {
            const float4 a = *reinterpret_cast<const float4*>(Ar + k * TR_R + 4 * q);
            acc[4 * q] = fmaf(wk, a.x, acc[4 * q]);
            acc[4 * q + 1] = fmaf(wk, a.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(wk, a.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(wk, a.w, acc[4 * q + 3]);
          }
        }
      }
      float* o = out + c * TR_R + rg * TR_RT;
      if (res) {
        const float* rr = res + c * TR_R + rg * TR_RT;
#pragma unroll
        for (int i = 0; i < TR_RT; ++i) acc[i] += rr[i];
      }
      if (accum) {
#pragma unroll
        for (int i = 0; i < TR_RT; ++i) acc[i] += o[i];
      }
#pragma unroll
      for (int i = 0; i < TR_RT; ++i) o[i] = acc[i];
    }
  }
}

// Weight gradient of one linear over the tile: g[c*K + k] (=|+=) sum_r delta[c][r] * A[k][r].
__device__ __forceinline__ void tr_wgrad(float* __restrict__ g, const float* __restrict__ delta,
                                         const float* __restrict__ A, int N, int K, bool first) {
  const int tid = threadIdx.x;
  if (K <= TR_THREADS && (TR_THREADS % K) == 0) {
    const int k = tid % K, cstep = TR_THREADS / K;
    float a[TR_R];
    const float4* A4 = reinterpret_cast<const float4*>(A + k * TR_R);
#pragma unroll
    for (int q = 0; q < TR_R / 4; ++q) {
      const float4 v = A4[q];
      a[4 * q] = v.x, a[4 * q + 1] = v.y, a[4 * q + 2] = v.z, a[4 * q + 3] = v.w;
    }
    for (int c = tid / K; c < N; c += cstep) {
      const float4* d4 = reinterpret_cast<const float4*>(delta + c * TR_R);
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < TR_R / 4; ++q) {
        const float4 v = d4[q];
        s = fmaf(v.x, a[4 * q], s);
        s = fmaf(v.y, a[4 * q + 1], s);
        s = fmaf(v.z, a[4 * q + 2], s);
        s = fmaf(v.w, a[4 * q + 3], s);
      }
      float* p = g + c * K + k;
      *p = first ? s : *p + s;
    }
  } else {
    for (int e = tid; e < N * K; e += TR_THREADS) {
      const int c = e / K, k = e - c * K;
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < TR_R; ++r) s = fmaf(delta[c * TR_R + r], A[k * TR_R + r], s);
      g[e] = first ? s : g[e] + s;
    }
  }
}

// Bias gradient: g[c] (=|+=) sum_r delta[c][r]
__device__ __forceinline__ void tr_bgrad(float* __restrict__ g, const float* __restrict__ delta,
                                         int N, bool first) {
  for (int c = threadIdx.x; c < N; c += TR_THREADS) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < TR_R; ++r) s += delta[c * TR_R + r];
    g[c] = first ? s : g[c] + s;
  }
}

// ---------------------------------------------------------------------------- async copies
// Global -> shared staging uses cp.async (LDGSTS): a thread queues all its copies without waiting,
// so a tile / a weight matrix costs ONE memory latency instead of one per loop iteration (with a
// single warp per SM sub-partition there is no other latency hiding).
#ifdef NB200_SIMT_SHIM
// host flavour for the CPU SIMT shim of tests/_hostcheck: the copies complete at once
__device__ __forceinline__ void tr_cp4(float* dst, const float* src) { *dst = *src; }
__device__ __forceinline__ void tr_cp16(float* dst, const float* src) {
  dst[0] = src[0], dst[1] = src[1], dst[2] = src[2], dst[3] = src[3];
}
__device__ __forceinline__ void tr_cp_commit() {}
template <int N>
__device__ __forceinline__ void tr_cp_wait() {}
#else
__device__ __forceinline__ uint32_t tr_s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tr_cp4(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(tr_s32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void tr_cp16(float* dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tr_s32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void tr_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tr_cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
#endif

// Store a [n][16] shared tile to global memory (both 16-byte aligned).
__device__ __forceinline__ void tr_copy(float* __restrict__ dst, const float* __restrict__ src,
                                        int n_feat) {
  float4* d = reinterpret_cast<float4*>(dst);
  const float4* s = reinterpret_cast<const float4*>(src);
  for (int i = threadIdx.x; i < n_feat * (TR_R / 4); i += TR_THREADS) d[i] = s[i];
}
// Queue the load of a [n][16] global tile into shared memory (caller commits and waits).
__device__ __forceinline__ void tr_copy_async(float* dst, const float* src, int n_feat) {
  for (int i = threadIdx.x; i < n_feat * (TR_R / 4); i += TR_THREADS) tr_cp16(dst + 4 * i, src + 4 * i);
}
// Queue n floats (no alignment assumed).
__device__ __forceinline__ void tr_copy_async4(float* dst, const float* src, int n) {
  for (int i = threadIdx.x; i < n; i += TR_THREADS) tr_cp4(dst + i, src + i);
}

// Queue nn.Linear weights (row-major [n_out][n_in]) into shared memory.
// transposed: W[k * (n_out | 1) + c]  (forward: thread index = output column c)
// natural:    W[c * n_in + k]          (backward: thread index = input column k)
__device__ __forceinline__ void tr_load_w_async(float* sW, const float* W, int n_out, int n_in,
                                                bool transposed) {
  const int n = n_out * n_in;
  if (transposed) {
    const int ld = n_out | 1;
    int c = threadIdx.x / n_in, k = threadIdx.x - c * n_in;
    const int dc = TR_THREADS / n_in, dk = TR_THREADS - dc * n_in;
    for (int i = threadIdx.x; i < n; i += TR_THREADS) {
      tr_cp4(sW + k * ld + c, W + i);
      c += dc, k += dk;
      if (k >= n_in) k -= n_in, ++c;
    }
  } else if ((((uintptr_t)W) & 15) == 0 && (n & 3) == 0) {
    for (int i = threadIdx.x; i < n / 4; i += TR_THREADS) tr_cp16(sW + 4 * i, W + 4 * i);
  } else {
    for (int i = threadIdx.x; i < n; i += TR_THREADS) tr_cp4(sW + i, W + i);
  }
}

// Dense W = Lo Up of nflows' LULinear (unit lower x upper with softplus diagonal) into shared
// memory, transposed as above.  The two triangles are first expanded into dense D x D matrices
// Lo = stage[0 .. D*D), Up = stage[D*D .. 2*D*D) (zeros included), which stay valid for the
// caller (the backward phase applies the chain rule with them): every global load independent,
// then a branch-free D x D x D product.
__device__ __forceinline__ void tr_lu_dense(float* __restrict__ sW, float* __restrict__ stage,
                                            const float* __restrict__ theta, const TrLayer& ly,
                                            int D, bool transposed) {
  float* Lo = stage;
  float* Up = stage + D * D;
  for (int e = threadIdx.x; e < D * D; e += TR_THREADS) {
    const int i = e / D, j = e - i * D;
    float lo = 0.f, up = 0.f;
    if (i == j) {
      lo = 1.f;
      up = tr_softplus(tr_ldp(theta + ly.lu_diag + i)) + TR_LU_EPS;
    } else if (j < i) {
      lo = tr_ldp(theta + ly.lu_lower + i * (i - 1) / 2 + j);
    } else {
      up = tr_ldp(theta + ly.lu_upper + i * D - i * (i + 1) / 2 + (j - i - 1));
    }
    Lo[e] = lo;
    Up[e] = up;
  }
  for (int j = threadIdx.x; j < D; j += TR_THREADS) sW[D * (D | 1) + j] = tr_ldp(theta + ly.lu_bias + j);
  __syncthreads();
  const int ld = transposed ? (D | 1) : D;
  for (int e = threadIdx.x; e < D * D; e += TR_THREADS) {
    const int i = e / D, j = e - i * D;
    const int mmax = i < j ? i : j;
    float s = 0.f;
    for (int m = 0; m <= mmax; ++m) s = fmaf(Lo[i * D + m], Up[m * D + j], s);
    if (transposed) sW[j * ld + i] = s;
    else sW[i * ld + j] = s;
  }
  __syncthreads();
}

// Chain rule of the LU parametrisation for this CTA's dense partial dW (D x D, row-major) into the
// partial gradients of the lower / upper / unconstrained-diagonal entries (linear in dW, so it
// commutes with the sum over CTAs); Lo / Up as tr_lu_dense left them.  `once`: this CTA also adds
// the row-constant log|det| term of the loss, d(-sum_i log Up_ii) (the row weights sum to 1).
__device__ __forceinline__ void tr_lu_chain(float* __restrict__ part, const float* __restrict__ dW,
                                            const float* __restrict__ Lo, const float* __restrict__ Up,
                                            const float* __restrict__ theta, const TrLayer& ly, int D,
                                            bool once) {
  for (int e = threadIdx.x; e < D * D; e += TR_THREADS) {
    const int i = e / D, j = e - i * D;
    if (j < i) {  // d lower[i][j] = sum_k dW[i][k] Up[j][k]
      float s = 0.f;
      for (int k = j; k < D; ++k) s = fmaf(dW[i * D + k], Up[j * D + k], s);
      part[ly.lu_lower + i * (i - 1) / 2 + j] = s;
    } else {  // d upper[i][j] = sum_k Lo[k][i] dW[k][j]
      float s = 0.f;
      for (int k = i; k < D; ++k) s = fmaf(Lo[k * D + i], dW[k * D + j], s);
      if (i == j) {
        if (once) s -= 1.f / Up[e];
        part[ly.lu_diag + i] = s * tr_sigmoid(tr_ldp(theta + ly.lu_diag + i));
      } else {
        part[ly.lu_upper + i * D - i * (i + 1) / 2 + (j - i - 1)] = s;
      }
    }
  }
}

// ---------------------------------------------------------------------------- shared memory map
struct TrSmem {
  float* W;     // staged weights of one linear, double buffered: W + (j & 1) * wbuf
  int wbuf;
  float* Wlu;   // dense LU matrix [D][D|1]
  float* H1;    // [D][16] layer input (after the permutation)
  float* H2;    // [D][16] after the LU linear
  float* Y;     // [D][16] coupling output (before BatchNorm)
  float* X;     // [D][16] scratch (x-hat / dy / dh2 ...)
  float* Pf;    // [D][16] prefetched tile (backward: previous layer's output)
  float* V;     // [vals][16] conditioner buffers
  float* Gv;    // [vals][16] their gradients (backward only)
  float* A;     // [max_in][16] staged GEMM operand
  float* A2;    // [max_in][16] second operand stage (forward) / input-gradient accumulator over output chunks (backward)
  float* bn;    // [4][D]: mean, rstd, w, beta  (+ [2][D] S1,S2 in backward)
  float* c;     // [16] row weights
  float* ld;    // [16]
  float* red;   // [3 * TR_THREADS] block reductions
  float* stage; // [TR_MAXG][2][D] per-CTA partials staged for the cross-CTA reductions
  int* itab;    // permutations / mask index lists
};
__host__ __device__ inline int tr_align4(int n) { return (n + 3) & ~3; }
__host__ __device__ inline size_t tr_smem_floats(int D, int vals, int max_in, int wmax, int n_itab, bool backward) {
  size_t n = 0;
  n += 2 * (tr_align4(wmax) + TR_CHUNK);
  n += tr_align4(D * (D | 1) + D);
  n += 5 * (size_t)D * TR_R;
  n += (size_t)vals * TR_R * (backward ? 2 : 1);
  n += (size_t)max_in * TR_R * 2;
  n += tr_align4(6 * D);
  n += 2 * TR_R + 3 * TR_THREADS;
  n += (size_t)TR_MAXG * (2 * D + 1);
  n += tr_align4(n_itab);
  return n;
}
__device__ __forceinline__ TrSmem tr_carve(float* s, const TrPlan& P, bool backward) {
  TrSmem m;
  const int D = P.D;
  m.W = s, m.wbuf = tr_align4(P.wmax) + TR_CHUNK, s += 2 * m.wbuf;
  m.Wlu = s, s += tr_align4(D * (D | 1) + D);
  m.H1 = s, s += D * TR_R;
  m.H2 = s, s += D * TR_R;
  m.Y = s, s += D * TR_R;
  m.X = s, s += D * TR_R;
  m.Pf = s, s += D * TR_R;
  m.V = s, s += P.vals_floats * TR_R;
  m.Gv = s;
  if (backward) s += P.vals_floats * TR_R;
  m.A = s, s += P.max_in * TR_R;
  m.A2 = s, s += P.max_in * TR_R;
  m.bn = s, s += tr_align4(6 * D);
  m.c = s, s += TR_R;
  m.ld = s, s += TR_R;
  m.red = s, s += 3 * TR_THREADS;
  m.stage = s, s += TR_MAXG * (2 * D + 1);
  m.itab = reinterpret_cast<int*>(s);
  return m;
}

// Index tables into shared memory (first thing every row kernel does).
__device__ __forceinline__ void tr_stage_itab(const TrBuffers& Bf, const TrPlan& P, const TrSmem& S) {
  tr_copy_async4(reinterpret_cast<float*>(S.itab), reinterpret_cast<const float*>(Bf.itab), P.n_itab);
  tr_cp_commit();
  tr_cp_wait<0>();
  __syncthreads();
}

__device__ __forceinline__ float tr_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the block (fixed order: lanes by shuffle, then the warps' sums by every thread).
template <int NT = TR_THREADS>
__device__ __forceinline__ float tr_block_sum(float v, float* red) {
  v = tr_warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) r += red[w];
  __syncthreads();
  return r;
}


// Pool the per-CTA (n, mean, M2) partials of BatchNorm layer `l` into bn[0][d] = mean,
// bn[1][d] = 1/sqrt(var + eps), and the unbiased variance through var_out:
//   mean = sum_g n_g mean_g / N,   M2 = sum_g (M2_g + n_g (mean_g - mean)^2)
// (the exact pooled form of Chan et al.'s update, without its sequential chain: one warp per
// feature, lanes over the CTAs, every load independent).
constexpr int TR_STAT_PER_LANE = (TR_MAXG + 31) / 32;
__device__ __forceinline__ void tr_reduce_stats(const TrBuffers& Bf, int l, int D, int B, float* bn,
                                                float* var_out, float* stage, float* red) {
  (void)stage, (void)red;
  const int G = Bf.G;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* part = Bf.stat_part + (size_t)l * G * 2 * D;
  float nb[TR_STAT_PER_LANE];
#pragma unroll
  for (int q = 0; q < TR_STAT_PER_LANE; ++q) {
    const int g = lane + 32 * q;
    nb[q] = g < G ? Bf.stat_n[g] : 0.f;
  }
  for (int d = warp; d < D; d += TR_THREADS / 32) {
    float mg[TR_STAT_PER_LANE], qg[TR_STAT_PER_LANE];
    float sn = 0.f, sm = 0.f;
#pragma unroll
    for (int q = 0; q < TR_STAT_PER_LANE; ++q) {
      const int g = lane + 32 * q;
      // (every CTA of the grid writes its partial, rows or not: the loads need not wait for nb)
      const bool ok = g < G;
      mg[q] = ok ? part[(size_t)g * 2 * D + d] : 0.f;
      qg[q] = ok ? part[(size_t)g * 2 * D + D + d] : 0.f;
      sn += nb[q];
      sm = fmaf(nb[q], mg[q], sm);
    }
    sn = tr_warp_sum(sn);
    sm = tr_warp_sum(sm);
    const float mean = sm / sn;
    float m2 = 0.f;
#pragma unroll
    for (int q = 0; q < TR_STAT_PER_LANE; ++q) {
      const float delta = mg[q] - mean;
      m2 += qg[q] + nb[q] * delta * delta;
    }
    m2 = tr_warp_sum(m2);
    if (lane == 0) {
      const float var = m2 / (float)(B - 1);
      bn[d] = mean;
      bn[D + d] = rsqrtf(var + TR_BN_EPS);
      if (var_out) var_out[d] = var;
    }
  }
  __syncthreads();
}

// out[c] = sum_g src[g * ncol + c]: one warp per column, lanes over the G partials (every load
// independent, fixed order), one barrier.
__device__ __forceinline__ void tr_colsum(const float* __restrict__ src, int G, int ncol, float* out,
                                          float* red) {
  (void)red;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = warp; c < ncol; c += TR_THREADS / 32) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < TR_STAT_PER_LANE; ++q) {
      const int g = lane + 32 * q;
      s += g < G ? src[(size_t)g * ncol + c] : 0.f;
    }
    s = tr_warp_sum(s);
    if (lane == 0) out[c] = s;
  }
  __syncthreads();
}

// BatchNorm layer l for the training pass: bn[0..3][D] = mean, rstd, w, beta.
// The affine parameters do not depend on the batch (they can be loaded while a grid barrier is pending):
__device__ __forceinline__ void tr_bn_affine(const TrBuffers& Bf, const TrLayer& ly, int D, float* bn) {
  for (int d = threadIdx.x; d < D; d += TR_THREADS) {
    bn[2 * D + d] = tr_softplus(tr_ldp(Bf.theta_p + ly.bn_uw + d)) + TR_BN_EPS;
    bn[3 * D + d] = tr_ldp(Bf.theta_p + ly.bn_bias + d);
  }
}
// first use of the layer's batch statistics: pool the per-CTA partials (scratch[D] receives the
// variance); the last CTA publishes them and EMA-updates the running buffers (nflows BatchNorm)
__device__ __forceinline__ void tr_bn_stats_first(const TrBuffers& Bf, const TrLayer& ly, int l, int D, int B,
                                                  float* bn, float* scratch) {
  tr_reduce_stats(Bf, l, D, B, bn, scratch, nullptr, nullptr);
  if (blockIdx.x == gridDim.x - 1)
    for (int d = threadIdx.x; d < D; d += TR_THREADS) {
      Bf.stats[(l * 2) * D + d] = bn[d];
      Bf.stats[(l * 2 + 1) * D + d] = scratch[d];
      float* rm = Bf.theta_b + ly.bn_rm;
      float* rv = Bf.theta_b + ly.bn_rv;
      rm[d] = (1.f - TR_BN_MOM) * rm[d] + TR_BN_MOM * bn[d];
      rv[d] = (1.f - TR_BN_MOM) * rv[d] + TR_BN_MOM * scratch[d];
    }
}
// later uses: the published statistics
__device__ __forceinline__ void tr_bn_stats_published(const TrBuffers& Bf, int l, int D, float* bn) {
  for (int d = threadIdx.x; d < D; d += TR_THREADS) {
    bn[d] = Bf.stats[(l * 2) * D + d];
    bn[D + d] = rsqrtf(Bf.stats[(l * 2 + 1) * D + d] + TR_BN_EPS);
  }
}

// Eval-mode BatchNorm constants (running statistics).
__device__ __forceinline__ void tr_bn_setup_eval(const TrBuffers& Bf, const TrLayer& ly, int D, float* bn) {
  for (int d = threadIdx.x; d < D; d += TR_THREADS) {
    bn[d] = Bf.theta_b[ly.bn_rm + d];
    bn[D + d] = rsqrtf(Bf.theta_b[ly.bn_rv + d] + TR_BN_EPS);
    bn[2 * D + d] = tr_softplus(tr_ldp(Bf.theta_p + ly.bn_uw + d)) + TR_BN_EPS;
    bn[3 * D + d] = tr_ldp(Bf.theta_p + ly.bn_bias + d);
  }
  __syncthreads();
}

// Gather the batch rows of one tile (layer-0 input), applying the layer permutation, and the
// per-row loss weights.  Padding rows are zero with weight zero.
__device__ __forceinline__ void tr_load_x(const TrBatch& bt, int tile, int D, const int* lperm,
                                          float* H1, float* c) {
  for (int e = threadIdx.x; e < D * TR_R; e += TR_THREADS) {
    const int r = e / D, j = e - r * D;  // consecutive threads read consecutive features of a row
    const int row = tile * TR_R + r;
    float v = 0.f;
    if (row < bt.B) {
      const int64_t gi = bt.perm ? bt.perm[bt.i0 + row] : bt.i0 + row;
      v = __ldg(bt.x + gi * D + (lperm ? lperm[j] : j));
    }
    H1[j * TR_R + r] = v;
  }
  if (threadIdx.x < TR_R) {
    const int row = tile * TR_R + threadIdx.x;
    float w = 0.f;
    if (row < bt.B) {
      const int64_t gi = bt.perm ? bt.perm[bt.i0 + row] : bt.i0 + row;
      w = bt.w ? __ldg(bt.w + gi) : 1.f;
    }
    c[threadIdx.x] = w;
  }
}

// H1[j][r] = BN(Yprev)[perm[j]][r]; optionally also X[d][r] = x-hat[d][r] (unpermuted).
__device__ __forceinline__ void tr_bn_apply_perm(const float* Yprev, const float* bn, int D,
                                                 const int* lperm, float* H1, float* xhat) {
  for (int e = threadIdx.x; e < D * TR_R; e += TR_THREADS) {
    const int j = e / TR_R, r = e - j * TR_R;
    const int d = lperm ? lperm[j] : j;
    const float xh = (Yprev[d * TR_R + r] - bn[d]) * bn[D + d];
    H1[j * TR_R + r] = fmaf(bn[2 * D + d], xh, bn[3 * D + d]);
    if (xhat) xhat[d * TR_R + r] = xh;
  }
}

// ---------------------------------------------------------------------------- RQ spline
// nflows' unconstrained_rational_quadratic_spline, linear tails, for one
// (row, transformed feature): p points at the feature's 3K-1 raw conditioner outputs in the
// feature-major tile (stride TR_R floats).  The backward pass is oracle/train_numpy.py::spline_backward.
constexpr int TR_MAXBINS = 16;
constexpr float TR_SPL_MIN = 1e-3f;  // min bin width / height / derivative

struct TrSpline {
  float pw[TR_MAXBINS], ph[TR_MAXBINS];  // softmax probabilities of widths / heights
  int k;                                  // bin
  bool inside;
  float W, Hh, d0, d1, th, dl, t, Nn, Dn, Q, c;
};

__device__ __forceinline__ void tr_spline_softmax(const float* p, int K, float isq, float* out) {
  float mx = -INFINITY;
  for (int j = 0; j < K; ++j) mx = fmaxf(mx, p[j * TR_R] * isq);
  float sum = 0.f;
  for (int j = 0; j < K; ++j) {
    out[j] = expf(p[j * TR_R] * isq - mx);
    sum += out[j];
  }
  const float inv = 1.f / sum;
  for (int j = 0; j < K; ++j) out[j] *= inv;
}
// knot j of a cumulative-size vector: s_0 = -B, s_K = +B (forced), else -B + 2B * cumsum
__device__ __forceinline__ float tr_spline_size(float prob, int K) {
  return TR_SPL_MIN + (1.f - TR_SPL_MIN * K) * prob;
}

__device__ __forceinline__ void tr_spline_eval(float x, const float* p, int K, float B, float isq,
                                               TrSpline& S, float& y, float& ld) {
  S.inside = (x >= -B) && (x <= B);
  tr_spline_softmax(p, K, isq, S.pw);
  tr_spline_softmax(p + K * TR_R, K, isq, S.ph);
  // bin = last knot <= x among s_0 .. s_{K-1} (nflows searchsorted with the last knot at B + 1e-6)
  const float xc = fminf(fmaxf(x, -B), B);
  float cum = 0.f, right = -B;
  int k = 0;
  float a = -B, a1 = B;
  for (int j = 0; j < K; ++j) {
    const float left = right;
    cum += tr_spline_size(S.pw[j], K);
    right = (j == K - 1) ? B : fmaf(2.f * B, cum, -B);
    if (j == 0 || xc >= left) k = j, a = left, a1 = right;
  }
  S.k = k;
  // height knots of bin k
  float c0 = -B, c1 = B;
  cum = 0.f;
  for (int j = 0; j < K; ++j) {
    const float lo = (j == 0) ? -B : fmaf(2.f * B, cum, -B);
    cum += tr_spline_size(S.ph[j], K);
    const float hi = (j == K - 1) ? B : fmaf(2.f * B, cum, -B);
    if (j == k) c0 = lo, c1 = hi;
  }
  const float* ud = p + 2 * K * TR_R;
  S.d0 = (k == 0) ? 1.f : TR_SPL_MIN + tr_softplus(ud[(k - 1) * TR_R]);
  S.d1 = (k == K - 1) ? 1.f : TR_SPL_MIN + tr_softplus(ud[k * TR_R]);
  S.W = a1 - a;
  S.Hh = c1 - c0;
  S.c = c0;
  S.th = (xc - a) / S.W;
  S.dl = S.Hh / S.W;
  S.t = S.th * (1.f - S.th);
  S.Nn = S.Hh * (S.dl * S.th * S.th + S.d0 * S.t);
  S.Dn = S.dl + (S.d0 + S.d1 - 2.f * S.dl) * S.t;
  S.Q = S.d1 * S.th * S.th + 2.f * S.dl * S.t + S.d0 * (1.f - S.th) * (1.f - S.th);
  if (S.inside) {
    y = S.c + S.Nn / S.Dn;
    ld = 2.f * logf(S.dl) + logf(S.Q) - 2.f * logf(S.Dn);
  } else {
    y = x;
    ld = 0.f;
  }
}

// Gradients of gy * y + gl * logdet w.r.t. x (returned) and the raw conditioner outputs
// (written to g, same layout as p).
__device__ __forceinline__ float tr_spline_backward(float x, const float* p, float* g, int K, float B,
                                                    float isq, float gy, float gl) {
  TrSpline S;
  float y, ld;
  tr_spline_eval(x, p, K, B, isq, S, y, ld);
  const int M = 3 * K - 1;
  if (!S.inside) {
    for (int j = 0; j < M; ++j) g[j * TR_R] = 0.f;
    return gy;
  }
  const float th = S.th, dl = S.dl, t = S.t, d0 = S.d0, d1 = S.d1;
  const float gN = gy / S.Dn, gD = -gy * S.Nn / (S.Dn * S.Dn);
  const float dDn = (d0 + d1 - 2.f * dl) * (1.f - 2.f * th);
  const float dQ = 2.f * d1 * th + 2.f * dl * (1.f - 2.f * th) - 2.f * d0 * (1.f - th);
  const float gth = gN * S.Hh * (2.f * dl * th + d0 * (1.f - 2.f * th)) + gD * dDn + gl * (dQ / S.Q - 2.f * dDn / S.Dn);
  const float gdl = gN * S.Hh * th * th + gD * (1.f - 2.f * t) +
                    gl * (2.f / dl + 2.f * t / S.Q - 2.f * (1.f - 2.f * t) / S.Dn);
  const float gH = gN * (dl * th * th + d0 * t) + gdl / S.W;
  const float gd0 = gN * S.Hh * t + gD * t + gl * ((1.f - th) * (1.f - th) / S.Q - 2.f * t / S.Dn);
  const float gd1 = gD * t + gl * (th * th / S.Q - 2.f * t / S.Dn);
  const float ga = -gth / S.W;
  const float gW = -gth * th / S.W - gdl * dl / S.W;
  const float gc = gy;
  const int k = S.k;
  // knots s_1 .. s_{K-1} are free (s_0, s_K constants): gs_k += g_left - g_size, gs_{k+1} += g_size;
  // d/d size_i = 2B * sum_{j > i, j <= K-1} gs_j; then the softmax backward
  const float scale = (1.f - TR_SPL_MIN * K) * 2.f * B;
  for (int which = 0; which < 2; ++which) {
    const float gl_ = which == 0 ? ga : gc, gs_ = which == 0 ? gW : gH;
    const float* pr = which == 0 ? S.pw : S.ph;
    const float gsk = (k >= 1) ? gl_ - gs_ : 0.f;       // knot k (constant when k == 0)
    const float gsk1 = (k + 1 <= K - 1) ? gs_ : 0.f;    // knot k + 1 (constant when k == K-1)
    // g size_i = scale * (gsk [i < k] + gsk1 [i < k + 1])
    float dot = 0.f;
    for (int i = 0; i < K; ++i) {
      const float gi = scale * ((i < k ? gsk : 0.f) + (i < k + 1 ? gsk1 : 0.f));
      dot += pr[i] * gi;
    }
    for (int i = 0; i < K; ++i) {
      const float gi = scale * ((i < k ? gsk : 0.f) + (i < k + 1 ? gsk1 : 0.f));
      g[(which * K + i) * TR_R] = pr[i] * (gi - dot) * isq;
    }
  }
  const float* ud = p + 2 * K * TR_R;
  for (int j = 0; j < K - 1; ++j) {
    // unnormalised derivative j is knot j + 1
    float gd = 0.f;
    if (j + 1 == k) gd = gd0;
    else if (j + 1 == k + 1) gd = gd1;
    g[(2 * K + j) * TR_R] = gd != 0.f ? gd * tr_sigmoid(ud[j * TR_R]) : 0.f;
  }
  return gth / S.W;
}

// Weights are staged in UNITS of one linear x one chunk of <= 64 output columns (so that a wide
// final layer -- 368 columns for the 32-D spline flow -- needs no more shared memory than a
// 64 x 64 one), double buffered: unit u lives in buffer u & 1.
__device__ __forceinline__ int tr_n_chunks(int n_out) { return (n_out + TR_CHUNK - 1) / TR_CHUNK; }
__device__ __forceinline__ void tr_prefetch_unit(const TrBuffers& Bf, const TrLayer& ly, const TrSmem& S,
                                                 int j, int q, int u, bool transposed, int max_dim) {
  const TrLinear& ln = ly.lin[j];
  const int c0 = q * TR_CHUNK, nc = min(TR_CHUNK, ln.n_out - c0);
  float* buf = S.W + (u & 1) * S.wbuf;
  tr_load_w_async(buf, Bf.theta_p + ln.w_off + (size_t)c0 * ln.n_in, nc, ln.n_in, transposed);
  if (transposed) tr_copy_async4(buf + S.wbuf - TR_CHUNK, Bf.theta_p + ln.b_off + c0, nc);
  tr_cp_commit();
}

// Conditioner + coupling of layer `ly` on one tile: H1 -> H2 (LU) -> V (net) -> Y.
// ld[r] += sum log s.  save != NULL: write the layer record (h2 | bufs 1.. | y) there.
// The caller has queued unit (linear 0, chunk 0) into buffer 0 as the most recent commit group.
__device__ __forceinline__ void tr_layer_forward(const TrBuffers& Bf, const TrPlan& P, const TrLayer& ly,
                                                 const TrSmem& S, float* save) {
  const int D = P.D, act = P.act;
  const int* itab = S.itab;
  if (ly.lu_bias >= 0) {
    tr_gemm(S.H2, S.H1, S.Wlu, D | 1, S.Wlu + D * (D | 1), nullptr, D, D, false);
  } else {
    for (int e = threadIdx.x; e < D * TR_R; e += TR_THREADS) S.H2[e] = S.H1[e];
  }
  __syncthreads();
  TR_T(1004);
  if (save) tr_copy(save, S.H2, D);
  for (int e = threadIdx.x; e < ly.d_id * TR_R; e += TR_THREADS) {
    const int i = e / TR_R, r = e - i * TR_R;
    S.V[e] = S.H2[itab[ly.id_off + i] * TR_R + r];
  }
  __syncthreads();
  TR_T(1005);
  int u = 0;
  for (int j = 0; j < ly.n_lin; ++j) {
    const TrLinear& ln = ly.lin[j];
    const float* src = S.V + ly.buf_off[ln.in_buf] * TR_R;
    const float* Aop = src;
    if (ln.pre_act) {
      // (writing the activated operand from the producing GEMM's epilogue instead of in this pass
      // was measured SLOWER: +0.4 us per linear)
      for (int e = threadIdx.x; e < ln.n_in * TR_R; e += TR_THREADS) S.A[e] = tr_act(act, src[e]);
      Aop = S.A;
    }
    const int nq = tr_n_chunks(ln.n_out);
    for (int q = 0; q < nq; ++q, ++u) {
      const bool more = q + 1 < nq || j + 1 < ly.n_lin;
      if (more) {
        tr_prefetch_unit(Bf, ly, S, q + 1 < nq ? j : j + 1, q + 1 < nq ? q + 1 : 0, u + 1, true, P.max_dim);
        tr_cp_wait<1>();
      } else {
        tr_cp_wait<0>();
      }
      __syncthreads();
      const int c0 = q * TR_CHUNK, nc = min(TR_CHUNK, ln.n_out - c0);
      const float* buf = S.W + (u & 1) * S.wbuf;
      tr_gemm(S.V + (ly.buf_off[ln.out_buf] + c0) * TR_R, Aop, buf, nc | 1, buf + S.wbuf - TR_CHUNK,
              ln.res_buf >= 0 ? S.V + (ly.buf_off[ln.res_buf] + c0) * TR_R : nullptr, nc, ln.n_in, false);
      __syncthreads();
      TR_T(1010 + j);
    }
  }
  const float* prm = S.V + ly.buf_off[ly.n_buf - 1] * TR_R;
  for (int e = threadIdx.x; e < ly.d_id * TR_R; e += TR_THREADS) {
    const int i = e / TR_R, r = e - i * TR_R;
    const int f = itab[ly.id_off + i];
    S.Y[f * TR_R + r] = S.H2[f * TR_R + r];
  }
  for (int e = threadIdx.x; e < ly.d_tr * TR_R; e += TR_THREADS) {
    const int i = e / TR_R, r = e - i * TR_R;
    const int f = itab[ly.tr_off + i];
    const float t = S.H2[f * TR_R + r];
    if (P.num_bins > 0) {
      TrSpline sp;
      float y, ld;
      tr_spline_eval(t, prm + (size_t)i * (3 * P.num_bins - 1) * TR_R + r, P.num_bins, __int_as_float(P.tail_bound),
                     rsqrtf((float)P.hidden), sp, y, ld);
      S.Y[f * TR_R + r] = y;
      S.A[e] = ld;
    } else if (P.additive == 2) {
      // MaskedAffineAutoregressiveTransform: params (D, 2) = (unconstrained scale, shift)
      const float s = tr_softplus(prm[(2 * i) * TR_R + r]) + 1e-3f;
      S.Y[f * TR_R + r] = fmaf(t, s, prm[(2 * i + 1) * TR_R + r]);
      S.A[e] = logf(s);
    } else if (P.additive == 1) {
      S.Y[f * TR_R + r] = t + prm[e];
      S.A[e] = 0.f;
    } else {
      const float s = tr_sigmoid(prm[ly.d_tr * TR_R + e] + 2.f) + 1e-3f;
      S.Y[f * TR_R + r] = fmaf(t, s, prm[e]);
      S.A[e] = logf(s);
    }
  }
  __syncthreads();
  TR_T(1006);
  if (threadIdx.x < TR_R) {
    float s = 0.f;
    for (int i = 0; i < ly.d_tr; ++i) s += S.A[i * TR_R + threadIdx.x];
    S.ld[threadIdx.x] += s;
  }
  if (save) {
    const int nb = ly.rec_floats - 2 * D;
    tr_copy(save + D * TR_R, S.V + ly.buf_off[1] * TR_R, nb);
    tr_copy(save + (D + nb) * TR_R, S.Y, D);
  }
  __syncthreads();
}

// Row-constant log|det| of the flow: LU diagonals + BatchNorm scales.  stats == NULL: eval mode
// (running variance from theta_b); otherwise the published batch variances.
__device__ __forceinline__ float tr_const_logdet(const TrBuffers& Bf, const TrPlan& P, const float* stats,
                                                 float* red) {
  const int D = P.D;
  float s = 0.f;
  for (int e = threadIdx.x; e < P.L * D; e += TR_THREADS) {
    const int l = e / D, d = e - l * D;
    const TrLayer& ly = P.layer[l];
    if (ly.lu_bias >= 0) s += logf(tr_softplus(tr_ldp(Bf.theta_p + ly.lu_diag + d)) + TR_LU_EPS);
    if (ly.bn_uw >= 0) {
      const float var = stats ? stats[(l * 2 + 1) * D + d] : Bf.theta_b[ly.bn_rv + d];
      s += logf(tr_softplus(tr_ldp(Bf.theta_p + ly.bn_uw + d)) + TR_BN_EPS) - 0.5f * logf(var + TR_BN_EPS);
    }
  }
  return tr_block_sum(s, red);
}

// ============================================================================ FWD(l)
#ifdef NB200_SIMT_SHIM
static float* const tr_smem_dyn = reinterpret_cast<float*>(simt::dynamic_smem);
#else
extern __shared__ __align__(16) float tr_smem_dyn[];
#endif

// Phases are __device__ functions over a carved shared-memory map whose index tables are staged:
// the persistent kernel (tr_train_kernel) runs them between grid barriers; the one-phase kernels
// below it launch them one at a time (CPU SIMT shim of tests/_hostcheck, NB200_TR_CHAIN=1).
// `part`: TR_PROLOGUE = what does not depend on the other CTAs' previous phase (the persistent kernel
// runs it while the grid barrier is pending), TR_MAIN = the rest, TR_WHOLE = both.
constexpr int TR_PROLOGUE = 1, TR_MAIN = 2, TR_WHOLE = 3;
constexpr int TR_KEEP_BN = 4;  // (backward prologue) bn[0..3][D] of this layer is still in shared memory

__device__ __forceinline__ void tr_fwd_phase(const TrPlan& P, const TrBuffers& Bf, const TrBatch& bt, int l,
                                             const TrSmem& S, int part = TR_WHOLE) {
  const TrLayer& ly = P.layer[l];
  const int D = P.D;
  const int* lperm = ly.perm_off >= 0 ? S.itab + ly.perm_off : nullptr;
  const bool bn_prev = l > 0 && P.layer[l - 1].bn_uw >= 0;
  if (part & TR_PROLOGUE) {
    if (ly.lu_bias >= 0) tr_lu_dense(S.Wlu, S.stage, Bf.theta_p, ly, D, true);
    if (bn_prev) tr_bn_affine(Bf, P.layer[l - 1], D, S.bn);
    TR_T(1001);
  }
  if (!(part & TR_MAIN)) return;
  if (bn_prev) tr_bn_stats_first(Bf, P.layer[l - 1], l - 1, D, bt.B, S.bn, S.bn + 4 * D);
  __syncthreads();
  TR_T(1002);
  float n_run = 0.f, mean_run = 0.f, m2_run = 0.f;  // BatchNorm partials of feature threadIdx.x
  float wsum = 0.f;
  int rows_cta = 0;
  for (int tile = blockIdx.x; tile < bt.n_tiles; tile += gridDim.x) {
    float* rec = Bf.ws + ((size_t)tile * P.rec_total + ly.ws_off) * TR_R;
    if (l == 0) {
      tr_prefetch_unit(Bf, ly, S, 0, 0, 0, true, P.max_dim);
      tr_load_x(bt, tile, D, lperm, S.H1, S.c);
      if (threadIdx.x < TR_R) S.ld[threadIdx.x] = 0.f;
      __syncthreads();
      if (threadIdx.x < TR_R) Bf.crow[tile * TR_R + threadIdx.x] = S.c[threadIdx.x];
      if (threadIdx.x == 0)
        for (int r = 0; r < TR_R; ++r) wsum += S.c[r];
    } else {
      const TrLayer& lp = P.layer[l - 1];
      const float* yprev = Bf.ws + ((size_t)tile * P.rec_total + lp.ws_off + lp.rec_floats - D) * TR_R;
      tr_copy_async(S.Y, yprev, D);
      tr_cp_commit();
      tr_prefetch_unit(Bf, ly, S, 0, 0, 0, true, P.max_dim);
      if (threadIdx.x < TR_R) S.ld[threadIdx.x] = Bf.ldrow[tile * TR_R + threadIdx.x];
      tr_cp_wait<1>();
      __syncthreads();
      if (bn_prev) {
        tr_bn_apply_perm(S.Y, S.bn, D, lperm, S.H1, nullptr);
      } else {
        for (int e = threadIdx.x; e < D * TR_R; e += TR_THREADS) {
          const int j = e / TR_R, r = e - j * TR_R;
          S.H1[e] = S.Y[(lperm ? lperm[j] : j) * TR_R + r];
        }
      }
      __syncthreads();
    }
    TR_T(1003);
    tr_layer_forward(Bf, P, ly, S, rec);
    TR_T(1009);
    if (threadIdx.x < TR_R) Bf.ldrow[tile * TR_R + threadIdx.x] = S.ld[threadIdx.x];
    const int nv = min(TR_R, bt.B - tile * TR_R);
    rows_cta += nv;
    if (ly.bn_uw >= 0 && (int)threadIdx.x < D) {
      const float* y = S.Y + threadIdx.x * TR_R;
      float m = 0.f;
      for (int r = 0; r < nv; ++r) m += y[r];
      m /= (float)nv;
      float q = 0.f;
      for (int r = 0; r < nv; ++r) q += (y[r] - m) * (y[r] - m);
      const float nb = (float)nv, nn = n_run + nb, delta = m - mean_run;
      mean_run += delta * (nb / nn);
      m2_run += q + delta * delta * (n_run * nb / nn);
      n_run = nn;
    }
    __syncthreads();
  }
  if (ly.bn_uw >= 0 && (int)threadIdx.x < D) {
    float* p = Bf.stat_part + ((size_t)(l * Bf.G + blockIdx.x) * 2) * D;
    p[threadIdx.x] = mean_run;
    p[D + threadIdx.x] = m2_run;
  }
  if (threadIdx.x == 0) {
    if (l == 0) {
      Bf.stat_n[blockIdx.x] = (float)rows_cta;
      Bf.wsum_part[blockIdx.x] = wsum;
    }
  }
}

// ============================================================================ LOSS
// z = BN_{L-1}(y_{L-1}); loss partial; dout_{L-1} = c_r z; BatchNorm backward sums of layer L-1.
__device__ __forceinline__ void tr_loss_phase(const TrPlan& P, const TrBuffers& Bf, const TrBatch& bt,
                                              const TrSmem& S, int part = TR_WHOLE) {
  const int D = P.D, L = P.L;
  const TrLayer& ly = P.layer[L - 1];
  const bool bn = ly.bn_uw >= 0;
  if ((part & TR_PROLOGUE) && bn) tr_bn_affine(Bf, ly, D, S.bn);
  if (!(part & TR_MAIN)) return;
  if (bn) tr_bn_stats_first(Bf, ly, L - 1, D, bt.B, S.bn, S.bn + 4 * D);
  __syncthreads();
  // the last layer's statistics are published by one block only: use our own copy for the constant
  float cld = 0.f;
  {
    float s = 0.f;
    for (int e = threadIdx.x; e < L * D; e += TR_THREADS) {
      const int l = e / D, d = e - l * D;
      const TrLayer& lq = P.layer[l];
      if (lq.lu_bias >= 0) s += logf(tr_softplus(tr_ldp(Bf.theta_p + lq.lu_diag + d)) + TR_LU_EPS);
      if (lq.bn_uw >= 0) {
        const float var = (l == L - 1) ? S.bn[4 * D + d] : Bf.stats[(l * 2 + 1) * D + d];
        s += logf(tr_softplus(tr_ldp(Bf.theta_p + lq.bn_uw + d)) + TR_BN_EPS) - 0.5f * logf(var + TR_BN_EPS);
      }
    }
    cld = tr_block_sum(s, S.red);
  }
  float csum = 0.f;
  {
    float s = 0.f;
    for (int g = threadIdx.x; g < Bf.G; g += TR_THREADS) s += Bf.wsum_part[g];
    csum = tr_block_sum(s, S.red);
  }
  const float inv_csum = 1.f / csum;
  // base distribution N(0, var I) (flows/distributions.py:45-56)
  const float bvar = __int_as_float(P.base_var), binv = 1.f / bvar;
  const float blogz = (float)D * (TR_HALF_LOG_2PI + 0.5f * logf(bvar));
  float s1 = 0.f, s2 = 0.f, loss = 0.f;
  float* dout = Bf.dout[(L - 1) & 1];
  for (int tile = blockIdx.x; tile < bt.n_tiles; tile += gridDim.x) {
    const float* y = Bf.ws + ((size_t)tile * P.rec_total + ly.ws_off + ly.rec_floats - D) * TR_R;
    tr_copy_async(S.Y, y, D);
    tr_cp_commit();
    if (threadIdx.x < TR_R) {
      S.c[threadIdx.x] = Bf.crow[tile * TR_R + threadIdx.x] * inv_csum;
      S.ld[threadIdx.x] = Bf.ldrow[tile * TR_R + threadIdx.x];
    }
    tr_cp_wait<0>();
    __syncthreads();
    if (threadIdx.x < TR_R) Bf.crow[tile * TR_R + threadIdx.x] = S.c[threadIdx.x];
    // X = x-hat, H1 = z, H2 = dout
    for (int e = threadIdx.x; e < D * TR_R; e += TR_THREADS) {
      const int d = e / TR_R, r = e - d * TR_R;
      float z = S.Y[e], xh = 0.f;
      if (bn) {
        xh = (z - S.bn[d]) * S.bn[D + d];
        z = fmaf(S.bn[2 * D + d], xh, S.bn[3 * D + d]);
      }
      S.X[e] = xh;
      S.H1[e] = z;
      S.H2[e] = S.c[r] * z * binv;
    }
    __syncthreads();
    tr_copy(dout + (size_t)tile * D * TR_R, S.H2, D);
    if (threadIdx.x < TR_R) {
      float q = 0.f;
      for (int d = 0; d < D; ++d) q += S.H1[d * TR_R + threadIdx.x] * S.H1[d * TR_R + threadIdx.x];
      const float logp = -0.5f * binv * q - blogz + S.ld[threadIdx.x] + cld;
      loss += S.c[threadIdx.x] * logp;  // c == 0 for padding rows
    }
    if (bn && (int)threadIdx.x < D) {
      for (int r = 0; r < TR_R; ++r) {
        const float g = S.H2[threadIdx.x * TR_R + r];
        s1 += g;
        s2 = fmaf(g, S.X[threadIdx.x * TR_R + r], s2);
      }
    }
    __syncthreads();
  }
  if (bn && (int)threadIdx.x < D) {
    float* sp = Bf.s_part[(L - 1) & 1] + (size_t)blockIdx.x * 2 * D;
    sp[threadIdx.x] = s1;
    sp[D + threadIdx.x] = s2;
  }
  const float lsum = tr_block_sum(threadIdx.x < TR_R ? loss : 0.f, S.red);
  if (threadIdx.x == 0) Bf.loss_part[blockIdx.x] = -lsum;
}

// ============================================================================ BWD(l)
// Queue (one commit group) the loads of a tile for BWD(l): the layer's record h2 | buffers | y, the
// gradient w.r.t. the layer's output, and the previous layer's output.
__device__ __forceinline__ void tr_bwd_loads(const TrPlan& P, const TrBuffers& Bf, const TrLayer& ly, int l,
                                             const TrSmem& S, int tile) {
  const int D = P.D;
  const float* rec = Bf.ws + ((size_t)tile * P.rec_total + ly.ws_off) * TR_R;
  const int nb = ly.rec_floats - 2 * D;
  tr_copy_async(S.H2, rec, D);
  tr_copy_async(S.V + ly.buf_off[1] * TR_R, rec + D * TR_R, nb);
  tr_copy_async(S.Y, rec + (D + nb) * TR_R, D);
  tr_copy_async(S.X, Bf.dout[l & 1] + (size_t)tile * D * TR_R, D);
  if (l > 0) {
    const TrLayer& lp = P.layer[l - 1];
    tr_copy_async(S.Pf, Bf.ws + ((size_t)tile * P.rec_total + lp.ws_off + lp.rec_floats - D) * TR_R, D);
  }
  tr_cp_commit();
}

__device__ __forceinline__ void tr_bwd_phase(const TrPlan& P, const TrBuffers& Bf, const TrBatch& bt, int l,
                                             const TrSmem& S, int part_sel = TR_WHOLE) {
  const TrLayer& ly = P.layer[l];
  const int D = P.D, act = P.act;
  const int* itab = S.itab;
  const int* lperm = ly.perm_off >= 0 ? S.itab + ly.perm_off : nullptr;
  const bool bn = ly.bn_uw >= 0;
  const bool bn_prev = l > 0 && P.layer[l - 1].bn_uw >= 0;
  float* bnp = S.bn;            // this layer: mean, rstd, w, beta
  float* sS = S.bn + 4 * D;     // S1, S2 of this layer
  float* part = Bf.part + (size_t)blockIdx.x * Bf.part_stride;
  if (part_sel & TR_PROLOGUE) {
    if (bn && !(part_sel & TR_KEEP_BN)) {
      tr_bn_stats_published(Bf, l, D, bnp);
      tr_bn_affine(Bf, ly, D, bnp);
    }
    if (ly.lu_bias >= 0) tr_lu_dense(S.Wlu, S.stage, Bf.theta_p, ly, D, false);
    // this CTA's first tile: its saved activations and the gradient arriving from the layer
    // above are its OWN data (written by its earlier phases), so their loads can be in flight
    // while the grid barrier is pending
    if ((int)blockIdx.x < bt.n_tiles) tr_bwd_loads(P, Bf, ly, l, S, blockIdx.x);
    TR_T(2001);
  }
  if (!(part_sel & TR_MAIN)) return;
  __syncthreads();
  if (bn) {
    tr_colsum(Bf.s_part[l & 1], Bf.G, 2 * D, sS, S.red);
    // the gradient of the BatchNorm parameters is complete here: CTA 0's partial carries it
    for (int d = threadIdx.x; d < D; d += TR_THREADS) {
      const float dw = sS[D + d] - 1.f / bnp[2 * D + d];
      part[ly.bn_bias + d] = blockIdx.x == 0 ? sS[d] : 0.f;
      part[ly.bn_uw + d] = blockIdx.x == 0 ? dw * tr_sigmoid(tr_ldp(Bf.theta_p + ly.bn_uw + d)) : 0.f;
    }
  }
  __syncthreads();
  TR_T(2002);
  const float invB = 1.f / (float)bt.B, invBm1 = 1.f / (float)(bt.B - 1);
  float p1 = 0.f, p2 = 0.f;  // BatchNorm backward sums of layer l-1 (feature threadIdx.x)
  bool first = true;
  const float* dout_in = Bf.dout[l & 1];
  float* dout_out = Bf.dout[(l - 1) & 1];
  for (int tile = blockIdx.x; tile < bt.n_tiles; tile += gridDim.x) {
    if (tile != (int)blockIdx.x) tr_bwd_loads(P, Bf, ly, l, S, tile);  // (the first tile: by the prologue)
    tr_prefetch_unit(Bf, ly, S, ly.n_lin - 1, 0, 0, false, P.max_dim);
    if (l == 0) tr_load_x(bt, tile, D, lperm, S.H1, S.ld);  // S.ld: scratch for the row weights
    if (threadIdx.x < TR_R) S.c[threadIdx.x] = Bf.crow[tile * TR_R + threadIdx.x];
    for (int e = threadIdx.x; e < P.vals_floats * TR_R; e += TR_THREADS) S.Gv[e] = 0.f;
    tr_cp_wait<1>();
    __syncthreads();
    TR_T(2003);
    // dy (in place in X): BatchNorm backward
    if (bn) {
      for (int e = threadIdx.x; e < D * TR_R; e += TR_THREADS) {
        const int d = e / TR_R, r = e - d * TR_R;
        const float w = bnp[2 * D + d];
        const float xh = (S.Y[e] - bnp[d]) * bnp[D + d];
        const float v = S.X[e] - xh * (sS[D + d] - 1.f / w) * invBm1 - sS[d] * invB;
        S.X[e] = (tile * TR_R + r < bt.B) ? w * bnp[D + d] * v : 0.f;
      }
    }
    for (int e = threadIdx.x; e < ly.d_id * TR_R; e += TR_THREADS) {
      const int i = e / TR_R, r = e - i * TR_R;
      S.V[e] = S.H2[itab[ly.id_off + i] * TR_R + r];
    }
    __syncthreads();
    // coupling backward: Gv[last] = d params, Y := d h2 (transformed half)
    {
      const int last = ly.buf_off[ly.n_buf - 1] * TR_R;
      const float* prm = S.V + last;
      float* gprm = S.Gv + last;
      for (int e = threadIdx.x; e < ly.d_tr * TR_R; e += TR_THREADS) {
        const int i = e / TR_R, r = e - i * TR_R;
        const int f = itab[ly.tr_off + i];
        const float dt2 = S.X[f * TR_R + r];
        if (P.num_bins > 0) {
          const size_t o = (size_t)i * (3 * P.num_bins - 1) * TR_R + r;
          S.Y[f * TR_R + r] = tr_spline_backward(S.H2[f * TR_R + r], prm + o, gprm + o, P.num_bins,
                                                 __int_as_float(P.tail_bound), rsqrtf((float)P.hidden), dt2,
                                                 -S.c[r]);
        } else if (P.additive == 2) {
          const float u = prm[(2 * i) * TR_R + r];
          const float s = tr_softplus(u) + 1e-3f;
          const float ds = dt2 * S.H2[f * TR_R + r] - S.c[r] / s;
          gprm[(2 * i) * TR_R + r] = ds * tr_sigmoid(u);
          gprm[(2 * i + 1) * TR_R + r] = dt2;
          S.Y[f * TR_R + r] = dt2 * s;
        } else if (P.additive == 1) {
          gprm[e] = dt2;
          S.Y[f * TR_R + r] = dt2;
        } else {
          const float sg = tr_sigmoid(prm[ly.d_tr * TR_R + e] + 2.f);
          const float s = sg + 1e-3f;
          const float ds = dt2 * S.H2[f * TR_R + r] - S.c[r] / s;
          gprm[e] = dt2;
          gprm[ly.d_tr * TR_R + e] = ds * sg * (1.f - sg);
          S.Y[f * TR_R + r] = dt2 * s;
        }
      }
    }
    __syncthreads();
    TR_T(2004);
    // conditioner backward
    int u = 0;
    for (int j = ly.n_lin - 1; j >= 0; --j) {
      const TrLinear& ln = ly.lin[j];
      const float* src = S.V + ly.buf_off[ln.in_buf] * TR_R;
      const float* delta = S.Gv + ly.buf_off[ln.out_buf] * TR_R;
      const float* Aop = src;
      if (ln.pre_act) {
        for (int e = threadIdx.x; e < ln.n_in * TR_R; e += TR_THREADS) S.A[e] = tr_act(act, src[e]);
        Aop = S.A;
      }
      const int nq = tr_n_chunks(ln.n_out);
      for (int q = 0; q < nq; ++q, ++u) {
        const bool more = q + 1 < nq || j > 0;
        if (more) {
          tr_prefetch_unit(Bf, ly, S, q + 1 < nq ? j : j - 1, q + 1 < nq ? q + 1 : 0, u + 1, false, P.max_dim);
          tr_cp_wait<1>();
        } else {
          tr_cp_wait<0>();
        }
        __syncthreads();
        const int c0 = q * TR_CHUNK, nc = min(TR_CHUNK, ln.n_out - c0);
        tr_wgrad(part + ln.w_off + (size_t)c0 * ln.n_in, delta + c0 * TR_R, Aop, nc, ln.n_in, first);
        tr_bgrad(part + ln.b_off + c0, delta + c0 * TR_R, nc, first);
        // A2 (+)= W[c0 : c0 + nc]^T delta[c0 : c0 + nc] (input gradient before the activation derivative)
        tr_gemm(S.A2, delta + c0 * TR_R, S.W + (u & 1) * S.wbuf, ln.n_in, nullptr, nullptr, ln.n_in, nc, q > 0);
        __syncthreads();
      }
      float* gin = S.Gv + ly.buf_off[ln.in_buf] * TR_R;
      for (int e = threadIdx.x; e < ln.n_in * TR_R; e += TR_THREADS)
        gin[e] += ln.pre_act ? S.A2[e] * tr_dact(act, src[e]) : S.A2[e];
      if (ln.res_buf >= 0) {
        float* gres = S.Gv + ly.buf_off[ln.res_buf] * TR_R;
        for (int e = threadIdx.x; e < ln.n_out * TR_R; e += TR_THREADS) gres[e] += delta[e];
      }
      __syncthreads();
      TR_T(2010 + j);
    }
    // d h2 (identity half) = dy + d(net input)
    for (int e = threadIdx.x; e < ly.d_id * TR_R; e += TR_THREADS) {
      const int i = e / TR_R, r = e - i * TR_R;
      const int f = itab[ly.id_off + i];
      // MAF: every feature is both conditioner input and transformed -> add the two paths
      S.Y[f * TR_R + r] = (P.additive == 2 ? S.Y[f * TR_R + r] : S.X[f * TR_R + r]) + S.Gv[e];
    }
    // layer input h1 (and x-hat of the previous BatchNorm, into X)
    __syncthreads();
    if (l > 0) {
      if (bn_prev) {
        // previous layer's BatchNorm constants: published by FWD(l) block 0
        const TrLayer& lp = P.layer[l - 1];
        for (int e = threadIdx.x; e < D * TR_R; e += TR_THREADS) {
          const int j = e / TR_R, r = e - j * TR_R;
          const int d = lperm ? lperm[j] : j;
          const float mean = Bf.stats[((l - 1) * 2) * D + d];
          const float rstd = rsqrtf(Bf.stats[((l - 1) * 2 + 1) * D + d] + TR_BN_EPS);
          const float w = tr_softplus(tr_ldp(Bf.theta_p + lp.bn_uw + d)) + TR_BN_EPS;
          const float xh = (S.Pf[d * TR_R + r] - mean) * rstd;
          S.H1[j * TR_R + r] = fmaf(w, xh, tr_ldp(Bf.theta_p + lp.bn_bias + d));
          S.X[d * TR_R + r] = xh;
        }
      } else {
        for (int e = threadIdx.x; e < D * TR_R; e += TR_THREADS) {
          const int j = e / TR_R, r = e - j * TR_R;
          S.H1[e] = S.Pf[(lperm ? lperm[j] : j) * TR_R + r];
        }
      }
    }
    __syncthreads();
    TR_T(2005);
    const float* dh1 = S.Y;
    if (ly.lu_bias >= 0) {
      tr_wgrad(part + ly.lu_part_off, S.Y, S.H1, D, D, first);
      tr_bgrad(part + ly.lu_bias, S.Y, D, first);
      tr_gemm(S.A, S.Y, S.Wlu, D, nullptr, nullptr, D, D, false);
      dh1 = S.A;
      __syncthreads();
    }
    if (l > 0) {
      // un-permute into H2, store, and accumulate the previous BatchNorm's backward sums
      for (int e = threadIdx.x; e < D * TR_R; e += TR_THREADS) {
        const int j = e / TR_R, r = e - j * TR_R;
        S.H2[(lperm ? lperm[j] : j) * TR_R + r] = dh1[e];
      }
      __syncthreads();
      tr_copy(dout_out + (size_t)tile * D * TR_R, S.H2, D);
      if (bn_prev && (int)threadIdx.x < D) {
        for (int r = 0; r < TR_R; ++r) {
          const float g = S.H2[threadIdx.x * TR_R + r];
          p1 += g;
          p2 = fmaf(g, S.X[threadIdx.x * TR_R + r], p2);
        }
      }
    }
    first = false;
    __syncthreads();
    TR_T(2006);
  }
  if (first) {
    // this CTA had no tile: its partial vector must still be defined for REDUCE
    for (int j = 0; j < ly.n_lin; ++j) {
      const TrLinear& ln = ly.lin[j];
      for (int e = threadIdx.x; e < ln.n_in * ln.n_out; e += TR_THREADS) part[ln.w_off + e] = 0.f;
      for (int e = threadIdx.x; e < ln.n_out; e += TR_THREADS) part[ln.b_off + e] = 0.f;
    }
    if (ly.lu_bias >= 0) {
      for (int e = threadIdx.x; e < D * D; e += TR_THREADS) part[ly.lu_part_off + e] = 0.f;
      for (int e = threadIdx.x; e < D; e += TR_THREADS) part[ly.lu_bias + e] = 0.f;
    }
  }
  if (ly.lu_bias >= 0) {
    // dense dW of this CTA (accumulated over its tiles above) -> lower / upper / diagonal partials
    __syncthreads();
    tr_lu_chain(part, part + ly.lu_part_off, S.stage, S.stage + D * D, Bf.theta_p, ly, D, blockIdx.x == 0);
  }
  if (bn_prev && (int)threadIdx.x < D) {
    float* sp = Bf.s_part[(l - 1) & 1] + (size_t)blockIdx.x * 2 * D;
    sp[threadIdx.x] = p1;
    sp[D + threadIdx.x] = p2;
  }
}

// ============================================================================ REDUCE
// grad[i] = sum over the CTAs of part[g][i] for every parameter (times the MADE mask), and
// gn_part[blockIdx.x] = this block's share of |grad|^2.  Every parameter's partial is complete in
// the partial vectors (the backward phases apply the LU chain rule and carry the BatchNorm
// gradients themselves), so this is one dense, fixed-order, vectorised sum: columns of four
// parameters are dealt out in contiguous slices, one per CTA; inside a CTA the G partials of a
// column are split over `nsub` threads -- every 16-byte load independent -- and combined through
// shared memory in a fixed order (no atomics: bit-reproducible for a given grid).
// scratch: >= TR_THREADS float4; red: >= TR_THREADS floats.
__device__ __forceinline__ void tr_reduce_phase(const TrPlan& P, const TrBuffers& Bf, float4* scratch, float* red) {
  const int n = P.n_params, n4 = (n + 3) / 4, G = Bf.G;
  const int per = (n4 + (int)gridDim.x - 1) / (int)gridDim.x;
  const int c_begin = blockIdx.x * per, c_end = min(n4, c_begin + per);
  float sq = 0.f;
  for (int c0 = c_begin; c0 < c_end; c0 += TR_THREADS) {
    const int ncol = min(TR_THREADS, c_end - c0);
    const int nsub = min(TR_THREADS / ncol, 16);
    const int sub = threadIdx.x / ncol, col = threadIdx.x - sub * ncol;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sub < nsub) {
      const float4* src = reinterpret_cast<const float4*>(Bf.part) + (c0 + col);
      const size_t stride4 = (size_t)Bf.part_stride / 4;
#pragma unroll 4
      for (int g = sub; g < G; g += nsub) {
        const float4 v = src[(size_t)g * stride4];
        acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
      }
    }
    scratch[threadIdx.x] = acc;  // [sub][col]
    __syncthreads();
    if ((int)threadIdx.x < ncol) {
      float4 t = scratch[threadIdx.x];
      for (int q = 1; q < nsub; ++q) {
        const float4 v = scratch[q * ncol + threadIdx.x];
        t.x += v.x, t.y += v.y, t.z += v.z, t.w += v.w;
      }
      const int i = 4 * (c0 + threadIdx.x);
      float g4[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (i + k < n) {
          if (Bf.pmask) g4[k] *= Bf.pmask[i + k];
          sq = fmaf(g4[k], g4[k], sq);
        } else {
          g4[k] = 0.f;
        }
      }
      reinterpret_cast<float4*>(Bf.grad)[c0 + threadIdx.x] = make_float4(g4[0], g4[1], g4[2], g4[3]);
    }
    __syncthreads();
  }
  const float tot = tr_block_sum<TR_THREADS>(sq, red);
  if (threadIdx.x == 0) Bf.gn_part[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(TR_THREADS) tr_reduce_kernel(const __grid_constant__ TrPlan P, TrBuffers Bf) {
  __shared__ float4 scratch[TR_THREADS];
  __shared__ float red[TR_THREADS];
  tr_reduce_phase(P, Bf, scratch, red);
}

// ============================================================================ ADAM
// clip_grad_norm_ + torch.optim.AdamW / Adam / SGD on the flat parameter vector.
// d_loss[0] = this step's loss, d_loss[1] = gradient norm before clipping.  red: >= 2 floats.
template <int NT>
__device__ __forceinline__ void tr_adam_phase(const TrBuffers& Bf, int n_params, const TrOptim& o,
                                              float* __restrict__ m, float* __restrict__ v, float* d_loss,
                                              float* d_loss_accum, float* red) {
  __syncthreads();
  if (threadIdx.x < 32) {
    // warp 0: |grad| from the REDUCE blocks' shares (fixed order), and the clip coefficient
    float gsum = 0.f;
    for (int i = threadIdx.x; i < Bf.n_reduce_blocks; i += 32) gsum += Bf.gn_part[i];
    const float gn = sqrtf(tr_warp_sum(gsum));
    // torch.nn.utils.clip_grad_norm_: clamp(clip / (norm + 1e-6), max = 1), NaN propagates
    const float c = o.clip / (gn + 1e-6f);
    if (threadIdx.x == 0) red[0] = o.clip > 0.f ? (c >= 1.f ? 1.f : c) : 1.f;
    if (blockIdx.x == 0) {
      float lsum = 0.f;
      for (int g = threadIdx.x; g < Bf.G; g += 32) lsum += Bf.loss_part[g];
      const float loss = tr_warp_sum(lsum);
      if (threadIdx.x == 0) {
        if (d_loss) d_loss[0] = loss, d_loss[1] = gn;
        if (d_loss_accum) d_loss_accum[0] += loss;
      }
    }
  }
  __syncthreads();
  const float coef = red[0];
  const int n = n_params;
  for (int i = blockIdx.x * NT + threadIdx.x; i < n; i += gridDim.x * NT) {
    float g = Bf.grad[i] * coef;
    if (o.kind < 0) {
      Bf.grad[i] = g;
      continue;
    }
    float p = Bf.theta_p[i];
    if (o.kind == 2) {
      if (o.weight_decay != 0.f) g = fmaf(o.weight_decay, p, g);
      Bf.theta_p[i] = p - o.lr * g;
      continue;
    }
    if (o.kind == 0) p *= 1.f - o.lr * o.weight_decay;
    else if (o.weight_decay != 0.f) g = fmaf(o.weight_decay, p, g);
    const float mi = o.beta1 * m[i] + (1.f - o.beta1) * g;
    const float vi = o.beta2 * v[i] + (1.f - o.beta2) * g * g;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(o.bc2) + o.eps;
    Bf.theta_p[i] = p - (o.lr / o.bc1) * (mi / denom);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) tr_adam_kernel(TrBuffers Bf, int n_params, TrOptim o, float* __restrict__ m,
                                                      float* __restrict__ v, float* d_loss,
                                                      float* d_loss_accum) {
  __shared__ float red[2];
  tr_adam_phase<256>(Bf, n_params, o, m, v, d_loss, d_loss_accum, red);
}

// theta_p *= mask: masked MADE weights are kept at exactly zero (they never influence the
// reference's outputs either: nflows multiplies by the mask in every forward pass)
__global__ void tr_mask_params_kernel(float* __restrict__ theta_p, const float* __restrict__ mask, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) theta_p[i] *= mask[i];
}

// ============================================================================ EVAL
// Eval-mode loss (FlowModel._validate, flowmodel/base.py:454-523): running statistics, every layer
// of a tile in one pass.  out_part[g] = {sum c logp, sum c}; also per-row log_prob when d_logp.
__device__ __forceinline__ void tr_eval_phase(const TrPlan& P, const TrBuffers& Bf, const TrBatch& bt,
                                              float* out_part, float* d_logp, const TrSmem& S) {
  const int D = P.D, L = P.L;
  const float cld = tr_const_logdet(Bf, P, nullptr, S.red);
  const float bvar = __int_as_float(P.base_var), binv = 1.f / bvar;
  const float blogz = (float)D * (TR_HALF_LOG_2PI + 0.5f * logf(bvar));
  float loss = 0.f, csum = 0.f;
  for (int tile = blockIdx.x; tile < bt.n_tiles; tile += gridDim.x) {
    for (int l = 0; l < L; ++l) {
      const TrLayer& ly = P.layer[l];
      const int* lperm = ly.perm_off >= 0 ? S.itab + ly.perm_off : nullptr;
      if (ly.lu_bias >= 0) tr_lu_dense(S.Wlu, S.stage, Bf.theta_p, ly, D, true);
      tr_prefetch_unit(Bf, ly, S, 0, 0, 0, true, P.max_dim);
      if (l == 0) {
        tr_load_x(bt, tile, D, lperm, S.H1, S.c);
        if (threadIdx.x < TR_R) S.ld[threadIdx.x] = 0.f;
      } else {
        const TrLayer& lp = P.layer[l - 1];
        if (lp.bn_uw >= 0) {
          tr_bn_setup_eval(Bf, lp, D, S.bn);
          tr_bn_apply_perm(S.Y, S.bn, D, lperm, S.H1, nullptr);
        } else {
          for (int e = threadIdx.x; e < D * TR_R; e += TR_THREADS) {
            const int j = e / TR_R, r = e - j * TR_R;
            S.H1[e] = S.Y[(lperm ? lperm[j] : j) * TR_R + r];
          }
        }
      }
      __syncthreads();
      tr_layer_forward(Bf, P, ly, S, nullptr);
    }
    const TrLayer& ll = P.layer[L - 1];
    if (ll.bn_uw >= 0) {
      tr_bn_setup_eval(Bf, ll, D, S.bn);
      tr_bn_apply_perm(S.Y, S.bn, D, nullptr, S.H1, nullptr);
    } else {
      for (int e = threadIdx.x; e < D * TR_R; e += TR_THREADS) S.H1[e] = S.Y[e];
    }
    __syncthreads();
    if (threadIdx.x < TR_R) {
      float q = 0.f;
      for (int d = 0; d < D; ++d) q += S.H1[d * TR_R + threadIdx.x] * S.H1[d * TR_R + threadIdx.x];
      const float logp = -0.5f * binv * q - blogz + S.ld[threadIdx.x] + cld;
      const int row = tile * TR_R + threadIdx.x;
      if (row < bt.B) {
        loss += S.c[threadIdx.x] * logp;
        csum += S.c[threadIdx.x];
        if (d_logp) d_logp[row] = logp;
      }
    }
    __syncthreads();
  }
  const float a = tr_block_sum(threadIdx.x < TR_R ? loss : 0.f, S.red);
  const float b = tr_block_sum(threadIdx.x < TR_R ? csum : 0.f, S.red);
  if (threadIdx.x == 0) out_part[2 * blockIdx.x] = a, out_part[2 * blockIdx.x + 1] = b;
}

__global__ void tr_eval_final_kernel(const float* part, int G, float* d_loss) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int g = 0; g < G; ++g) a += part[2 * g], b += part[2 * g + 1];
    d_loss[0] = -a / b;
  }
}

// ============================================================================ one-phase kernels
// One phase per launch: what the CPU SIMT shim of tests/_hostcheck runs (blocks one after the
// other cannot wait on each other) and the NB200_TR_CHAIN=1 debugging path of nb200_train_epoch.
__global__ void __launch_bounds__(TR_THREADS) tr_fwd_kernel(const __grid_constant__ TrPlan P, TrBuffers Bf, TrBatch bt, int l) {
  TrSmem S = tr_carve(tr_smem_dyn, P, false);
  tr_stage_itab(Bf, P, S);
  tr_fwd_phase(P, Bf, bt, l, S);
}
__global__ void __launch_bounds__(TR_THREADS) tr_loss_kernel(const __grid_constant__ TrPlan P, TrBuffers Bf, TrBatch bt) {
  TrSmem S = tr_carve(tr_smem_dyn, P, false);
  tr_stage_itab(Bf, P, S);
  tr_loss_phase(P, Bf, bt, S);
}
__global__ void __launch_bounds__(TR_THREADS) tr_bwd_kernel(const __grid_constant__ TrPlan P, TrBuffers Bf, TrBatch bt, int l) {
  TrSmem S = tr_carve(tr_smem_dyn, P, true);
  tr_stage_itab(Bf, P, S);
  tr_bwd_phase(P, Bf, bt, l, S);
}
__global__ void __launch_bounds__(TR_THREADS) tr_eval_kernel(const __grid_constant__ TrPlan P, TrBuffers Bf, TrBatch bt, float* out_part,
                                                             float* d_logp) {
  TrSmem S = tr_carve(tr_smem_dyn, P, false);
  tr_stage_itab(Bf, P, S);
  tr_eval_phase(P, Bf, bt, out_part, d_logp, S);
}

// ============================================================================ persistent kernel
// A run of epochs of FlowModel.train (flowmodel/base.py:620-680) in ONE cooperative launch: every
// optimisation step of every epoch (phases separated by grid barriers exactly where a grid-wide
// reduction forces one: the BatchNorm statistics, the loss normaliser, the gradient partials, the
// gradient norm, the updated parameters), the validation loss, and the loop control of the
// reference -- best validation loss, best-weights snapshot, patience -- decided identically by
// every CTA from the same reduced numbers.  The host reads TrCtl / hist once per launch.
constexpr int TR_MAX_CHUNK = 64;  // epochs per launch

struct TrCtl {
  float best_val;   // best validation loss so far (+inf at the start of train())
  int best_epoch;   // 1-based epoch that reached it (0: none)
  int epochs_done;  // epochs finished so far (1-based index of the last one)
  int stop;         // patience exceeded
};

struct TrRun {
  const float* x;       // training rows [n_rows][D]
  const float* w;       // their weights or NULL
  const int64_t* perm;  // [n_epochs][n_rows] row order of every epoch, or NULL
  int64_t n_rows;
  int batch_size;
  const float* xv;      // validation rows [n_val][D] (n_val == 0: none, the validation loss is NaN)
  const float* wv;
  int64_t n_val;
  int n_epochs, epoch0, validate, patience;
  int kind;             // TrOptim::kind
  float beta1, beta2, eps, weight_decay, clip;
  double beta1d, beta2d;
  double b1pow0, b2pow0;  // beta^step0 (the bias corrections 1 - beta^step continue from there)
  int64_t step0;        // optimiser steps taken before this launch
  float lr[TR_MAX_CHUNK];  // learning rate of every epoch of the launch
  float* m;
  float* v;
  float* step_info;     // [steps][2] = {loss, |grad|} or NULL
  float* hist;          // [n_epochs][2] = {sum of the batch losses, validation loss} or NULL
  float* loss_accum;    // scalar: += every batch loss (the caller zeroes it)
  float* eval_part;     // [2 * G]
  TrCtl* ctl;           // NULL: no loop control (single steps)
  float* best_p;        // snapshot of theta_p / theta_b at the best epoch
  float* best_b;
  int n_b;              // floats in theta_b
  unsigned* bar;        // grid barrier counter (zeroed by the host before the launch)
};

#ifndef NB200_SIMT_SHIM
// Grid barrier over the co-resident CTAs of a cooperative launch.  `target` counts arrivals since
// the launch; thread 0 publishes this CTA's writes (release add) and waits for everybody's (acquire loads).
__device__ __forceinline__ void tr_grid_arrive(unsigned* bar, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
  }
}
__device__ __forceinline__ void tr_grid_wait(unsigned* bar, const unsigned& target) {
  if (threadIdx.x == 0) {
    unsigned seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
    } while ((int)(seen - target) < 0);
  }
  __syncthreads();
}
__device__ __forceinline__ void tr_grid_sync(unsigned* bar, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    // release: cumulative over the CTA's writes ordered before it by the bar.sync above;
    // acquire + the bar.sync below order every thread's later reads after the other CTAs' writes
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    unsigned seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
    } while ((int)(seen - target) < 0);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(TR_THREADS) tr_train_kernel(const __grid_constant__ TrPlan P, TrBuffers Bf,
                                                              const __grid_constant__ TrRun R) {
  TrSmem S = tr_carve(tr_smem_dyn, P, true);
  tr_stage_itab(Bf, P, S);
  const int L = P.L;
  const bool lead = blockIdx.x == 0 && threadIdx.x == 0;
  unsigned target = 0;
  tr_mark(Bf, 0);
  if (Bf.pmask) {
    for (int i = blockIdx.x * TR_THREADS + threadIdx.x; i < P.n_params; i += gridDim.x * TR_THREADS)
      Bf.theta_p[i] *= Bf.pmask[i];
    tr_grid_sync(R.bar, target);
  }
  float best_val = R.ctl ? R.ctl->best_val : INFINITY;
  int best_epoch = R.ctl ? R.ctl->best_epoch : 0;
  int64_t step = R.step0;
  double b1pow = R.b1pow0, b2pow = R.b2pow0;
  int stop = 0, epoch = R.epoch0;
  for (int e = 0; e < R.n_epochs && !stop; ++e) {
    TrBatch bt;
    bt.x = R.x;
    bt.w = R.w;
    bt.perm = R.perm ? R.perm + (size_t)e * R.n_rows : nullptr;
    for (int64_t i0 = 0; i0 < R.n_rows; i0 += R.batch_size) {
      bt.i0 = i0;
      bt.B = (int)min((int64_t)R.batch_size, R.n_rows - i0);
      bt.n_tiles = (bt.B + TR_R - 1) / TR_R;
      // Between two phases the barrier is split: arrive, then the next phase's prologue (dense LU
      // matrix, BatchNorm affine constants: nothing another CTA is still producing), then wait.
      for (int l = 0; l < L; ++l) {
        // (FWD(0) follows the optimiser step: its parameters only exist after that barrier)
        tr_fwd_phase(P, Bf, bt, l, S, l == 0 ? TR_WHOLE : TR_MAIN);
        // the next phase needs layer l's batch statistics (or, before the loss, the weight sum)
        const bool grid = l == L - 1 || P.layer[l].bn_uw >= 0;
        if (grid) tr_grid_arrive(R.bar, target);
        else __syncthreads();
        if (l + 1 < L) tr_fwd_phase(P, Bf, bt, l + 1, S, TR_PROLOGUE);
        else tr_loss_phase(P, Bf, bt, S, TR_PROLOGUE);
        if (grid) tr_grid_wait(R.bar, target);
        else __syncthreads();
        tr_mark(Bf, 100 + l);
      }
      tr_loss_phase(P, Bf, bt, S, TR_MAIN);
      for (int l = L - 1; l >= 0; --l) {
        // BatchNorm backward of layer l needs its sums over the batch (from the loss / BWD(l + 1))
        const bool grid = P.layer[l].bn_uw >= 0;
        if (grid) tr_grid_arrive(R.bar, target);
        else __syncthreads();
        // (after the loss phase the last layer's BatchNorm constants are still in shared memory)
        tr_bwd_phase(P, Bf, bt, l, S, TR_PROLOGUE | (l == L - 1 ? TR_KEEP_BN : 0));
        if (grid) tr_grid_wait(R.bar, target);
        else __syncthreads();
        tr_mark(Bf, l == L - 1 ? 200 : 301 + l);
        tr_bwd_phase(P, Bf, bt, l, S, TR_MAIN);
      }
      tr_grid_sync(R.bar, target);  // REDUCE reads every CTA's partial vector
      tr_mark(Bf, 300);
      tr_reduce_phase(P, Bf, reinterpret_cast<float4*>(tr_smem_dyn), S.red);
      tr_grid_sync(R.bar, target);
      tr_mark(Bf, 400);
      ++step;
      b1pow *= R.beta1d, b2pow *= R.beta2d;  // beta^step
      TrOptim o;
      o.kind = R.kind, o.lr = R.lr[e], o.beta1 = R.beta1, o.beta2 = R.beta2, o.eps = R.eps;
      o.weight_decay = R.weight_decay, o.clip = R.clip;
      o.bc1 = (float)(1.0 - b1pow);
      o.bc2 = (float)(1.0 - b2pow);
      tr_adam_phase<TR_THREADS>(Bf, P.n_params, o, R.m, R.v,
                                R.step_info ? R.step_info + 2 * (step - R.step0 - 1) : nullptr, R.loss_accum,
                                S.red);
      tr_grid_sync(R.bar, target);
      tr_mark(Bf, 500);
    }
    ++epoch;
    if (!R.ctl) continue;
    // validation loss (eval mode: running statistics), the same number in every CTA
    float val = NAN;
    if (R.n_val > 0) {
      TrBatch bv;
      bv.x = R.xv, bv.w = R.wv, bv.perm = nullptr, bv.i0 = 0, bv.B = (int)R.n_val;
      bv.n_tiles = (bv.B + TR_R - 1) / TR_R;
      tr_eval_phase(P, Bf, bv, R.eval_part, nullptr, S);
      tr_grid_sync(R.bar, target);
      float a = 0.f, b = 0.f;
      for (int g = 0; g < (int)gridDim.x; ++g) a += R.eval_part[2 * g], b += R.eval_part[2 * g + 1];
      val = -a / b;
    }
    if (lead && R.hist) {
      R.hist[2 * e] = R.loss_accum[0];
      R.hist[2 * e + 1] = val;
      R.loss_accum[0] = 0.f;
    }
    if (R.validate) {
      if (val < best_val) {
        best_val = val;
        best_epoch = epoch;
        for (int i = blockIdx.x * TR_THREADS + threadIdx.x; i < P.n_params; i += gridDim.x * TR_THREADS)
          R.best_p[i] = Bf.theta_p[i];
        for (int i = blockIdx.x * TR_THREADS + threadIdx.x; i < R.n_b; i += gridDim.x * TR_THREADS)
          R.best_b[i] = Bf.theta_b[i];
      }
      if (epoch - best_epoch > R.patience) stop = 1;
    }
    tr_mark(Bf, 600);
  }
  if (lead && R.ctl) {
    R.ctl->best_val = best_val;
    R.ctl->best_epoch = best_epoch;
    R.ctl->epochs_done = epoch;
    R.ctl->stop = stop;
  }
}
#endif  // NB200_SIMT_SHIM

}  // namespace nb200
