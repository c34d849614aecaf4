// Generic fp32 interpreter of a folded flow program: one thread per row.
//
// Replaces, for the eval-mode flow, the per-call Module walk of the reference
// (/root/reference/src/nessai/flows/base.py:209-221 NFlow.forward/inverse ->
//  nflows CompositeTransform / CouplingTransform / LULinear / BatchNorm and the
//  conditioner nets /root/reference/src/nessai/flows/nets.py:83-126).
//
// Layout: every per-row vector lives in shared memory as a column:
//   buf[k * BS + tid]   (BS = blockDim.x)  -> conflict-free, private to the thread.
// A linear op stages its weights (k-major [K][Npad] + bias[Npad]) into shared
// memory once per CTA; every thread then reads them as warp-wide broadcasts
// (LDS.128) while accumulating CH outputs in registers.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "flow_program.h"

namespace nb200 {

template <int ACT>
__device__ __forceinline__ float actf(float v) {
  if (ACT == ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == ACT_TANH) return tanhf(v);
  return v / (1.f + expf(-v));  // silu: x * sigmoid(x)
}

// cooperative copy of one op's weights + bias into shared memory
__device__ __forceinline__ void stage_weights(float* Ws, const float* __restrict__ blob,
                                              const FlowOp& op) {
  const int nw = op.K * op.Npad;  // multiple of 4 (Npad % 8 == 0)
  const float4* gw = reinterpret_cast<const float4*>(blob + op.w_off);
  float4* sw = reinterpret_cast<float4*>(Ws);
  for (int i = threadIdx.x; i < nw / 4; i += blockDim.x) sw[i] = __ldg(gw + i);
  const float4* gb = reinterpret_cast<const float4*>(blob + op.b_off);
  float4* sb = reinterpret_cast<float4*>(Ws + nw);
  for (int i = threadIdx.x; i < op.Npad / 4; i += blockDim.x) sb[i] = __ldg(gb + i);
}

template <int ACT, int CH>
__device__ __forceinline__ void linear_chunk(const float* Ws, const float* bs, const float* src,
                                             float* dst, int K, int n0, int Npad, int BS,
                                             int flags) {
  float acc[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) acc[j] = bs[n0 + j];
  if (flags & FLAG_ACCUM) {
#pragma unroll
    for (int j = 0; j < CH; ++j) acc[j] += dst[(n0 + j) * BS];
  }
  const bool in_act = flags & FLAG_IN_ACT;
  for (int k = 0; k < K; ++k) {
    float a = src[k * BS];
    if (in_act) a = actf<ACT>(a);
    const float4* w4 = reinterpret_cast<const float4*>(Ws + k * Npad + n0);
#pragma unroll
    for (int j = 0; j < CH / 4; ++j) {
      const float4 w = w4[j];
      acc[4 * j + 0] = fmaf(a, w.x, acc[4 * j + 0]);
      acc[4 * j + 1] = fmaf(a, w.y, acc[4 * j + 1]);
      acc[4 * j + 2] = fmaf(a, w.z, acc[4 * j + 2]);
      acc[4 * j + 3] = fmaf(a, w.w, acc[4 * j + 3]);
    }
  }
  const bool out_act = flags & FLAG_OUT_ACT;
#pragma unroll
  for (int j = 0; j < CH; ++j) {
    float v = acc[j];
    if (out_act) v = actf<ACT>(v);
    dst[(n0 + j) * BS] = v;
  }
}

template <int ACT>
__device__ __forceinline__ void op_linear(const FlowOp& op, const float* Ws, float* const* bufs,
                                          int BS) {
  const float* bs = Ws + op.K * op.Npad;
  const float* src = bufs[op.src] + op.src_off * BS;
  float* dst = bufs[op.dst];
  int n0 = 0;
  for (; n0 + 32 <= op.Npad; n0 += 32)
    linear_chunk<ACT, 32>(Ws, bs, src, dst, op.K, n0, op.Npad, BS, op.flags);
  for (; n0 + 8 <= op.Npad; n0 += 8)
    linear_chunk<ACT, 8>(Ws, bs, src, dst, op.K, n0, op.Npad, BS, op.flags);
}

// final conditioner layer fused with the affine coupling
// (nflows AffineCouplingTransform: scale = sigmoid(u + 2) + 1e-3)
__device__ __forceinline__ void op_coupling_affine(const FlowOp& op, const float* Ws,
                                                   float* const* bufs, int BS, float& ld) {
  const float* bs = Ws + op.K * op.Npad;
  const float* src = bufs[op.src];
  const float* x = bufs[op.x_buf] + op.d_id * BS;  // values the affine acts on
  float* xo = bufs[op.dst] + op.d_id * BS;         // == x except for the MAF inverse passes
  const bool inverse = op.flags & FLAG_INVERSE;
  const bool additive = op.flags & FLAG_ADDITIVE;
  const bool softplus_scale = op.flags & FLAG_SOFTPLUS_SCALE;
  const bool count_ld = !(op.flags & FLAG_NO_LOGDET);
  for (int n0 = 0; n0 < op.Npad; n0 += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bs[n0 + j];
    for (int k = 0; k < op.K; ++k) {
      const float a = src[k * BS];
      const float4* w4 = reinterpret_cast<const float4*>(Ws + k * op.Npad + n0);
      const float4 w0 = w4[0], w1 = w4[1];
      acc[0] = fmaf(a, w0.x, acc[0]);
      acc[1] = fmaf(a, w0.y, acc[1]);
      acc[2] = fmaf(a, w0.z, acc[2]);
      acc[3] = fmaf(a, w0.w, acc[3]);
      acc[4] = fmaf(a, w1.x, acc[4]);
      acc[5] = fmaf(a, w1.y, acc[5]);
      acc[6] = fmaf(a, w1.z, acc[6]);
      acc[7] = fmaf(a, w1.w, acc[7]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int f = n0 / 2 + i;
      if (f < op.d_tr) {
        const float t = acc[2 * i];
        float s = 1.f, ls = 0.f;
        if (!additive) {
          const float u = acc[2 * i + 1];
          s = softplus_scale ? (u > 20.f ? u : log1pf(expf(u))) + 1e-3f
                             : 1.f / (1.f + expf(-(u + 2.f))) + 1e-3f;
          ls = count_ld ? logf(s) : 0.f;
        }
        const float xi = x[f * BS];
        if (inverse) {
          xo[f * BS] = (xi - t) / s;
          ld -= ls;
        } else {
          xo[f * BS] = fmaf(xi, s, t);
          ld += ls;
        }
      }
    }
  }
}

// --- rational-quadratic spline coupling (nflows PiecewiseRationalQuadratic
//     CouplingTransform, linear tails) ------------------------------
__device__ __forceinline__ float softplusf(float v) {
  return v > 20.f ? v : log1pf(expf(v));
}

template <int MAXNB>
__device__ __forceinline__ void spline_feature(const float* p, int nb, float B, bool inverse,
                                               float& xio, float& ld) {
  const float x = xio;
  if (!(x >= -B && x <= B)) return;  // linear tails: identity, logabsdet 0 (NaN falls through)
  const float min_w = 1e-3f, min_h = 1e-3f, min_d = 1e-3f;
  float cw[MAXNB + 1], ch[MAXNB + 1], d[MAXNB + 1];
  // softmax widths / heights
  float mw = -INFINITY, mh = -INFINITY;
#pragma unroll
  for (int k = 0; k < MAXNB; ++k)
    if (k < nb) {
      mw = fmaxf(mw, p[k]);
      mh = fmaxf(mh, p[nb + k]);
    }
  float sw = 0.f, sh = 0.f;
  float ew[MAXNB], eh[MAXNB];
#pragma unroll
  for (int k = 0; k < MAXNB; ++k)
    if (k < nb) {
      ew[k] = expf(p[k] - mw);
      eh[k] = expf(p[nb + k] - mh);
      sw += ew[k];
      sh += eh[k];
    }
  float aw = 0.f, ah = 0.f;
  cw[0] = -B;
  ch[0] = -B;
#pragma unroll
  for (int k = 0; k < MAXNB; ++k)
    if (k < nb) {
      aw += min_w + (1.f - min_w * nb) * (ew[k] / sw);
      ah += min_h + (1.f - min_h * nb) * (eh[k] / sh);
      cw[k + 1] = (k == nb - 1) ? B : fmaf(2.f * B, aw, -B);
      ch[k + 1] = (k == nb - 1) ? B : fmaf(2.f * B, ah, -B);
    }
  // derivatives: edges are min_d + softplus(log(exp(1 - min_d) - 1)) == 1
  d[0] = 1.f;
#pragma unroll
  for (int k = 1; k < MAXNB; ++k)
    if (k < nb) d[k] = min_d + softplusf(p[2 * nb + k - 1]);
#pragma unroll
  for (int k = 0; k <= MAXNB; ++k)
    if (k == nb) d[k] = 1.f;
  // bin search (searchsorted with the last knot nudged by 1e-6)
  int b = -1;
#pragma unroll
  for (int k = 0; k <= MAXNB; ++k)
    if (k <= nb) {
      float knot = inverse ? ch[k] : cw[k];
      if (k == nb) knot += 1e-6f;
      b += (x >= knot) ? 1 : 0;
    }
  b = min(max(b, 0), nb - 1);
  float icw = 0, iw = 0, ich = 0, ih = 0, d0 = 0, d1 = 0;
#pragma unroll
  for (int k = 0; k < MAXNB; ++k)
    if (k == b) {
      icw = cw[k];
      iw = cw[k + 1] - cw[k];
      ich = ch[k];
      ih = ch[k + 1] - ch[k];
      d0 = d[k];
      d1 = d[k + 1];
    }
  const float delta = ih / iw;
  if (inverse) {
    const float dy = x - ich;
    const float q = d0 + d1 - 2.f * delta;
    const float a = dy * q + ih * (delta - d0);
    const float bb = ih * d0 - dy * q;
    const float c = -delta * dy;
    const float disc = bb * bb - 4.f * a * c;  // < 0 -> NaN row (reference asserts)
    const float root = (2.f * c) / (-bb - sqrtf(disc));
    xio = fmaf(root, iw, icw);
    const float t1m = root * (1.f - root);
    const float den = delta + q * t1m;
    const float dnum = delta * delta * (d1 * root * root + 2.f * delta * t1m + d0 * (1.f - root) * (1.f - root));
    ld -= logf(dnum) - 2.f * logf(den);
  } else {
    const float th = (x - icw) / iw;
    const float t1m = th * (1.f - th);
    const float q = d0 + d1 - 2.f * delta;
    const float num = ih * (delta * th * th + d0 * t1m);
    const float den = delta + q * t1m;
    xio = ich + num / den;
    const float dnum = delta * delta * (d1 * th * th + 2.f * delta * t1m + d0 * (1.f - th) * (1.f - th));
    ld += logf(dnum) - 2.f * logf(den);
  }
}

template <int MAXNB>
__device__ __forceinline__ void op_coupling_spline(const FlowOp& op, const float* Ws,
                                                   float* const* bufs, int BS, float& ld) {
  constexpr int MAXG = (3 * MAXNB - 1 + 7) / 8 * 8;
  const float* bs = Ws + op.K * op.Npad;
  const float* src = bufs[op.src];
  float* x = bufs[op.x_buf] + op.d_id * BS;
  const int nb = op.e0, G = op.e1;
  const float B = __int_as_float(op.e2);
  const bool inverse = op.flags & FLAG_INVERSE;
  for (int f = 0; f < op.d_tr; ++f) {
    float acc[MAXG];
    const int n0 = f * G;
#pragma unroll
    for (int j = 0; j < MAXG; ++j) acc[j] = (j < G) ? bs[n0 + j] : 0.f;
    for (int k = 0; k < op.K; ++k) {
      const float a = src[k * BS];
      const float4* w4 = reinterpret_cast<const float4*>(Ws + k * op.Npad + n0);
#pragma unroll
      for (int j = 0; j < MAXG / 4; ++j)
        if (4 * j < G) {
          const float4 w = w4[j];
          acc[4 * j + 0] = fmaf(a, w.x, acc[4 * j + 0]);
          acc[4 * j + 1] = fmaf(a, w.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(a, w.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(a, w.w, acc[4 * j + 3]);
        }
    }
    float xi = x[f * BS];
    spline_feature<MAXNB>(acc, nb, B, inverse, xi, ld);
    x[f * BS] = xi;
  }
}

// Run the whole program on the rows held in X0; returns the row's log|det J|
// (without the constant).  Contains __syncthreads(): every thread of the CTA
// must call it.
template <int ACT>
__device__ __forceinline__ float run_program(const FlowProgramDev& P, float* Ws,
                                             float* const* bufs, int BS) {
  float ld = 0.f;
  for (int i = 0; i < P.n_ops; ++i) {
    const FlowOp op = P.ops[i];
    __syncthreads();  // previous op done with Ws
    stage_weights(Ws, P.blob, op);
    __syncthreads();
    if (op.type == OP_LINEAR) {
      op_linear<ACT>(op, Ws, bufs, BS);
    } else if (op.type == OP_COUPLING_AFFINE) {
      op_coupling_affine(op, Ws, bufs, BS, ld);
    } else {
      if (op.e0 <= 8)
        op_coupling_spline<8>(op, Ws, bufs, BS, ld);
      else
        op_coupling_spline<16>(op, Ws, bufs, BS, ld);
    }
  }
  return ld;
}

__device__ __forceinline__ void carve_buffers(float* smem, const FlowProgramDev& P, int BS,
                                              float*& Ws, float** bufs) {
  Ws = smem;
  float* p = smem + ((P.wmax + 3) & ~3);
  bufs[BUF_X0] = p + threadIdx.x;
  p += P.Dpad * BS;
  bufs[BUF_X1] = p + threadIdx.x;
  p += P.Dpad * BS;
  bufs[BUF_A0] = p + threadIdx.x;
  p += P.Hpad * BS;
  bufs[BUF_A1] = p + threadIdx.x;
}

inline size_t interp_smem_bytes(const FlowProgramDev& P, int BS) {
  return sizeof(float) * (size_t)(((P.wmax + 3) & ~3) + 2 * P.Dpad * BS + 2 * P.Hpad * BS);
}

}  // namespace nb200
