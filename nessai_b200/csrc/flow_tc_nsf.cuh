// tcgen05 / TMEM kernel for nessai's neural spline flow (config C3):
// /root/reference/src/nessai/flows/nsf.py:60-130 -- per layer a permutation, then nflows'
// PiecewiseRationalQuadraticCouplingTransform (8 bins, linear tails) conditioned by a
// ResidualNet(64) whose final layer emits 23 spline parameters per transformed feature.
//
// Same machinery as flow_tc_res.cuh (row == TMEM lane, split-bf16 3-pass MMAs with the A
// operand in tensor memory, residual stream accumulating in TMEM, converged issuer warps, eight
// epilogue warps per 128-row tile, layer passes through an L2 scratch), plus:
//   * the ROW STATE (up to 32 features) lives in TENSOR MEMORY (32 columns per tile) instead of
//     registers: both epilogue warps of a lane quarter can reach it, so the splines of two
//     features run in parallel per row, and a permutation between layers is free -- the state
//     never moves, each layer just addresses different columns (the permutations are composed on
//     the host; the first conditioner GEMM gets its weight columns scattered to physical slots);
//   * the wide final layer (d_tr x 24 outputs) runs in chunks of 48 columns = 2 features, each
//     followed by the rational-quadratic spline epilogue of those two features (raw SFU ops,
//     branch-free bin selection).
// Covered: D <= 32, width <= 64, ReLU, 1-2 residual blocks, 8 bins, permutation / no linear
// transform, no BatchNorm between layers (nessai's NSF defaults).  Everything else runs the
// generic kernel.
//
// AC mode (template parameter): the same tile machinery for RealNVP's AFFINE coupling with the
// default ResidualNet conditioner at 17 .. 32 features (flows/realnvp.py:76-214; up to 16 features
// flow_tc_res.cuh keeps the state in registers).  Differences: the dense D x D affine in front of
// every coupling (LU x permutation x BatchNorm, folded) rides the first conditioner GEMM as 32
// extra output columns (N = 96: hidden pre-activations | the state after the affine), which the
// epilogue copies into the state columns; identity features live in slots 0 .. 15 and transformed
// ones in 16 .. 31 of every layer (the affines are re-laid-out on the host); the final layer is ONE
// chunk (shift | unconstrained scale) followed by the coupling on the transformed slots, eight per
// twin warp; the affine after the last coupling runs on the CUDA cores of the output stage.
#pragma once
#include "flow_tc_res.cuh"

namespace nb200 {

constexpr int NS_DP = 32;                 // state slots
constexpr int NS_K = 8;                   // spline bins
constexpr int NS_G = 24;                  // padded parameters per feature (3K - 1 = 23)
constexpr int NS_CN = 2 * NS_G;           // columns of one final-layer chunk (2 features)
constexpr int NS_MAXCH = NS_DP / 2;
constexpr int NS_COLS = 256;              // TMEM columns per tile
constexpr int NS_COL_D = 0, NS_COL_D2 = 64, NS_COL_AH = 128, NS_COL_AL = 160, NS_COL_ST = 192, NS_COL_LD = 224;
constexpr int NS_W0 = TC_H * 16 * (NS_DP / 8);        // 64 rows, K = 32: 4 KB
constexpr int NS_WF = NS_CN * 16 * (TC_H / 8);        // 48 rows, K = 64: 6 KB
constexpr int NS_BF = NS_CN * 16;                     // bias operand of a chunk: 768 B

constexpr int AC_ROWS0 = TC_H + NS_DP;    // AC mode: rows of the first GEMM's B operand (hidden | affine)
constexpr int AC_TR0 = 16;                // AC mode: first state slot of the transformed features
constexpr int AC_ZERO_BYTES = AC_ROWS0 * 16;  // the bias operands' all-zero K-chunk must cover 96 rows
constexpr int AC_AFF_BYTES = (NS_DP * NS_DP + NS_DP) * 4;  // fp32 affine after the last coupling, k-major + bias

struct NsLayout {
  int w0hi, w0lo, blk, wf, b0, bblk, bf, aff, layer_bytes;
};
__host__ __device__ inline NsLayout ns_layout(int NB, int nch, bool ac = false) {
  NsLayout o;
  const int rows0 = ac ? AC_ROWS0 : TC_H;
  const int w0 = rows0 * 16 * (NS_DP / 8);
  o.w0hi = 0;
  o.w0lo = w0;
  o.blk = 2 * w0;
  o.wf = o.blk + NB * 4 * RS_W_BIG;      // per chunk: hi, lo
  o.b0 = o.wf + nch * 2 * NS_WF;
  o.bblk = o.b0 + rows0 * 16;
  o.bf = o.bblk + NB * 2 * RS_BIAS;
  o.aff = o.bf + nch * NS_BF;
  o.layer_bytes = o.aff + (ac ? AC_AFF_BYTES : 0);
  return o;
}

struct NsLayerInfo {
  int d_tr;
  int8_t trslot[NS_DP];  // physical state slot of transformed feature i
};
struct NsProgram {
  bool valid = false;
  int narrow = 0;  // conditioner width <= 32: two K-steps, 32-column hidden epilogues (flow_tc.cuh)
  int ac = 0;      // RealNVP affine coupling (AC mode) instead of the spline coupling
  int additive = 0;
  int L = 0, D = 0, NB = 0, nch = 0, inverse = 0;
  float tail_bound = 0.f, const_logdet = 0.f;
  NsLayerInfo layer[TC_MAXL];
  int8_t out_slot[NS_DP];  // output feature j = state slot out_slot[j]
  uint8_t* d_image[TC_MAXL] = {nullptr};
  int image_bytes = 0;
  float* d_scratch = nullptr;  // [rows][32] state | [rows] log|det| | [rows] sum z^2 (sign: alive)
  int64_t scratch_rows = 0;
};
inline void ns_free(NsProgram& t) {
  for (int l = 0; l < TC_MAXL; ++l)
    if (t.d_image[l]) cudaFree(t.d_image[l]);
  if (t.d_scratch) cudaFree(t.d_scratch);
  t = NsProgram();
}

struct NsParams {
  const uint8_t* image;
  int image_bytes;
  int D, NB, nch, inverse, first, last;
  int additive;
  NsLayerInfo ly;
  int8_t out_slot[NS_DP];
  float tail_bound, const_logdet;
  float* sc_h;
  float* sc_ld;
  float* sc_ss;
};

// Is `f` (k-major [K][Npad] + bias) a pure permutation?  perm[n] = source index of output n.
inline bool ns_perm_of(const FlowOp& f, const float* blob, int D, int* perm) {
  if (f.type != OP_LINEAR || f.K != D || f.N != D || f.flags != 0 || f.src_off != 0) return false;
  for (int n = 0; n < D; ++n) {
    if (blob[f.b_off + n] != 0.f) return false;
    int src = -1;
    for (int k = 0; k < D; ++k) {
      const float w = blob[f.w_off + k * f.Npad + n];
      if (w == 1.f && src < 0) src = k;
      else if (w != 0.f) return false;
    }
    if (src < 0) return false;
    perm[n] = src;
  }
  return true;
}

inline int ns_build(NsProgram& t, const FlowOp* ops, int n_ops, const float* blob, int D, int H,
                    int activation) {
  t.valid = false;
  if (D > NS_DP || D < 2 || H < 1 || H > TC_H || activation != ACT_RELU || n_ops < 6) return 0;
  int NB = -1;
  for (int nb = 1; nb <= RS_MAXNB; ++nb)
    if ((n_ops - 1) % (2 * nb + 3) == 0 && ops[2 * nb + 2].type == OP_COUPLING_SPLINE) NB = nb;
  if (NB < 0) return 0;
  const int per = 2 * NB + 3;
  const int L = (n_ops - 1) / per;
  if (L < 1 || L > TC_MAXL) return 0;
  int m[NS_DP], pm[NS_DP], nm[NS_DP];  // h_l[j] = state[m[j]]
  if (!ns_perm_of(ops[0], blob, D, pm)) return 0;
  for (int j = 0; j < D; ++j) m[j] = pm[j];
  int inverse = -1, nch = 0;
  float B = 0.f;
  tc_put_overflow() = false;
  for (int l = 0; l < L; ++l) {
    const FlowOp* o = ops + 1 + per * l;
    const FlowOp& a = o[0];
    if (a.type != OP_LINEAR || a.src > BUF_X1 || a.dst < BUF_A0 || a.N != H || a.flags != 0 ||
        a.src_off != 0 || a.K < 1 || a.K > D)
      return 0;
    for (int b = 0; b < NB; ++b) {
      const FlowOp& x = o[1 + 2 * b];
      const FlowOp& y = o[2 + 2 * b];
      if (x.type != OP_LINEAR || x.src != a.dst || x.dst == a.dst || x.dst < BUF_A0 || x.K != H ||
          x.N != H || x.flags != (FLAG_IN_ACT | FLAG_OUT_ACT))
        return 0;
      if (y.type != OP_LINEAR || y.src != x.dst || y.dst != a.dst || y.K != H || y.N != H ||
          y.flags != FLAG_ACCUM)
        return 0;
    }
    const FlowOp& c = o[1 + 2 * NB];
    if (c.type != OP_COUPLING_SPLINE || c.src != a.dst || c.K != H || c.d_id != a.K || c.d_tr < 1 ||
        c.d_id + c.d_tr != D || c.e0 != NS_K || c.e1 != NS_G || c.N != c.d_tr * NS_G || c.x_buf != c.dst)
      return 0;
    if ((c.flags & ~FLAG_INVERSE) != 0) return 0;
    const int inv = (c.flags & FLAG_INVERSE) ? 1 : 0;
    if (inverse >= 0 && inverse != inv) return 0;
    inverse = inv;
    float Bl;
    memcpy(&Bl, &c.e2, 4);
    if (l > 0 && Bl != B) return 0;
    B = Bl;
    t.layer[l].d_tr = c.d_tr;
    for (int i = 0; i < c.d_tr; ++i) t.layer[l].trslot[i] = (int8_t)m[c.d_id + i];
    nch = std::max(nch, (c.d_tr + 1) / 2);
    // the affine after this coupling
    if (!ns_perm_of(o[2 + 2 * NB], blob, D, pm)) return 0;
    for (int j = 0; j < D; ++j) nm[j] = m[pm[j]];
    // (image of this layer is built below, with the map m of THIS layer)
    const NsLayout lay = ns_layout(NB, (c.d_tr + 1) / 2);
    (void)lay;
    for (int j = 0; j < D; ++j) pm[j] = m[j];  // keep this layer's map for the image
    // build image
    {
      const int nchl = (c.d_tr + 1) / 2;
      const NsLayout ly = ns_layout(NB, nchl);
      const int ones_off = ly.layer_bytes;
      const int bytes = ones_off + TC_ONES_BYTES + TC_ZERO_BYTES;
      if (((bytes + 1023) & ~1023) + 4096 > 227 * 1024) return 0;
      std::vector<uint8_t> img((size_t)bytes, 0);
      for (int r = 0; r < 128; ++r) {
        const uint16_t one[2] = {TC_ONE16, TC_ONE16};
        memcpy(img.data() + ones_off + (size_t)r * 16, one, 4);
      }
      auto put_bias = [&](uint8_t* base, int n, float bv) {
        if (!tc_h16_representable(bv)) tc_put_overflow() = true;
        const uint16_t hi = tc_h16_rn(bv);
        const uint16_t lo = tc_h16_rn(bv - tc_h16_to_f(hi));
        memcpy(base + (size_t)n * 16, &hi, 2);
        memcpy(base + (size_t)n * 16 + 2, &lo, 2);
      };
      uint8_t* lb = img.data();
      // initial layer: identity feature i sits in physical slot m[i]
      for (int n = 0; n < H; ++n) {
        for (int i = 0; i < a.K; ++i)
          tc_put(lb + ly.w0hi, lb + ly.w0lo, TC_H, n, pm[i], blob[a.w_off + i * a.Npad + n]);
        put_bias(lb + ly.b0, n, blob[a.b_off + n]);
      }
      for (int b = 0; b < NB; ++b) {
        uint8_t* wb = lb + ly.blk + (size_t)b * 4 * RS_W_BIG;
        const FlowOp& x = o[1 + 2 * b];
        const FlowOp& y = o[2 + 2 * b];
        for (int n = 0; n < H; ++n)
          for (int k = 0; k < H; ++k) {
            tc_put(wb, wb + RS_W_BIG, TC_H, n, k, blob[x.w_off + k * x.Npad + n]);
            tc_put(wb + 2 * RS_W_BIG, wb + 3 * RS_W_BIG, TC_H, n, k, blob[y.w_off + k * y.Npad + n]);
          }
        for (int n = 0; n < H; ++n) {
          put_bias(lb + ly.bblk + (size_t)(2 * b) * RS_BIAS, n, blob[x.b_off + n]);
          put_bias(lb + ly.bblk + (size_t)(2 * b + 1) * RS_BIAS, n, blob[y.b_off + n]);
        }
      }
      for (int j = 0; j < nchl; ++j) {
        uint8_t* wh = lb + ly.wf + (size_t)j * 2 * NS_WF;
        for (int n = 0; n < NS_CN; ++n) {
          const int col = j * NS_CN + n;  // column of the program's final layer
          if (col >= c.N) continue;
          for (int k = 0; k < H; ++k) tc_put(wh, wh + NS_WF, NS_CN, n, k, blob[c.w_off + k * c.Npad + col]);
          put_bias(lb + ly.bf + (size_t)j * NS_BF, n, blob[c.b_off + col]);
        }
      }
      if (tc_put_overflow()) return 0;
      if (cudaMalloc(&t.d_image[l], bytes) != cudaSuccess) return 2;
      if (cudaMemcpy(t.d_image[l], img.data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess) return 2;
      t.image_bytes = std::max(t.image_bytes, bytes);
    }
    for (int j = 0; j < D; ++j) m[j] = nm[j];
  }
  for (int l = 0; l < L; ++l)
    if ((t.layer[l].d_tr + 1) / 2 != nch) return 0;  // one chunk count for every layer
  for (int j = 0; j < D; ++j) t.out_slot[j] = (int8_t)m[j];
  t.L = L;
  t.D = D;
  t.NB = NB;
  t.nch = nch;
  t.inverse = inverse;
  t.tail_bound = B;
  t.narrow = H <= TC_H / 2;
  t.valid = true;
  return 0;
}

// AC mode: recognise  affine (resnet-coupling affine)*  as rs_build does, for 17 .. 32 features (up to
// 16 the register-state kernel of flow_tc_res.cuh is faster), and build one image per layer.
inline int ac_build(NsProgram& t, const FlowOp* ops, int n_ops, const float* blob, int D, int H, int activation) {
  t.valid = false;
  if (D > NS_DP || D <= TC_DP || H < 1 || H > TC_H || activation != ACT_RELU || n_ops < 6) return 0;
  int NB = -1;
  for (int nb = 1; nb <= RS_MAXNB; ++nb)
    if ((n_ops - 1) % (2 * nb + 3) == 0 && ops[2].type == OP_LINEAR && ops[2 * nb + 2].type == OP_COUPLING_AFFINE)
      NB = nb;
  if (NB < 0) return 0;
  const int per = 2 * NB + 3;
  const int L = (n_ops - 1) / per;
  if (L < 1 || L > TC_MAXL) return 0;
  auto is_affine = [&](const FlowOp& o) {
    return o.type == OP_LINEAR && o.src <= BUF_X1 && o.dst <= BUF_X1 && o.K == D && o.N == D && o.flags == 0 &&
           o.src_off == 0;
  };
  if (!is_affine(ops[0])) return 0;
  int inverse = -1, additive = -1;
  int d_id[TC_MAXL], d_tr[TC_MAXL];
  for (int l = 0; l < L; ++l) {
    const FlowOp* o = ops + 1 + per * l;
    const FlowOp& a = o[0];
    if (a.type != OP_LINEAR || a.src > BUF_X1 || a.dst < BUF_A0 || a.N != H || a.flags != 0 || a.src_off != 0 ||
        a.K < 1 || a.K > AC_TR0)
      return 0;
    for (int b = 0; b < NB; ++b) {
      const FlowOp& x = o[1 + 2 * b];
      const FlowOp& y = o[2 + 2 * b];
      if (x.type != OP_LINEAR || x.src != a.dst || x.dst == a.dst || x.dst < BUF_A0 || x.K != H || x.N != H ||
          x.flags != (FLAG_IN_ACT | FLAG_OUT_ACT))
        return 0;
      if (y.type != OP_LINEAR || y.src != x.dst || y.dst != a.dst || y.K != H || y.N != H || y.flags != FLAG_ACCUM)
        return 0;
    }
    const FlowOp& c = o[1 + 2 * NB];
    if (c.type != OP_COUPLING_AFFINE || c.src != a.dst || c.K != H || c.d_id != a.K || c.d_tr < 1 ||
        c.d_tr > NS_DP - AC_TR0 || c.d_id + c.d_tr != D || c.N != 2 * c.d_tr)
      return 0;
    const int inv = (c.flags & FLAG_INVERSE) ? 1 : 0, add = (c.flags & FLAG_ADDITIVE) ? 1 : 0;
    if ((c.flags & ~(FLAG_INVERSE | FLAG_ADDITIVE)) != 0) return 0;
    if ((inverse >= 0 && inverse != inv) || (additive >= 0 && additive != add)) return 0;
    inverse = inv;
    additive = add;
    if (!is_affine(o[2 + 2 * NB])) return 0;
    d_id[l] = c.d_id;
    d_tr[l] = c.d_tr;
  }
  // state slot of natural feature j inside coupling layer `layer` (the flow's input / output: natural)
  auto slot = [&](int layer, int j) {
    if (layer < 0 || layer >= L) return j;
    return j < d_id[layer] ? j : AC_TR0 + (j - d_id[layer]);
  };
  auto put_bias = [&](uint8_t* base, int n, float bv) {
    if (!tc_h16_representable(bv)) tc_put_overflow() = true;
    const uint16_t hi = tc_h16_rn(bv);
    const uint16_t lo = tc_h16_rn(bv - tc_h16_to_f(hi));
    memcpy(base + (size_t)n * 16, &hi, 2);
    memcpy(base + (size_t)n * 16 + 2, &lo, 2);
  };
  tc_put_overflow() = false;
  const NsLayout ly = ns_layout(NB, 1, true);
  const int ones_off = ly.layer_bytes;
  const int bytes = ones_off + TC_ONES_BYTES + AC_ZERO_BYTES;
  if (((bytes + 1023) & ~1023) + 4096 > 227 * 1024) return 0;
  for (int l = 0; l < L; ++l) {
    const FlowOp& f = ops[per * l];  // the affine in front of coupling l
    const FlowOp* o = ops + 1 + per * l;
    const FlowOp& a = o[0];
    std::vector<uint8_t> img((size_t)bytes, 0);
    for (int r = 0; r < 128; ++r) {
      const uint16_t one[2] = {TC_ONE16, TC_ONE16};
      memcpy(img.data() + ones_off + (size_t)r * 16, one, 4);
    }
    uint8_t* lb = img.data();
    // first conditioner layer with the affine folded in (float64): consumes the state BEFORE the affine
    for (int n = 0; n < H; ++n) {
      for (int k = 0; k < D; ++k) {
        double acc = 0.0;
        for (int j = 0; j < a.K; ++j)
          acc += (double)blob[a.w_off + j * a.Npad + n] * (double)blob[f.w_off + k * f.Npad + j];
        tc_put(lb + ly.w0hi, lb + ly.w0lo, AC_ROWS0, n, slot(l - 1, k), (float)acc);
      }
      double bacc = blob[a.b_off + n];
      for (int j = 0; j < a.K; ++j) bacc += (double)blob[a.w_off + j * a.Npad + n] * (double)blob[f.b_off + j];
      put_bias(lb + ly.b0, n, (float)bacc);
    }
    // rows 64 .. 95: the affine itself, slots of layer l - 1 -> slots of layer l
    for (int n = 0; n < D; ++n) {
      for (int k = 0; k < D; ++k)
        tc_put(lb + ly.w0hi, lb + ly.w0lo, AC_ROWS0, TC_H + slot(l, n), slot(l - 1, k), blob[f.w_off + k * f.Npad + n]);
      put_bias(lb + ly.b0, TC_H + slot(l, n), blob[f.b_off + n]);
    }
    for (int b = 0; b < NB; ++b) {
      uint8_t* wb = lb + ly.blk + (size_t)b * 4 * RS_W_BIG;
      const FlowOp& x = o[1 + 2 * b];
      const FlowOp& y = o[2 + 2 * b];
      for (int n = 0; n < H; ++n)
        for (int k = 0; k < H; ++k) {
          tc_put(wb, wb + RS_W_BIG, TC_H, n, k, blob[x.w_off + k * x.Npad + n]);
          tc_put(wb + 2 * RS_W_BIG, wb + 3 * RS_W_BIG, TC_H, n, k, blob[y.w_off + k * y.Npad + n]);
        }
      for (int n = 0; n < H; ++n) {
        put_bias(lb + ly.bblk + (size_t)(2 * b) * RS_BIAS, n, blob[x.b_off + n]);
        put_bias(lb + ly.bblk + (size_t)(2 * b + 1) * RS_BIAS, n, blob[y.b_off + n]);
      }
    }
    // final layer: the program's columns are (shift_i, scale_i) pairs -> rows i and 16 + i of ONE chunk
    const FlowOp& c = o[1 + 2 * NB];
    for (int i = 0; i < c.d_tr; ++i)
      for (int part = 0; part < (additive ? 1 : 2); ++part) {
        const int col = 2 * i + part, row = 16 * part + i;  // (additive coupling: the scale column is unused)
        for (int k = 0; k < H; ++k) tc_put(lb + ly.wf, lb + ly.wf + NS_WF, NS_CN, row, k, blob[c.w_off + k * c.Npad + col]);
        put_bias(lb + ly.bf, row, blob[c.b_off + col]);
      }
    if (l == L - 1) {
      // the affine after the last coupling: fp32, k-major over the slots of the last layer
      const FlowOp& fl = ops[per * L];
      float* A = reinterpret_cast<float*>(lb + ly.aff);
      for (int k = 0; k < D; ++k)
        for (int n = 0; n < D; ++n) A[slot(L - 1, k) * NS_DP + n] = blob[fl.w_off + k * fl.Npad + n];
      for (int n = 0; n < D; ++n) A[NS_DP * NS_DP + n] = blob[fl.b_off + n];
    }
    if (tc_put_overflow()) return 0;
    if (cudaMalloc(&t.d_image[l], bytes) != cudaSuccess) return 2;
    if (cudaMemcpy(t.d_image[l], img.data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess) return 2;
    t.layer[l].d_tr = d_tr[l];
  }
  t.image_bytes = bytes;
  for (int j = 0; j < NS_DP; ++j) t.out_slot[j] = (int8_t)j;
  t.L = L;
  t.D = D;
  t.NB = NB;
  t.nch = 1;
  t.inverse = inverse;
  t.additive = additive;
  t.ac = 1;
  t.narrow = H <= TC_H / 2;
  t.tail_bound = 0.f;
  t.valid = true;
  return 0;
}

// ------------------------------------------------------------------ device
__device__ __forceinline__ void ns_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ uint32_t ns_ld1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return r;
}
__device__ __forceinline__ void ns_st1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void ns_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
      "%14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ float ns_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ns_softplus(float u) {
  // log(1 + e^u); for u > 20 it is u to fp32 precision
  const float v = 0.6931471805599453f * tc_lg2(1.f + tc_ex2(u * 1.4426950408889634f));
  return u > 20.f ? u : v;
}

// nflows' unconstrained_rational_quadratic_spline for ONE feature, linear tails, 8 bins:
// p[0..7] unnormalised widths, p[8..15] heights (both already divided by sqrt(hidden) in the
// folded weights), p[16..22] unnormalised derivatives.  Returns the output; ld gets +-log|dy/dx|.
template <bool INVERSE>
__device__ __forceinline__ float ns_spline(const float (&p)[NS_G], float x, float B, float& ld) {
  constexpr float MINS = 1e-3f, L2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  float ew[NS_K], eh[NS_K];
  float mw = p[0], mh = p[NS_K];
#pragma unroll
  for (int k = 1; k < NS_K; ++k) mw = fmaxf(mw, p[k]), mh = fmaxf(mh, p[NS_K + k]);
  float sw = 0.f, sh = 0.f;
#pragma unroll
  for (int k = 0; k < NS_K; ++k) {
    ew[k] = tc_ex2((p[k] - mw) * L2E);
    eh[k] = tc_ex2((p[NS_K + k] - mh) * L2E);
    sw += ew[k];
    sh += eh[k];
  }
  const float cw_ = (1.f - MINS * NS_K) * tc_rcp(sw), chh = (1.f - MINS * NS_K) * tc_rcp(sh);
  // knots, bin search and selection in one sweep (knot 0 = -B, knot 8 = +B)
  float aw = 0.f, ah = 0.f, lw = -B, lh = -B;
  float icw = -B, iw = 1.f, ich = -B, ih = 1.f, u0 = 0.f, u1 = 0.f;
  int b = 0;
#pragma unroll
  for (int k = 0; k < NS_K; ++k) {
    aw += MINS + cw_ * ew[k];
    ah += MINS + chh * eh[k];
    const float rw = (k == NS_K - 1) ? B : fmaf(2.f * B, aw, -B);
    const float rh = (k == NS_K - 1) ? B : fmaf(2.f * B, ah, -B);
    const bool in_or_above = k == 0 || x >= (INVERSE ? lh : lw);  // knots increase: the last true wins
    if (in_or_above) {
      b = k;
      icw = lw, iw = rw - lw, ich = lh, ih = rh - lh;
      u0 = k == 0 ? 0.f : p[2 * NS_K + k - 1];
      u1 = k == NS_K - 1 ? 0.f : p[2 * NS_K + k];
    }
    lw = rw, lh = rh;
  }
  const float d0 = b == 0 ? 1.f : MINS + ns_softplus(u0);
  const float d1 = b == NS_K - 1 ? 1.f : MINS + ns_softplus(u1);
  const float delta = ih * tc_rcp(iw);
  const float q = d0 + d1 - 2.f * delta;
  float out, lad;
  if (INVERSE) {
    const float dy = x - ich;
    const float a = dy * q + ih * (delta - d0);
    const float bb = ih * d0 - dy * q;
    const float c = -delta * dy;
    const float disc = bb * bb - 4.f * a * c;  // < 0 -> NaN (the reference asserts)
    const float root = (2.f * c) * tc_rcp(-bb - ns_sqrt(disc));
    out = fmaf(root, iw, icw);
    const float t1m = root * (1.f - root);
    const float den = delta + q * t1m;
    const float dnum = delta * delta * (d1 * root * root + 2.f * delta * t1m + d0 * (1.f - root) * (1.f - root));
    lad = -LN2 * (tc_lg2(dnum) - 2.f * tc_lg2(den));
  } else {
    const float th = (x - icw) * tc_rcp(iw);
    const float t1m = th * (1.f - th);
    const float num = ih * (delta * th * th + d0 * t1m);
    const float den = delta + q * t1m;
    out = fmaf(num, tc_rcp(den), ich);
    const float dnum = delta * delta * (d1 * th * th + 2.f * delta * t1m + d0 * (1.f - th) * (1.f - th));
    lad = LN2 * (tc_lg2(dnum) - 2.f * tc_lg2(den));
  }
  const bool inside = x >= -B && x <= B;  // linear tails: identity (NaN falls through unchanged)
  ld += inside ? lad : 0.f;
  return inside ? out : x;
}

struct NsShared {
  uint64_t bar_in[RS_NG];
  uint64_t bar_out[RS_NG];
  uint64_t bar_img;  // completion of the weight image's bulk copy
  uint64_t bar_cr[RS_NG][2];  // final-layer chunk ready in D2 / D (tcgen05.commit -> epilogue)
  uint64_t bar_cf[RS_NG][2];  // ... and read out again (epilogue arrivals -> issuer)
  uint32_t tmem_base;
  uint32_t pad;
  double cst[4][NS_DP];
  double log_const;
};

__device__ __forceinline__ void ns_tile_sync(int g) {
  // the eight epilogue warps of tile g (TMEM written by one warp, read by its lane-quarter twin)
  tc_fence_before();
  asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(RS_EW * 32) : "memory");
  tc_fence_after();
}

// One layer pass for one row; the state is in TMEM columns NS_COL_ST.., the log|det| partial of
// this thread is returned (c == 0 and c == 1 threads of a row each sum their own features).
// Barriers of one tile group (shared-memory addresses) and the phase bits of this thread.
struct NsBars {
  uint32_t in, out, cr0, cr1, cf0, cf1;
  uint32_t ph, pc0, pc1;
};

// Spline of feature i = 2 j + c from the chunk parameters in TMEM columns `col`; the chunk buffer
// is handed back to the issuer as soon as the parameters are in registers.
__device__ __forceinline__ void ns_chunk(const NsParams& P, uint32_t tg, int c, int j, uint32_t col, uint32_t bar_cr,
                                         uint32_t& pc, uint32_t bar_cf, float& ld) {
  rs_wait(bar_cr, pc);
  const int i = 2 * j + c;
  const bool mine = i < P.ly.d_tr;
  uint32_t ra[16], rb[8], xv = 0u;
  const uint32_t st = tg + NS_COL_ST + (mine ? P.ly.trslot[i] : 0);
  if (mine) {
    tc_ld16(tg + col + NS_G * c, ra);
    ns_ld8(tg + col + NS_G * c + 16, rb);
    xv = ns_ld1(st);
    tc_wait_ld();
    tc_pin16(ra);
  }
  if (j + 2 < P.nch) rs_arrive(bar_cf);
  if (mine) {
    float p[NS_G];
#pragma unroll
    for (int k = 0; k < 16; ++k) p[k] = __uint_as_float(ra[k]);
#pragma unroll
    for (int k = 0; k < 8; ++k) p[16 + k] = __uint_as_float(rb[k]);
    const float y = P.inverse ? ns_spline<true>(p, __uint_as_float(xv), P.tail_bound, ld)
                              : ns_spline<false>(p, __uint_as_float(xv), P.tail_bound, ld);
    ns_st1(st, __float_as_uint(y));
    tc_wait_st();
  }
}

// AC mode: nflows' AffineCouplingTransform on the transformed slots AC_TR0 + i, i = 8 c .. 8 c + 7 of
// this thread; the final layer's single chunk holds shift_i in column i and the unconstrained
// scale_i in column 16 + i (scale = sigmoid(u + 2) + 1e-3; additive coupling: scale = 1).  Slots
// >= d_tr carry zero weights and a zero state: only their log-scale has to be masked.
__device__ __forceinline__ void ac_coupling(const NsParams& P, uint32_t tg, int c, uint32_t bar_cr, uint32_t& pc,
                                            float& ld) {
  rs_wait(bar_cr, pc);
  uint32_t sh[8], sc[8], st[8];
  ns_ld8(tg + NS_COL_D2 + 8 * c, sh);
  ns_ld8(tg + NS_COL_D2 + 16 + 8 * c, sc);
  ns_ld8(tg + NS_COL_ST + AC_TR0 + 8 * c, st);
  tc_wait_ld();
  float s[8];
#pragma unroll
  for (int f = 0; f < 8; ++f) s[f] = tc_ex2((__uint_as_float(sc[f]) + 2.f) * -1.4426950408889634f);
#pragma unroll
  for (int f = 0; f < 8; ++f) s[f] = P.additive ? 1.f : tc_rcp(1.f + s[f]) + 1e-3f;
  float acc = 0.f;
#pragma unroll
  for (int f = 0; f < 8; ++f) {
    const float t = __uint_as_float(sh[f]), x = __uint_as_float(st[f]);
    const float y = P.inverse ? (x - t) * tc_rcp(s[f]) : fmaf(x, s[f], t);
    st[f] = __float_as_uint(y);
    const float ls = tc_lg2(s[f]) * 0.6931471805599453f;
    acc += (8 * c + f < P.ly.d_tr && !P.additive) ? ls : 0.f;
  }
  ld += P.inverse ? -acc : acc;
  tc_st8(tg + NS_COL_ST + AC_TR0 + 8 * c, st);
  tc_wait_st();
}

// One layer pass for one row; the state is in TMEM columns NS_COL_ST.., the log|det| partial of
// this thread is returned (c == 0 and c == 1 threads of a row each sum their own features).
template <bool NARROW, bool AC>
__device__ __forceinline__ float ns_run_layer(const NsParams& P, const NsLayout& lay, uint32_t tg, int g, int c,
                                              NsBars& B) {
  float ld = 0.f;
  {  // state slots 16c .. 16c+15 -> K chunk c of the A operand
    uint32_t r[16], hi[8], lo[8];
    tc_ld16(tg + NS_COL_ST + 16 * c, r);
    tc_wait_ld();
    tc_pin16(r);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      tc_split2<false>(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]), hi[j], lo[j]);
    tc_st8(tg + NS_COL_AH + 8 * c, hi);
    tc_st8(tg + NS_COL_AL + 8 * c, lo);
    tc_wait_st();
  }
  rs_arrive(B.in);  // -> G0
  for (int b = 0; b < P.NB; ++b) {
    rs_wait(B.out, B.ph);
    if (AC && b == 0) {
      // the state after this layer's affine: G0's columns 64 .. 95 (the first columns of D2, which
      // the first block's GEMM overwrites only after this epilogue has arrived) -> own state slots
      uint32_t r[16];
      tc_ld16(tg + NS_COL_D2 + 16 * c, r);
      tc_wait_ld();
      tc_pin16(r);
      ns_st16(tg + NS_COL_ST + 16 * c, r);
    }
    rs_hidden<true, NARROW>(tg, NS_COL_D, c);
    rs_arrive(B.in);
    rs_wait(B.out, B.ph);
    rs_hidden<true, NARROW>(tg, NS_COL_D2, c);
    rs_arrive(B.in);
  }
  rs_wait(B.out, B.ph);
  rs_hidden<false, NARROW>(tg, NS_COL_D, c);
  // The final layer's chunks alternate between the two accumulators (the residual stream in D is
  // dead once its activation is the A operand), so chunk j + 1 is computed while chunk j's
  // splines run: Gf chunk j: (D2 | D)[0:48] = Wf_j a + bf_j
  rs_arrive(B.in);
  if (AC) {
    ac_coupling(P, tg, c, B.cr0, B.pc0, ld);
  } else {
    for (int j = 0; j < P.nch; j += 2) {
      ns_chunk(P, tg, c, j, NS_COL_D2, B.cr0, B.pc0, B.cf0, ld);
      if (j + 1 < P.nch) ns_chunk(P, tg, c, j + 1, NS_COL_D, B.cr1, B.pc1, B.cf1, ld);
    }
  }
  ns_tile_sync(g);  // the state written by the twin warp is visible before the next split
  return ld;
}

template <int NKS, bool AC>
__device__ __forceinline__ void ns_issuer(const NsParams& P, const NsLayout& lay, uint32_t img_s, uint32_t tg,
                                          const NsBars& B, int64_t my_tiles) {
  const uint32_t bar_in = B.in, bar_out = B.out;
  constexpr uint32_t ID64 = tc_idesc(128, TC_H), ID48 = tc_idesc(128, NS_CN);
  const uint32_t d = tg + NS_COL_D, d2 = tg + NS_COL_D2, ah = tg + NS_COL_AH, al = tg + NS_COL_AL;
  const uint32_t ones_s = img_s + lay.layer_bytes;
  const uint32_t zero_s = ones_s + TC_ONES_BYTES;
  const uint64_t ones = tc_desc(ones_s, 2048, 128);
  auto adv = [](uint64_t desc, uint32_t off) { return desc + (uint64_t)(off >> 4); };
  auto bias = [&](uint32_t addr) { return tc_desc(addr, zero_s - addr, 128); };
  auto gemm64 = [&](uint32_t dst, uint64_t bdesc, uint64_t whi, uint64_t wlo, uint32_t rows, uint32_t idesc,
                    uint32_t acc0) {
    tc_mma_ss_e(dst, ones, bdesc, idesc, acc0);
#pragma unroll
    for (int ks = 0; ks < NKS; ++ks) {  // K-steps of 16 hidden units (4; 2 for a conditioner of width <= 32)
      tc_mma_ts_e(dst, ah + 8 * ks, adv(whi, ks * 2 * rows * 16), idesc, 1);
      tc_mma_ts_e(dst, al + 8 * ks, adv(whi, ks * 2 * rows * 16), idesc, 1);
      tc_mma_ts_e(dst, ah + 8 * ks, adv(wlo, ks * 2 * rows * 16), idesc, 1);
    }
  };
  const uint32_t lb = img_s;
  const uint64_t d64 = tc_desc(lb, TC_H * 16, 128);
  const uint64_t d48 = tc_desc(lb, NS_CN * 16, 128);
  uint32_t ph = 0, pf0 = 0, pf1 = 0;
  for (int64_t it = 0; it < my_tiles; ++it) {
    // G0: K = 32 state slots (2 k-steps)
    tc_mbar_wait(bar_in, ph);
    ph ^= 1;
    tc_fence_after();
    // (AC: N = 96 -- columns 64 .. 95, the first of D2, receive the state after the layer's affine)
    constexpr int ROWS0 = AC ? AC_ROWS0 : TC_H;
    constexpr uint32_t ID0 = tc_idesc(128, ROWS0);
    const uint64_t d0 = tc_desc(lb, ROWS0 * 16, 128);
    tc_mma_ss_e(d, ones, bias(lb + lay.b0), ID0, 0);
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      tc_mma_ts_e(d, ah + 8 * ks, adv(d0, lay.w0hi + ks * 2 * ROWS0 * 16), ID0, 1);
      tc_mma_ts_e(d, al + 8 * ks, adv(d0, lay.w0hi + ks * 2 * ROWS0 * 16), ID0, 1);
      tc_mma_ts_e(d, ah + 8 * ks, adv(d0, lay.w0lo + ks * 2 * ROWS0 * 16), ID0, 1);
    }
    tc_commit_e(bar_out);
    for (int b = 0; b < P.NB; ++b) {
      const uint32_t wb = lay.blk + b * 4 * RS_W_BIG;
      tc_mbar_wait(bar_in, ph);
      ph ^= 1;
      tc_fence_after();
      gemm64(d2, bias(lb + lay.bblk + 2 * b * RS_BIAS), adv(d64, wb), adv(d64, wb + RS_W_BIG), TC_H, ID64, 0);
      tc_commit_e(bar_out);
      tc_mbar_wait(bar_in, ph);
      ph ^= 1;
      tc_fence_after();
      gemm64(d, bias(lb + lay.bblk + (2 * b + 1) * RS_BIAS), adv(d64, wb + 2 * RS_W_BIG),
             adv(d64, wb + 3 * RS_W_BIG), TC_H, ID64, 1);
      tc_commit_e(bar_out);
    }
    tc_mbar_wait(bar_in, ph);
    ph ^= 1;
    tc_fence_after();
    for (int j = 0; j < P.nch; ++j) {
      const bool odd = j & 1;
      if (j >= 2) {  // the buffer's previous chunk has been read out
        if (odd) {
          tc_mbar_wait(B.cf1, pf1);
          pf1 ^= 1;
        } else {
          tc_mbar_wait(B.cf0, pf0);
          pf0 ^= 1;
        }
        tc_fence_after();
      }
      gemm64(odd ? d : d2, bias(lb + lay.bf + j * NS_BF), adv(d48, lay.wf + j * 2 * NS_WF),
             adv(d48, lay.wf + j * 2 * NS_WF + NS_WF), NS_CN, ID48, 0);
      tc_commit_e(odd ? B.cr1 : B.cr0);
    }
  }
}

// MODE 0: apply (rows supplied), MODE 1: populate.  One coupling layer per launch.
template <int MODE, bool NARROW, bool AC>
__global__ void __launch_bounds__(RS_THREADS, 1) flow_tc_nsf_kernel(NsParams P, TcIO io, PopulateArgs A) {
  extern __shared__ __align__(1024) uint8_t ns_smem[];
  NsShared* sh = reinterpret_cast<NsShared*>(ns_smem + tc_image_pad(P.image_bytes));
  const int tid = threadIdx.x;
  const NsLayout lay = ns_layout(P.NB, P.nch, AC);
  if (MODE == 1 && P.last) {
    if (tid < 4 * NS_DP) {
      const int which = tid / NS_DP, d = tid % NS_DP;
      const double* src = which == 0 ? A.scale : which == 1 ? A.shift : which == 2 ? A.lo : A.hi;
      sh->cst[which][d] = d < P.D ? src[d] : 0.0;
    }
    if (tid == 4 * NS_DP) sh->log_const = populate_log_const(A, P.D);
  }
  if (tid == 0) {
    tc_image_load(tc_smem_u32(ns_smem), P.image, (uint32_t)P.image_bytes, tc_smem_u32(&sh->bar_img));
    for (int g = 0; g < RS_NG; ++g) {
      tc_mbar_init(tc_smem_u32(&sh->bar_in[g]), RS_EW * 32);
      tc_mbar_init(tc_smem_u32(&sh->bar_out[g]), 1);
      for (int b = 0; b < 2; ++b) {
        tc_mbar_init(tc_smem_u32(&sh->bar_cr[g][b]), 1);
        tc_mbar_init(tc_smem_u32(&sh->bar_cf[g][b]), RS_EW * 32);
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int warp = tid >> 5;
  if (warp == RS_NG * RS_EW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     tc_smem_u32(&sh->tmem_base)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  tc_mbar_wait(tc_smem_u32(&sh->bar_img), 0);  // the weight image has landed
  const int64_t n = MODE == 1 ? A.n : io.n;
  const int64_t ntiles = (n + 127) / 128;
  const uint32_t tmem = sh->tmem_base;
  if (warp < RS_NG * RS_EW) {
    const int g = warp / RS_EW, w8 = warp % RS_EW, q = w8 & 3, c = w8 >> 2;
    const uint32_t tg = tmem + g * NS_COLS + ((uint32_t)(q * 32) << 16);
    NsBars B{tc_smem_u32(&sh->bar_in[g]),    tc_smem_u32(&sh->bar_out[g]),   tc_smem_u32(&sh->bar_cr[g][0]),
             tc_smem_u32(&sh->bar_cr[g][1]), tc_smem_u32(&sh->bar_cf[g][0]), tc_smem_u32(&sh->bar_cf[g][1]),
             0u, 0u, 0u};
    double vmax = -INFINITY, vcount = 0.0;
    const int64_t stride = (int64_t)gridDim.x * RS_NG;
    for (int64_t tile = (int64_t)blockIdx.x * RS_NG + g; tile < ntiles; tile += stride) {
      const int64_t row = tile * 128 + q * 32 + (tid & 31);
      const bool valid = row < n;
      float ss = 0.f, ld0 = 0.f;
      bool alive = true;
      // ---- load the state: half c owns slots 16c .. 16c+15
      {
        uint32_t r[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) r[k] = 0u;
        if (!P.first) {
          if (valid) {
            const float4* p4 = reinterpret_cast<const float4*>(P.sc_h + row * NS_DP + 16 * c);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float4 v = p4[k];
              r[4 * k] = __float_as_uint(v.x), r[4 * k + 1] = __float_as_uint(v.y);
              r[4 * k + 2] = __float_as_uint(v.z), r[4 * k + 3] = __float_as_uint(v.w);
            }
            if (c == 0) {
              ld0 = P.sc_ld[row];
              const float e = P.sc_ss[row];
              alive = e >= 0.f;
              ss = alive ? e : -e - 1.f;
            }
          }
        } else if (MODE == 0) {
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const int d = 16 * c + k;
            const float v = (valid && d < P.D) ? __ldg(io.in + row * P.D + d) : 0.f;
            r[k] = __float_as_uint(v);
            ss = fmaf(v, v, ss);
          }
        } else {
#pragma unroll
          for (int k0 = 0; k0 < 16; k0 += 4) {
            const int d0 = 16 * c + k0;
            float v[4];
            const Philox4 rr = philox4x32_10(A.seed, A.row_offset + row, d0 / 4, 0);
            box_muller(rr.x, rr.y, v[0], v[1]);
            box_muller(rr.z, rr.w, v[2], v[3]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const bool use = d0 + j < P.D;
              ss = use ? fmaf(v[j], v[j], ss) : ss;
              const float zz = use ? v[j] * A.sqrt_t : 0.f;
              r[k0 + j] = __float_as_uint(zz);
              if (A.z && use && valid) A.z[row * P.D + d0 + j] = zz;
            }
          }
        }
        ns_st16(tg + NS_COL_ST + 16 * c, r);
        if (P.first) {
          // the row's sum of squares needs both halves: through the LD columns
          ns_st1(tg + NS_COL_LD + c, __float_as_uint(ss));
        }
        tc_wait_st();
        ns_tile_sync(g);
        if (P.first) {
          const float other = __uint_as_float(ns_ld1(tg + NS_COL_LD + (1 - c)));
          tc_wait_ld();
          ss += other;
          if (MODE == 1) {
            const float rad = sqrtf(ss) * A.sqrt_t;
            alive = !(A.r_max > 0.f) || (rad <= A.r_max);
          }
          ns_tile_sync(g);  // LD columns are reused below
        }
      }
      float ld = ns_run_layer<NARROW, AC>(P, lay, tg, g, c, B);
      // ---- combine the two halves' log|det| and hand the row on
      ns_st1(tg + NS_COL_LD + c, __float_as_uint(ld));
      tc_wait_st();
      ns_tile_sync(g);
      if (c == 0) {
        const float other = __uint_as_float(ns_ld1(tg + NS_COL_LD + 1));
        uint32_t s0[16], s1[16];
        tc_ld16(tg + NS_COL_ST, s0);
        tc_ld16(tg + NS_COL_ST + 16, s1);
        tc_wait_ld();
        tc_pin16(s0);
        tc_pin16(s1);
        ld += other + ld0;
        if (!P.last) {
          if (valid) {
            float4* o4 = reinterpret_cast<float4*>(P.sc_h + row * NS_DP);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              o4[k] = make_float4(__uint_as_float(s0[4 * k]), __uint_as_float(s0[4 * k + 1]),
                                  __uint_as_float(s0[4 * k + 2]), __uint_as_float(s0[4 * k + 3]));
              o4[4 + k] = make_float4(__uint_as_float(s1[4 * k]), __uint_as_float(s1[4 * k + 1]),
                                      __uint_as_float(s1[4 * k + 2]), __uint_as_float(s1[4 * k + 3]));
            }
            P.sc_ld[row] = ld;
            P.sc_ss[row] = alive ? ss : -ss - 1.f;
          }
        } else {
          // output feature j = state slot out_slot[j]: gather through registers with a
          // compile-time sweep (no dynamic register indexing)
          float outv[NS_DP];
          if (AC) {
            // the affine after the last coupling (slots of the last layer -> natural order), fp32 on
            // the CUDA cores: k-major [32][32] + bias in the image, warp-wide broadcast loads
            const float* aff = reinterpret_cast<const float*>(ns_smem + lay.aff);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              float o[16];
              const float4* b4 = reinterpret_cast<const float4*>(aff + NS_DP * NS_DP + 16 * half);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 b = b4[q];
                o[4 * q] = b.x, o[4 * q + 1] = b.y, o[4 * q + 2] = b.z, o[4 * q + 3] = b.w;
              }
#pragma unroll
              for (int k = 0; k < NS_DP; ++k) {
                const float hk = __uint_as_float(k < 16 ? s0[k & 15] : s1[k & 15]);
                const float4* w4 = reinterpret_cast<const float4*>(aff + k * NS_DP + 16 * half);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float4 w = w4[q];
                  tc_fma2(o[4 * q + 0], o[4 * q + 1], w.x, w.y, hk, hk);
                  tc_fma2(o[4 * q + 2], o[4 * q + 3], w.z, w.w, hk, hk);
                }
              }
#pragma unroll
              for (int j = 0; j < 16; ++j) outv[16 * half + j] = o[j];
            }
          } else {
#pragma unroll
          for (int j = 0; j < NS_DP; ++j) {
            float v = 0.f;
            const int s = j < P.D ? P.out_slot[j] : -1;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              v = s == k ? __uint_as_float(s0[k]) : v;
              v = s == 16 + k ? __uint_as_float(s1[k]) : v;
            }
            outv[j] = v;
          }
          }
          const float logj = ld + P.const_logdet;
          if (MODE == 0) {
            float ss_out = 0.f;
#pragma unroll
            for (int j = 0; j < NS_DP; ++j) ss_out = j < P.D ? fmaf(outv[j], outv[j], ss_out) : ss_out;
            if (valid) {
              if (io.out) {
#pragma unroll
                for (int j = 0; j < NS_DP; ++j)
                  if (j < P.D) io.out[row * P.D + j] = outv[j];
              }
              if (io.out_logj) io.out_logj[row] = logj;
              if (io.out_lp) {
                const float cn = io.base_log_z, hv = 0.5f * io.base_inv_var;
                io.out_lp[row] = (io.lp_mode == 1) ? (-hv * ss - cn) - logj : (-hv * ss_out - cn) + logj;
              }
            }
          } else {
            const float base_lp = -0.5f * ss - 0.5f * P.D * TC_LOG_2PI;
            populate_row<NS_DP>(A, P.D, [&](int d) { return outv[d]; }, row, alive, base_lp, logj, vmax, vcount,
                                sh->cst[0], sh->cst[1], sh->cst[2], sh->cst[3], sh->log_const);
          }
        }
      }
      ns_tile_sync(g);  // the next tile's state store must not overtake these reads
    }
    if (MODE == 1 && P.last && c == 0) populate_publish(A, vmax, vcount);
  } else {
    const int g = __shfl_sync(0xffffffffu, warp - RS_NG * RS_EW, 0);
    const NsBars B{tc_smem_u32(&sh->bar_in[g]),    tc_smem_u32(&sh->bar_out[g]),   tc_smem_u32(&sh->bar_cr[g][0]),
                   tc_smem_u32(&sh->bar_cr[g][1]), tc_smem_u32(&sh->bar_cf[g][0]), tc_smem_u32(&sh->bar_cf[g][1]),
                   0u, 0u, 0u};
    ns_issuer<NARROW ? 2 : 4, AC>(P, lay, tc_smem_u32(ns_smem), __shfl_sync(0xffffffffu, tmem, 0) + g * NS_COLS, B,
              rs_my_tiles(ntiles, g));
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == RS_NG * RS_EW) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(sh->tmem_base), "r"(512u)
                 : "memory");
  }
}

inline size_t ns_smem_bytes(int image_bytes) {
  return (((size_t)image_bytes + 1023) & ~(size_t)1023) + sizeof(NsShared) + 64;
}
inline int ns_reserve(NsProgram& t, int64_t n) {
  if (t.L <= 1 || n <= t.scratch_rows) return 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return 1;
  if (t.d_scratch) cudaFree(t.d_scratch);
  t.d_scratch = nullptr;
  t.scratch_rows = 0;
  if (cudaMalloc(&t.d_scratch, sizeof(float) * (size_t)n * (NS_DP + 2)) != cudaSuccess) return 1;
  t.scratch_rows = n;
  return 0;
}

// One launch per coupling layer; returns the number of launches (0 on failure).
template <int MODE>
inline int ns_launch(NsProgram& t, const TcIO& io, const PopulateArgs& A, int64_t n, int num_sms, cudaStream_t st) {
  if (ns_reserve(t, n)) return 0;
  const size_t smem = ns_smem_bytes(t.image_bytes);
  const void* kern = t.ac ? (t.narrow ? (const void*)flow_tc_nsf_kernel<MODE, true, true>
                                      : (const void*)flow_tc_nsf_kernel<MODE, false, true>)
                          : (t.narrow ? (const void*)flow_tc_nsf_kernel<MODE, true, false>
                                      : (const void*)flow_tc_nsf_kernel<MODE, false, false>);
  if (tc_prep(kern, smem)) return 0;
  for (int l = 0; l < t.L; ++l) {
    NsParams P;
    P.image = t.d_image[l];
    const NsLayout lay = ns_layout(t.NB, t.nch, t.ac);
    P.image_bytes = lay.layer_bytes + TC_ONES_BYTES + (t.ac ? AC_ZERO_BYTES : TC_ZERO_BYTES);
    P.additive = t.additive;
    P.D = t.D;
    P.NB = t.NB;
    P.nch = t.nch;
    P.inverse = t.inverse;
    P.first = l == 0;
    P.last = l == t.L - 1;
    P.ly = t.layer[l];
    for (int j = 0; j < NS_DP; ++j) P.out_slot[j] = t.out_slot[j];
    P.tail_bound = t.tail_bound;
    P.const_logdet = t.const_logdet;
    P.sc_h = t.d_scratch;
    P.sc_ld = t.d_scratch ? t.d_scratch + (size_t)t.scratch_rows * NS_DP : nullptr;
    P.sc_ss = t.d_scratch ? t.d_scratch + (size_t)t.scratch_rows * (NS_DP + 1) : nullptr;
    const int grid = rs_grid(n, num_sms);
    if (t.ac && t.narrow) flow_tc_nsf_kernel<MODE, true, true><<<grid, RS_THREADS, smem, st>>>(P, io, A);
    else if (t.ac) flow_tc_nsf_kernel<MODE, false, true><<<grid, RS_THREADS, smem, st>>>(P, io, A);
    else if (t.narrow) flow_tc_nsf_kernel<MODE, true, false><<<grid, RS_THREADS, smem, st>>>(P, io, A);
    else flow_tc_nsf_kernel<MODE, false, false><<<grid, RS_THREADS, smem, st>>>(P, io, A);
    if (cudaGetLastError() != cudaSuccess) return 0;
  }
  return t.L;
}

}  // namespace nb200
