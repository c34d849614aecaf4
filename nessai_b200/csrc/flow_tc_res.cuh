// tcgen05 / TMEM kernel for RealNVP with the DEFAULT conditioner, nflows' ResidualNet
// (/root/reference/src/nessai/flows/realnvp.py:133-146):  d_id -> H, NB <= 3 residual blocks
// h += lin1(relu(lin0(relu(h)))), H -> 2*d_tr  (ReLU, D <= 16, H <= 64: narrower nets -- the
// reference's default is H = 2 D -- are zero-padded to 64 hidden units in the weight image).
//
// Same scheme as flow_tc.cuh (row == TMEM lane, activations as the A operand in tensor memory,
// split-bf16 3-pass MMAs, converged issuer warps), plus what the residual net needs:
//   * the residual stream h lives in a TMEM accumulator for the whole layer; the second linear
//     of a block ACCUMULATES into it (accumulate = 1 on every MMA incl. the bias MMA), so the
//     residual add costs nothing;
//   * three 64-column regions per tile (stream D, scratch accumulator D2, A operand), i.e. two
//     128-row tiles in flight per SM; each tile gets EIGHT epilogue warps -- two per TMEM lane
//     quarter, splitting the 64 hidden columns -- so the ReLU + split epilogues are half as long;
//   * one layer is ~77 KB of bf16 hi/lo weights, so only a few layers fit in shared memory: the
//     flow runs in PASSES of consecutive layers, handing the row state (16 floats + log|det|)
//     to the next pass through an L2-resident scratch buffer (72 B/row against ~150 kFLOP/row).
#pragma once
#include "flow_tc.cuh"

namespace nb200 {

constexpr int RS_MAXNB = 3;                    // residual blocks per conditioner (n_layers of the reference)
constexpr int RS_MAXPASS = 8;
constexpr int RS_NG = 2;                      // tiles in flight per SM
constexpr int RS_EW = 8;                      // epilogue warps per tile
constexpr int RS_ETHREADS = RS_NG * RS_EW * 32;
constexpr int RS_THREADS = RS_ETHREADS + RS_NG * 32;
constexpr int RS_COLS = 192;                  // TMEM columns per tile
constexpr int RS_COL_D = 0;                   // residual stream
constexpr int RS_COL_D2 = 64;                 // block-internal / final-layer accumulator
constexpr int RS_COL_AH = 128;                // A operand, hi
constexpr int RS_COL_AL = 160;                // A operand, lo
constexpr int RS_W_SMALL = TC_N1 * 16 * 2;    // K = 16 operand of G0: 64 rows (+ 16: the affine, flow_tc.cuh TC_AFFMMA)
constexpr int RS_BIAS0 = TC_N1 * 16;          // G0's bias operand
constexpr int RS_W_BIG = TC_H * 16 * (TC_H / 8);    // 64 x 64: 8 KB
constexpr int RS_W_FIN = TC_N3 * 16 * (TC_H / 8);   // 16 x 64: 2 KB
constexpr int RS_BIAS = TC_H * 16;            // bias operand (K-chunk 0 only): 1 KB

struct RsLayout {  // byte offsets inside one layer of the image
  int w0hi, w0lo, blk, wfhi, wflo, b0, bblk, bf, layer_bytes;
  int hrows, wbig, wfin, bias;  // hidden rows of a B operand; bytes of a hidden / final weight matrix, a bias operand
};
// narrow (conditioner width <= 32): the hidden operands are 32 rows x 32 K (2 KB instead of 8 KB), so a
// layer is 27 KB instead of 77 KB and the whole flow fits ONE pass (no scratch hand-off)
__host__ __device__ inline RsLayout rs_layout(int NB, bool narrow = false) {
  RsLayout o;
  o.hrows = narrow ? TC_H / 2 : TC_H;
  o.wbig = o.hrows * 16 * (o.hrows / 8);
  o.wfin = TC_N3 * 16 * (o.hrows / 8);
  o.bias = o.hrows * 16;
  o.w0hi = 0;
  o.w0lo = RS_W_SMALL;
  o.blk = 2 * RS_W_SMALL;                 // per block: WA hi, WA lo, WB hi, WB lo
  o.wfhi = o.blk + NB * 4 * o.wbig;
  o.wflo = o.wfhi + o.wfin;
  o.b0 = o.wflo + o.wfin;
  o.bblk = o.b0 + RS_BIAS0;               // per block: bA, bB
  o.bf = o.bblk + NB * 2 * o.bias;
  o.layer_bytes = o.bf + TC_N3 * 16;
  return o;
}

struct RsPass {
  uint8_t* d_image = nullptr;
  int image_bytes = 0;
  int l0 = 0, nl = 0;
};
struct RsProgram {
  bool valid = false;
  int L = 0, D = 0, NB = 0, n_pass = 0;
  int d_id[TC_MAXL] = {0}, d_tr[TC_MAXL] = {0};
  int additive = 0, inverse = 0;
  int narrow = 0;  // conditioner width <= 32 (see rs_hidden_quarter)
  int act = ACT_RELU;  // conditioner activation (a compile-time parameter of the kernel)
  float const_logdet = 0.f;
  RsPass pass[RS_MAXPASS];
  float* d_scratch = nullptr;  // [rows][16] state | [rows] log|det| | [rows] sum z^2 (sign: alive)
  int64_t scratch_rows = 0;
};
inline void rs_free(RsProgram& t) {
  for (int p = 0; p < RS_MAXPASS; ++p)
    if (t.pass[p].d_image) cudaFree(t.pass[p].d_image);
  if (t.d_scratch) cudaFree(t.d_scratch);
  t = RsProgram();
}

struct RsParams {
  const uint8_t* image;
  int image_bytes;
  int l0, nl, L, D, NB;
  int d_id[TC_MAXL], d_tr[TC_MAXL];
  int additive, inverse, first, last;
  int narrow;
  float const_logdet;
  float* sc_h;
  float* sc_ld;
  float* sc_ss;
};

// Recognise  affine (resnet-coupling affine)*  and build one image per pass.
inline int rs_build(RsProgram& t, const FlowOp* ops, int n_ops, const float* blob, int D, int H,
                    int activation) {
  t.valid = false;
  if (D > TC_DP || H < 1 || H > TC_H || activation < ACT_RELU || activation > ACT_SILU || n_ops < 6) return 0;
  // ops per layer: affine, initial linear, 2 per block, coupling
  int NB = -1;
  for (int nb = 1; nb <= RS_MAXNB; ++nb)
    if ((n_ops - 1) % (2 * nb + 3) == 0 && ops[2].type == OP_LINEAR &&
        ops[2 * nb + 2].type == OP_COUPLING_AFFINE)
      NB = nb;
  if (NB < 0) return 0;
  const int per = 2 * NB + 3;
  const int L = (n_ops - 1) / per;
  if (L < 1 || L > TC_MAXL) return 0;
  auto is_affine = [&](const FlowOp& o) {
    return o.type == OP_LINEAR && o.src <= BUF_X1 && o.dst <= BUF_X1 && o.K == D && o.N == D &&
           o.flags == 0 && o.src_off == 0;
  };
  if (!is_affine(ops[0])) return 0;
  int inverse = -1, additive = -1;
  for (int l = 0; l < L; ++l) {
    const FlowOp* o = ops + 1 + per * l;
    const FlowOp& a = o[0];
    if (a.type != OP_LINEAR || a.src > BUF_X1 || a.dst < BUF_A0 || a.N != H || a.flags != 0 ||
        a.src_off != 0 || a.K < 1 || a.K > TC_TR0)
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
    if (c.type != OP_COUPLING_AFFINE || c.src != a.dst || c.K != H || c.d_id != a.K || c.d_tr < 1 ||
        2 * c.d_tr > TC_N3 || c.d_id + c.d_tr != D || c.N != 2 * c.d_tr)
      return 0;
    const int inv = (c.flags & FLAG_INVERSE) ? 1 : 0, add = (c.flags & FLAG_ADDITIVE) ? 1 : 0;
    if ((c.flags & ~(FLAG_INVERSE | FLAG_ADDITIVE)) != 0) return 0;
    if ((inverse >= 0 && inverse != inv) || (additive >= 0 && additive != add)) return 0;
    inverse = inv;
    additive = add;
    if (!is_affine(o[2 + 2 * NB])) return 0;
    t.d_id[l] = c.d_id;
    t.d_tr[l] = c.d_tr;
  }
  tc_put_overflow() = false;
  const bool narrow = H <= TC_H / 2;
  const RsLayout lay = rs_layout(NB, narrow);
  const int fixed = 2 * TC_AFF_BYTES + TC_ONES_BYTES + TC_ZERO_BYTES + 4096;
  int per_pass = (227 * 1024 - fixed) / (lay.layer_bytes + TC_AFF_BYTES);
  if (per_pass < 1) return 0;
  if (per_pass > L) per_pass = L;
  const int n_pass = (L + per_pass - 1) / per_pass;
  if (n_pass > RS_MAXPASS) return 0;
  per_pass = (L + n_pass - 1) / n_pass;  // balance the passes
  auto slot = [&](int layer, int j) {
    if (layer < 0 || layer >= L) return j;
    return j < t.d_id[layer] ? j : TC_TR0 + (j - t.d_id[layer]);
  };
  auto put_bias = [&](uint8_t* base, int n, float b) {
    if (!tc_h16_representable(b)) tc_put_overflow() = true;
    const uint16_t hi = tc_h16_rn(b);
    const uint16_t lo = tc_h16_rn(b - tc_h16_to_f(hi));
    memcpy(base + (size_t)n * 16, &hi, 2);
    memcpy(base + (size_t)n * 16 + 2, &lo, 2);
  };
  auto put_affine = [&](float* A, int i) {  // affine i: slots of layer i-1 -> slots of layer i
    const FlowOp& f = ops[per * i];
    for (int k = 0; k < D; ++k)
      for (int n = 0; n < D; ++n)
        A[slot(i - 1, k) * TC_DP + slot(i, n)] = blob[f.w_off + k * f.Npad + n];
    for (int n = 0; n < D; ++n) A[TC_DP * TC_DP + slot(i, n)] = blob[f.b_off + n];
  };
  for (int p = 0; p < n_pass; ++p) {
    const int l0 = p * per_pass, nl = (l0 + per_pass <= L ? per_pass : L - l0);
    const bool last = p == n_pass - 1;
    const int n_aff = nl + (last ? 1 : 0);
    const int ones_off = nl * lay.layer_bytes + n_aff * TC_AFF_BYTES;
    const int bytes = ones_off + TC_ONES_BYTES + TC_ZERO_BYTES;
    std::vector<uint8_t> img((size_t)bytes, 0);
    for (int m = 0; m < 128; ++m) {
      const uint16_t one[2] = {TC_ONE16, TC_ONE16};
      memcpy(img.data() + ones_off + (size_t)m * 16, one, 4);
    }
    for (int li = 0; li < nl; ++li) {
      const int l = l0 + li;
      uint8_t* lb = img.data() + (size_t)li * lay.layer_bytes;
      const FlowOp& f = ops[per * l];
      const FlowOp* o = ops + 1 + per * l;
      const FlowOp& a = o[0];
      // initial layer with the preceding affine folded in (float64): consumes the pre-affine state
      for (int n = 0; n < H; ++n) {
        for (int k = 0; k < D; ++k) {
          double acc = 0.0;
          for (int j = 0; j < a.K; ++j)
            acc += (double)blob[a.w_off + j * a.Npad + n] * (double)blob[f.w_off + k * f.Npad + j];
          tc_put(lb + lay.w0hi, lb + lay.w0lo, TC_N1, n, slot(l - 1, k), (float)acc);
        }
        double bacc = blob[a.b_off + n];
        for (int j = 0; j < a.K; ++j)
          bacc += (double)blob[a.w_off + j * a.Npad + n] * (double)blob[f.b_off + j];
        put_bias(lb + lay.b0, n, (float)bacc);
      }
      if (TC_AFFMMA) {
        // rows 64 .. 79 of G0's B operand: the affine itself (slots of layer l - 1 -> slots of layer l)
        for (int n = 0; n < D; ++n) {
          for (int k = 0; k < D; ++k)
            tc_put(lb + lay.w0hi, lb + lay.w0lo, TC_N1, TC_H + slot(l, n), slot(l - 1, k),
                   blob[f.w_off + k * f.Npad + n]);
          put_bias(lb + lay.b0, TC_H + slot(l, n), blob[f.b_off + n]);
        }
      }
      for (int b = 0; b < NB; ++b) {
        uint8_t* wb = lb + lay.blk + (size_t)b * 4 * lay.wbig;
        const FlowOp& x = o[1 + 2 * b];
        const FlowOp& y = o[2 + 2 * b];
        for (int n = 0; n < H; ++n)
          for (int k = 0; k < H; ++k) {
            tc_put(wb, wb + lay.wbig, lay.hrows, n, k, blob[x.w_off + k * x.Npad + n]);
            tc_put(wb + 2 * lay.wbig, wb + 3 * lay.wbig, lay.hrows, n, k, blob[y.w_off + k * y.Npad + n]);
          }
        for (int n = 0; n < H; ++n) {
          put_bias(lb + lay.bblk + (size_t)(2 * b) * lay.bias, n, blob[x.b_off + n]);
          put_bias(lb + lay.bblk + (size_t)(2 * b + 1) * lay.bias, n, blob[y.b_off + n]);
        }
      }
      const FlowOp& c = o[1 + 2 * NB];
      for (int n = 0; n < c.N; ++n) {
        for (int k = 0; k < H; ++k)
          tc_put(lb + lay.wfhi, lb + lay.wflo, TC_N3, n, k, blob[c.w_off + k * c.Npad + n]);
        put_bias(lb + lay.bf, n, blob[c.b_off + n]);
      }
    }
    float* aff = reinterpret_cast<float*>(img.data() + (size_t)nl * lay.layer_bytes);
    for (int li = 0; li < n_aff; ++li) put_affine(aff + (size_t)li * (TC_AFF_BYTES / 4), l0 + li);
    if (tc_put_overflow()) return 0;
    RsPass& P = t.pass[p];
    if (cudaMalloc(&P.d_image, bytes) != cudaSuccess) return 2;
    if (cudaMemcpy(P.d_image, img.data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess) return 2;
    P.image_bytes = bytes;
    P.l0 = l0;
    P.nl = nl;
  }
  t.L = L;
  t.D = D;
  t.NB = NB;
  t.n_pass = n_pass;
  t.inverse = inverse;
  t.additive = additive;
  t.narrow = narrow;
  t.act = activation;
  t.valid = true;
  return 0;
}

// ------------------------------------------------------------------ device
// One half (32 columns) of a 64-column accumulator -> (ReLU) -> split -> the matching half of
// the A operand.  Warp column-half c handles K chunks 2c and 2c + 1.
template <int ACT>
__device__ __forceinline__ void rs_hidden_half(uint32_t tg, int src_col, int c) {
  uint32_t ra[16], rb[16];
  tc_ld16(tg + src_col + 32 * c, ra);
  tc_ld16(tg + src_col + 32 * c + 16, rb);
  tc_wait_ld();
  tc_pin16(ra);
  tc_pin16(rb);
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
    tc_split_act2<ACT>(__uint_as_float(ra[2 * j]), __uint_as_float(ra[2 * j + 1]), hi[j], lo[j]);
  tc_st8(tg + RS_COL_AH + 8 * (2 * c), hi);
  tc_st8(tg + RS_COL_AL + 8 * (2 * c), lo);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    tc_split_act2<ACT>(__uint_as_float(rb[2 * j]), __uint_as_float(rb[2 * j + 1]), hi[j], lo[j]);
  tc_st8(tg + RS_COL_AH + 8 * (2 * c + 1), hi);
  tc_st8(tg + RS_COL_AL + 8 * (2 * c + 1), lo);
  tc_wait_st();
}

// Conditioner width <= 32: only accumulator columns 0 .. 31 carry hidden units (the rest is the
// zero padding of the weight image, never an MMA operand: the hidden GEMMs run K-steps 0 and 1).
// The twin warps split those 32 columns: warp column-half c handles K chunk c.
template <int ACT>
__device__ __forceinline__ void rs_hidden_quarter(uint32_t tg, int src_col, int c) {
  uint32_t ra[16];
  tc_ld16(tg + src_col + 16 * c, ra);
  tc_wait_ld();
  tc_pin16(ra);
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
    tc_split_act2<ACT>(__uint_as_float(ra[2 * j]), __uint_as_float(ra[2 * j + 1]), hi[j], lo[j]);
  tc_st8(tg + RS_COL_AH + 8 * c, hi);
  tc_st8(tg + RS_COL_AL + 8 * c, lo);
  tc_wait_st();
}
template <int ACT, bool NARROW>
__device__ __forceinline__ void rs_hidden_act(uint32_t tg, int src_col, int c) {
  if (NARROW) rs_hidden_quarter<ACT>(tg, src_col, c);
  else rs_hidden_half<ACT>(tg, src_col, c);
}
template <bool RELU, bool NARROW>
__device__ __forceinline__ void rs_hidden(uint32_t tg, int src_col, int c) {
  rs_hidden_act<RELU ? (int)ACT_RELU : TC_ACT_NONE, NARROW>(tg, src_col, c);
}

struct RsShared {
  uint64_t bar_in[RS_NG];
  uint64_t bar_out[RS_NG];
  uint64_t bar_img;  // completion of the weight image's bulk copy
  uint32_t tmem_base;
  uint32_t pad;
  double cst[4][TC_DP];
  double log_const;
};

__device__ __forceinline__ void rs_wait(uint32_t bar_out, uint32_t& ph) {
  tc_mbar_wait(bar_out, ph);
  ph ^= 1;
  tc_fence_after();
}
__device__ __forceinline__ void rs_arrive(uint32_t bar_in) {
  tc_fence_before();
  tc_mbar_arrive(bar_in);
}

// All layers of this pass for one row.  c == 0 threads own the row state h[]; c == 1 threads
// only help with the hidden epilogues.  Returns the row log|det J| accumulated in this pass.
template <bool NARROW, int ACT>
__device__ __forceinline__ float rs_run_row(const RsParams& P, const uint8_t* img, const RsLayout& lay,
                                            uint32_t tg, int c, uint32_t bar_in, uint32_t bar_out,
                                            uint32_t& ph, float (&h)[TC_DP]) {
  const float* aff = reinterpret_cast<const float*>(img + (size_t)P.nl * lay.layer_bytes);
  float ld = 0.f;
  for (int li = 0; li < P.nl; ++li) {
    const int d_tr = P.d_tr[P.l0 + li];
    if (c == 0) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) tc_split2<false>(h[2 * j], h[2 * j + 1], hi[j], lo[j]);
      tc_st8(tg + RS_COL_AH, hi);
      tc_st8(tg + RS_COL_AL, lo);
      tc_wait_st();
    }
    rs_arrive(bar_in);                                           // -> G0: D = W0' h + b0'
    if (c == 0 && !TC_AFFMMA) tc_affine(aff + (size_t)li * (TC_AFF_BYTES / 4), h);  // in the shadow of G0
    for (int b = 0; b < P.NB; ++b) {
      rs_wait(bar_out, ph);
      if (TC_AFFMMA && b == 0 && c == 0) {
        // the state after this layer's affine: G0's columns 64 .. 79 (the first columns of D2,
        // which the first block's GEMM overwrites only after this epilogue has arrived)
        uint32_t r[16];
        tc_ld16(tg + RS_COL_D2, r);
        tc_wait_ld();
        tc_pin16(r);
#pragma unroll
        for (int d = 0; d < TC_DP; ++d) h[d] = __uint_as_float(r[d]);
      }
      rs_hidden_act<ACT, NARROW>(tg, RS_COL_D, c);
      rs_arrive(bar_in);                                         // -> Ga: D2 = Wa relu(D) + ba
      rs_wait(bar_out, ph);
      rs_hidden_act<ACT, NARROW>(tg, RS_COL_D2, c);
      rs_arrive(bar_in);                                         // -> Gb: D += Wb relu(D2) + bb
    }
    rs_wait(bar_out, ph);
    rs_hidden_act<TC_ACT_NONE, NARROW>(tg, RS_COL_D, c);
    rs_arrive(bar_in);                                           // -> Gf: D2[0:16] = Wf D + bf
    rs_wait(bar_out, ph);
    if (c == 0) {
      uint32_t r[16];
      tc_ld16(tg + RS_COL_D2, r);
      tc_wait_ld();
      tc_pin16(r);
      ld += tc_coupling(r, h, d_tr, P.additive, P.inverse);
    }
  }
  if (c == 0 && P.last) tc_affine(aff + (size_t)P.nl * (TC_AFF_BYTES / 4), h);
  return ld;
}

template <int NKS>
__device__ __forceinline__ void rs_issuer_n(const RsParams& P, const RsLayout& lay, uint32_t img_s,
                                            uint32_t tg, uint32_t bar_in, uint32_t bar_out,
                                            int64_t my_tiles) {
  // hidden operands: 64 rows (N = 64), or 32 (N = 32) in the narrow instantiation
  constexpr int HROWS = NKS == 2 ? TC_H / 2 : TC_H;
  constexpr uint32_t ID64 = tc_idesc(128, HROWS), ID16 = tc_idesc(128, TC_N3);
  const uint32_t d = tg + RS_COL_D, d2 = tg + RS_COL_D2, ah = tg + RS_COL_AH, al = tg + RS_COL_AL;
  const int n_aff = P.nl + (P.last ? 1 : 0);
  const uint32_t ones_s = img_s + P.nl * lay.layer_bytes + n_aff * TC_AFF_BYTES;
  const uint32_t zero_s = ones_s + TC_ONES_BYTES;
  const uint64_t ones = tc_desc(ones_s, 2048, 128);
  auto adv = [](uint64_t desc, uint32_t off) { return desc + (uint64_t)(off >> 4); };
  auto bias = [&](uint32_t addr) { return tc_desc(addr, zero_s - addr, 128); };
  constexpr int nks = NKS;  // K-steps of 16 hidden units (4; 2 for a conditioner of width <= 32)
  // one K = 64 GEMM: acc0 = accumulate flag of the bias MMA
  auto gemm64 = [&](uint32_t dst, uint64_t bdesc, uint64_t whi, uint64_t wlo, uint32_t rows,
                    uint32_t idesc, uint32_t acc0) {
    tc_mma_ss_e(dst, ones, bdesc, idesc, acc0);
#pragma unroll
    for (int ks = 0; ks < nks; ++ks) {
      tc_mma_ts_e(dst, ah + 8 * ks, adv(whi, ks * 2 * rows * 16), idesc, 1);
      tc_mma_ts_e(dst, al + 8 * ks, adv(whi, ks * 2 * rows * 16), idesc, 1);
      tc_mma_ts_e(dst, ah + 8 * ks, adv(wlo, ks * 2 * rows * 16), idesc, 1);
    }
  };
  uint32_t ph = 0;
  for (int64_t it = 0; it < my_tiles; ++it) {
    for (int li = 0; li < P.nl; ++li) {
      const uint32_t lb = img_s + li * lay.layer_bytes;
      const uint64_t d64 = tc_desc(lb, HROWS * 16, 128);
      const uint64_t d16 = tc_desc(lb, TC_N3 * 16, 128);
      // G0
      tc_mbar_wait(bar_in, ph);
      ph ^= 1;
      tc_fence_after();
      constexpr uint32_t ID1 = tc_idesc(128, TC_N1);
      const uint64_t d1 = tc_desc(lb, TC_N1 * 16, 128);
      tc_mma_ss_e(d, ones, bias(lb + lay.b0), ID1, 0);
      tc_mma_ts_e(d, ah, adv(d1, lay.w0hi), ID1, 1);
      tc_mma_ts_e(d, al, adv(d1, lay.w0hi), ID1, 1);
      tc_mma_ts_e(d, ah, adv(d1, lay.w0lo), ID1, 1);
      tc_commit_e(bar_out);
      for (int b = 0; b < P.NB; ++b) {
        const uint32_t wb = lay.blk + b * 4 * lay.wbig;
        tc_mbar_wait(bar_in, ph);
        ph ^= 1;
        tc_fence_after();
        gemm64(d2, bias(lb + lay.bblk + 2 * b * lay.bias), adv(d64, wb), adv(d64, wb + lay.wbig), HROWS,
               ID64, 0);
        tc_commit_e(bar_out);
        tc_mbar_wait(bar_in, ph);
        ph ^= 1;
        tc_fence_after();
        gemm64(d, bias(lb + lay.bblk + (2 * b + 1) * lay.bias), adv(d64, wb + 2 * lay.wbig),
               adv(d64, wb + 3 * lay.wbig), HROWS, ID64, 1);
        tc_commit_e(bar_out);
      }
      tc_mbar_wait(bar_in, ph);
      ph ^= 1;
      tc_fence_after();
      gemm64(d2, bias(lb + lay.bf), adv(d16, lay.wfhi), adv(d16, lay.wflo), TC_N3, ID16, 0);
      tc_commit_e(bar_out);
    }
  }
}

__device__ __forceinline__ int64_t rs_my_tiles(int64_t ntiles, int g) {
  const int64_t first = (int64_t)blockIdx.x * RS_NG + g;
  const int64_t stride = (int64_t)gridDim.x * RS_NG;
  return first < ntiles ? (ntiles - first + stride - 1) / stride : 0;
}

// MODE 0: apply (rows supplied), MODE 1: populate (Philox draw in the first pass, float64 tail
// in the last).  A must be valid for MODE 1, io for MODE 0.
// NARROW: conditioner width <= 32, a compile-time switch (see flow_tc.cuh).
template <int MODE, bool NARROW, int ACT>
__global__ void __launch_bounds__(RS_THREADS, 1) flow_tc_res_kernel(RsParams P, TcIO io, PopulateArgs A) {
  extern __shared__ __align__(1024) uint8_t rs_smem[];
  RsShared* sh = reinterpret_cast<RsShared*>(rs_smem + tc_image_pad(P.image_bytes));
  const int tid = threadIdx.x;
  const RsLayout lay = rs_layout(P.NB, NARROW);
  if (MODE == 1 && P.last) {
    if (tid < 4 * TC_DP) {
      const int which = tid / TC_DP, d = tid % TC_DP;
      const double* src = which == 0 ? A.scale : which == 1 ? A.shift : which == 2 ? A.lo : A.hi;
      sh->cst[which][d] = d < P.D ? src[d] : 0.0;
    }
    if (tid == 4 * TC_DP) sh->log_const = populate_log_const(A, P.D);
  }
  if (tid == 0) {
    tc_image_load(tc_smem_u32(rs_smem), P.image, (uint32_t)P.image_bytes, tc_smem_u32(&sh->bar_img));
    for (int g = 0; g < RS_NG; ++g) {
      tc_mbar_init(tc_smem_u32(&sh->bar_in[g]), RS_EW * 32);
      tc_mbar_init(tc_smem_u32(&sh->bar_out[g]), 1);
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
    const uint32_t tg = tmem + g * RS_COLS + ((uint32_t)(q * 32) << 16);
    const uint32_t bar_in = tc_smem_u32(&sh->bar_in[g]), bar_out = tc_smem_u32(&sh->bar_out[g]);
    uint32_t ph = 0;
    double vmax = -INFINITY, vcount = 0.0;
    const int64_t stride = (int64_t)gridDim.x * RS_NG;
    for (int64_t tile = (int64_t)blockIdx.x * RS_NG + g; tile < ntiles; tile += stride) {
      const int64_t row = tile * 128 + q * 32 + (tid & 31);
      const bool valid = row < n;
      float h[TC_DP];
      float ss = 0.f, ld0 = 0.f;
      bool alive = true;
#pragma unroll
      for (int d = 0; d < TC_DP; ++d) h[d] = 0.f;
      if (c == 0) {
        if (!P.first) {
          if (valid) {
            const float4* p4 = reinterpret_cast<const float4*>(P.sc_h + row * TC_DP);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float4 v = p4[k];
              h[4 * k] = v.x, h[4 * k + 1] = v.y, h[4 * k + 2] = v.z, h[4 * k + 3] = v.w;
            }
            ld0 = P.sc_ld[row];
            const float e = P.sc_ss[row];
            alive = e >= 0.f;
            ss = alive ? e : -e - 1.f;
          }
        } else if (MODE == 0) {
          if (P.D == TC_DP) {
            const float4* p4 = reinterpret_cast<const float4*>(io.in + row * TC_DP);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float4 v = valid ? __ldg(p4 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
              h[4 * k] = v.x, h[4 * k + 1] = v.y, h[4 * k + 2] = v.z, h[4 * k + 3] = v.w;
            }
          } else {
#pragma unroll
            for (int d = 0; d < TC_DP; ++d) h[d] = (valid && d < P.D) ? __ldg(io.in + row * P.D + d) : 0.f;
          }
#pragma unroll
          for (int d = 0; d < TC_DP; ++d) ss = fmaf(h[d], h[d], ss);
        } else {
#pragma unroll
          for (int d0 = 0; d0 < TC_DP; d0 += 4) {
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
              if (A.z && use && valid) A.z[row * P.D + d0 + j] = h[d0 + j];
            }
          }
          const float rad = sqrtf(ss) * A.sqrt_t;
          alive = !(A.r_max > 0.f) || (rad <= A.r_max);
        }
      }
      const float ld = ld0 + rs_run_row<NARROW, ACT>(P, rs_smem, lay, tg, c, bar_in, bar_out, ph, h);
      if (c != 0) continue;
      if (!P.last) {
        if (valid) {
          float4* o4 = reinterpret_cast<float4*>(P.sc_h + row * TC_DP);
#pragma unroll
          for (int k = 0; k < 4; ++k) o4[k] = make_float4(h[4 * k], h[4 * k + 1], h[4 * k + 2], h[4 * k + 3]);
          P.sc_ld[row] = ld;
          P.sc_ss[row] = alive ? ss : -ss - 1.f;
        }
      } else if (MODE == 0) {
        const float logj = ld + P.const_logdet;
        float ss_out = 0.f;
#pragma unroll
        for (int d = 0; d < TC_DP; ++d) ss_out = d < P.D ? fmaf(h[d], h[d], ss_out) : ss_out;
        if (valid) {
          if (io.out) {
            if (P.D == TC_DP) {
              float4* o4 = reinterpret_cast<float4*>(io.out + row * TC_DP);
#pragma unroll
              for (int k = 0; k < 4; ++k) o4[k] = make_float4(h[4 * k], h[4 * k + 1], h[4 * k + 2], h[4 * k + 3]);
            } else {
#pragma unroll
              for (int d = 0; d < TC_DP; ++d)
                if (d < P.D) io.out[row * P.D + d] = h[d];
            }
          }
          if (io.out_logj) io.out_logj[row] = logj;
          if (io.out_lp) {
            const float cn = io.base_log_z, hv = 0.5f * io.base_inv_var;
            io.out_lp[row] = (io.lp_mode == 1) ? (-hv * ss - cn) - logj : (-hv * ss_out - cn) + logj;
          }
        }
      } else {
        const float logj = ld + P.const_logdet;
        const float base_lp = -0.5f * ss - 0.5f * P.D * TC_LOG_2PI;
        populate_row<TC_DP>(A, P.D, [&](int d) { return h[d]; }, row, alive, base_lp, logj, vmax, vcount,
                            sh->cst[0], sh->cst[1], sh->cst[2], sh->cst[3], sh->log_const);
      }
    }
    if (MODE == 1 && P.last && c == 0) populate_publish(A, vmax, vcount);
  } else {
    const int g = __shfl_sync(0xffffffffu, warp - RS_NG * RS_EW, 0);
    rs_issuer_n<NARROW ? 2 : 4>(P, lay, tc_smem_u32(rs_smem), __shfl_sync(0xffffffffu, tmem, 0) + g * RS_COLS,
              tc_smem_u32(&sh->bar_in[g]), tc_smem_u32(&sh->bar_out[g]), rs_my_tiles(ntiles, g));
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

inline size_t rs_smem_bytes(int image_bytes) {
  return (((size_t)image_bytes + 1023) & ~(size_t)1023) + sizeof(RsShared) + 64;
}
inline int rs_grid(int64_t n, int num_sms) {
  const int64_t ntiles = (n + 127) / 128;
  const int64_t want = (ntiles + RS_NG - 1) / RS_NG;
  return (int)(want < num_sms ? want : num_sms);
}
inline int rs_reserve(RsProgram& t, int64_t n) {
  if (t.n_pass <= 1 || n <= t.scratch_rows) return 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return 1;
  if (t.d_scratch) cudaFree(t.d_scratch);
  t.d_scratch = nullptr;
  t.scratch_rows = 0;
  if (cudaMalloc(&t.d_scratch, sizeof(float) * (size_t)n * (TC_DP + 2)) != cudaSuccess) return 1;
  t.scratch_rows = n;
  return 0;
}

// Runs every pass; returns the number of kernel launches (0 on failure).
template <int MODE>
inline int rs_launch(RsProgram& t, const TcIO& io, const PopulateArgs& A, int64_t n, int num_sms,
                     cudaStream_t st) {
  if (rs_reserve(t, n)) return 0;
  for (int p = 0; p < t.n_pass; ++p) {
    const RsPass& ps = t.pass[p];
    const size_t smem = rs_smem_bytes(ps.image_bytes);

    RsParams P;
    P.image = ps.d_image;
    P.image_bytes = ps.image_bytes;
    P.l0 = ps.l0;
    P.nl = ps.nl;
    P.L = t.L;
    P.D = t.D;
    P.NB = t.NB;
    for (int i = 0; i < TC_MAXL; ++i) P.d_id[i] = t.d_id[i], P.d_tr[i] = t.d_tr[i];
    P.additive = t.additive;
    P.inverse = t.inverse;
    P.first = p == 0;
    P.last = p == t.n_pass - 1;
    P.narrow = t.narrow;
    P.const_logdet = t.const_logdet;
    P.sc_h = t.d_scratch;
    P.sc_ld = t.d_scratch ? t.d_scratch + (size_t)t.scratch_rows * TC_DP : nullptr;
    P.sc_ss = t.d_scratch ? t.d_scratch + (size_t)t.scratch_rows * (TC_DP + 1) : nullptr;
#define NB200_RS_LAUNCH(NARROW, ACT)                                                              \
  {                                                                                                 \
    if (tc_prep((const void*)flow_tc_res_kernel<MODE, NARROW, ACT>, smem)) return 0;                \
    flow_tc_res_kernel<MODE, NARROW, ACT><<<rs_grid(n, num_sms), RS_THREADS, smem, st>>>(P, io, A); \
  }
    NB200_TC_DISPATCH(NB200_RS_LAUNCH, t.narrow, t.act)
#undef NB200_RS_LAUNCH
    if (cudaGetLastError() != cudaSuccess) return 0;
  }
  return t.n_pass;
}

}  // namespace nb200
