// Folded flow program shared by every kernel (see nessai_b200/spec.py).
//
// The reference walks a torch Module tree per call
// (/root/reference/src/nessai/flows/base.py:209-221 -> glasflow.nflows
// CompositeTransform); here the eval-mode flow is a flat list of ops over four
// per-row column buffers (X0, X1: features; A0, A1: conditioner activations).
#pragma once
#include <cstdint>

namespace nb200 {

enum : int { OP_LINEAR = 0, OP_COUPLING_AFFINE = 1, OP_COUPLING_SPLINE = 2 };
enum : int {
  FLAG_IN_ACT = 1,    // activation applied to the op's input
  FLAG_OUT_ACT = 2,   // activation applied to the op's output
  FLAG_ACCUM = 4,     // dst += (residual connection)
  FLAG_INVERSE = 8,   // coupling applied in the inverse direction
  FLAG_ADDITIVE = 16,       // volume preserving coupling (scale == 1)
  FLAG_SOFTPLUS_SCALE = 32, // scale = softplus(u) + 1e-3 (nflows MaskedAffineAutoregressiveTransform)
  FLAG_NO_LOGDET = 64       // do not accumulate log|det| (intermediate passes of the MAF inverse)
};
enum : int { BUF_X0 = 0, BUF_X1 = 1, BUF_A0 = 2, BUF_A1 = 3 };
enum : int { ACT_RELU = 0, ACT_TANH = 1, ACT_SILU = 2 };

// 16 ints; must match OP_INTS / the encoding in spec.py::FoldedFlow.program
struct FlowOp {
  int type, src, dst, src_off;
  int K, N, Npad, w_off;
  int b_off, flags, x_buf, d_id;
  int d_tr, e0, e1, e2;  // spline: e0 = bins, e1 = group stride, e2 = tail bound (float bits)
};
static_assert(sizeof(FlowOp) == 64, "FlowOp layout");

struct FlowProgramDev {
  const FlowOp* ops;   // device
  const float* blob;   // device
  int n_ops;
  int D, Dpad, Hpad;
  int activation;
  int final_buf;
  int wmax;            // floats of the largest staged weight block (K*Npad + Npad)
  float const_logdet;
};

}  // namespace nb200
