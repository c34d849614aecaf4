// TEST HARNESS ONLY: one optimisation step of the fused training kernels (nessai_b200/csrc/train.cuh:
// FWD(0..L-1) -> LOSS -> BWD(L-1..0) -> REDUCE -> ADAM, the CUDA source, unchanged apart from the
// NB200_SIMT_SHIM flavour of its cp.async helpers and of the dynamic shared-memory symbol) under
// the CPU SIMT shim.  The launch sequence and the workspace sizes mirror nb200_trainer_create /
// nb200_train_epoch of nessai_b200.cu for ONE batch.  Compile with -I tests/_hostcheck/fake_cuda.
#define NB200_SIMT_SHIM 1
#define nb200 nb200_simt_train
#include "simt_shim.h"

inline float __int_as_float(int v) {
  float f;
  std::memcpy(&f, &v, sizeof f);
  return f;
}

#include "../../nessai_b200/csrc/train.cuh"

// One gradient / optimiser step on the n_rows rows of x.  opt_kind: 0 AdamW, 1 Adam, 2 SGD,
// -1 gradient only.  grad_out[n_params]; loss_accum[1] += batch loss; step_info[2] = {loss, |g|}.
extern "C" int simt_train_step(const int32_t* h_plan, int n_plan_ints, const int32_t* itab, int n_itab,
                               const int32_t* reduce_idx, int n_reduce, float* theta_p, float* theta_b, float* m,
                               float* v, const float* x, const float* w, int n_rows, int opt_kind, double lr,
                               double beta1, double beta2, double eps, double weight_decay, double clip,
                               int64_t step0, const float* pmask, float* loss_accum, float* step_info,
                               float* grad_out, int num_sms) {
  using namespace nb200;
  if (n_plan_ints != TR_PLAN_INTS) return 1;
  TrPlan P;
  std::memcpy(&P, h_plan, sizeof(TrPlan));
  if (P.n_itab != n_itab || P.n_reduce != n_reduce) return 2;
  const size_t smem_fwd = 4 * tr_smem_floats(P.D, P.vals_floats, P.max_in, P.wmax, P.n_itab, false);
  const size_t smem_bwd = 4 * tr_smem_floats(P.D, P.vals_floats, P.max_in, P.wmax, P.n_itab, true);
  if (smem_bwd > sizeof(simt::dynamic_smem) || smem_fwd > sizeof(simt::dynamic_smem)) return 5;
  const int Gmax = TR_MAXG;
  const int n_tiles = (n_rows + TR_R - 1) / TR_R;
  std::vector<float> ws((size_t)n_tiles * P.rec_total * TR_R), dout0((size_t)n_tiles * P.D * TR_R),
      dout1((size_t)n_tiles * P.D * TR_R), ldrow((size_t)n_tiles * TR_R), crow((size_t)n_tiles * TR_R),
      stat_part((size_t)P.L * Gmax * 2 * P.D), stats((size_t)P.L * 2 * P.D), s_part0((size_t)Gmax * 2 * P.D),
      s_part1((size_t)Gmax * 2 * P.D), wsum_part(Gmax), loss_part(Gmax), part((size_t)Gmax * ((P.n_part + 3) & ~3), 0.f),
      grad((P.n_params + 3) & ~3, 0.f), gn_part(Gmax), stat_n(Gmax);
  TrBatch bt;
  bt.x = x, bt.perm = nullptr, bt.w = w, bt.i0 = 0, bt.B = n_rows, bt.n_tiles = n_tiles;
  const int G = std::min(n_tiles, std::min(TR_MAXG, num_sms));
  TrBuffers B;
  B.itab = itab, B.reduce_idx = reduce_idx, B.theta_p = theta_p, B.theta_b = theta_b;
  B.ws = ws.data(), B.dout[0] = dout0.data(), B.dout[1] = dout1.data(), B.ldrow = ldrow.data(), B.crow = crow.data();
  B.stat_part = stat_part.data(), B.stat_n = stat_n.data(), B.stats = stats.data();
  B.s_part[0] = s_part0.data(), B.s_part[1] = s_part1.data(), B.wsum_part = wsum_part.data();
  B.loss_part = loss_part.data(), B.part = part.data(), B.grad = grad.data(), B.gn_part = gn_part.data();
  B.G = G, B.pmask = pmask;
  B.part_stride = (P.n_part + 3) & ~3;
  B.n_reduce_blocks = G;
  if (pmask)
    simt_launch(tr_mask_params_kernel, (unsigned)std::min((P.n_params + 255) / 256, 2 * num_sms), 256u, theta_p, pmask,
                P.n_params);
  for (int l = 0; l < P.L; ++l) simt_launch(tr_fwd_kernel, (unsigned)G, (unsigned)TR_THREADS, P, B, bt, l);
  simt_launch(tr_loss_kernel, (unsigned)G, (unsigned)TR_THREADS, P, B, bt);
  for (int l = P.L - 1; l >= 0; --l) simt_launch(tr_bwd_kernel, (unsigned)G, (unsigned)TR_THREADS, P, B, bt, l);
  simt_launch(tr_reduce_kernel, (unsigned)G, (unsigned)TR_THREADS, P, B);
  const int64_t step = step0 + 1;
  TrOptim o;
  o.kind = opt_kind, o.lr = (float)lr, o.beta1 = (float)beta1, o.beta2 = (float)beta2, o.eps = (float)eps;
  o.weight_decay = (float)weight_decay, o.clip = (float)clip;
  o.bc1 = (float)(1.0 - std::pow(beta1, (double)step));
  o.bc2 = (float)(1.0 - std::pow(beta2, (double)step));
  simt_launch(tr_adam_kernel, (unsigned)std::min((P.n_params + 255) / 256, 2 * num_sms), 256u, B, P.n_params, o, m, v,
              step_info, loss_accum);
  std::memcpy(grad_out, grad.data(), sizeof(float) * P.n_params);
  return 0;
}

// The validation loss (flowmodel/base.py:454-523): eval mode, running statistics, all layers of a
// tile in one pass; mirrors nb200_eval_loss.  d_loss[1] and / or logp[n].
extern "C" int simt_eval_loss(const int32_t* h_plan, int n_plan_ints, const int32_t* itab, int n_itab,
                              const int32_t* reduce_idx, int n_reduce, float* theta_p, float* theta_b, const float* x,
                              const float* w, int n, const float* pmask, float* d_loss, float* logp, int num_sms) {
  using namespace nb200;
  if (n_plan_ints != TR_PLAN_INTS) return 1;
  TrPlan P;
  std::memcpy(&P, h_plan, sizeof(TrPlan));
  if (P.n_itab != n_itab || P.n_reduce != n_reduce) return 2;
  TrBatch bt;
  bt.x = x, bt.perm = nullptr, bt.w = w, bt.i0 = 0, bt.B = n, bt.n_tiles = (n + TR_R - 1) / TR_R;
  const int G = std::min(bt.n_tiles, 2 * num_sms);
  std::vector<float> eval_part(2 * 2 * num_sms);
  TrBuffers B{};
  B.itab = itab, B.reduce_idx = reduce_idx, B.theta_p = theta_p, B.theta_b = theta_b, B.G = G, B.pmask = pmask;
  if (pmask)
    simt_launch(tr_mask_params_kernel, (unsigned)std::min((P.n_params + 255) / 256, 2 * num_sms), 256u, theta_p, pmask,
                P.n_params);
  simt_launch(tr_eval_kernel, (unsigned)G, (unsigned)TR_THREADS, P, B, bt, eval_part.data(), logp);
  if (d_loss) simt_launch(tr_eval_final_kernel, 1u, 32u, (const float*)eval_part.data(), G, d_loss);
  return 0;
}

extern "C" double nb200_host_erfcinv(double) { return 0.0; }  // declared by the shim; unused here
