/* nessai_b200 -- C ABI of the B200-native flow-proposal hot path.
 *
 * The reference (mj-will/nessai) has no FFI: its boundary for this path is the
 * Python class nessai.flowmodel.FlowModel (plugin point P2).  The
 * entry points below are what a binding for that class calls; each cites the
 * reference method it replaces (paths relative to /root/reference/src/nessai).
 *
 * Conventions: plain pointers and sizes only.  Pointers named d_* are DEVICE
 * pointers (fp32 row-major unless stated), h_* are HOST pointers.  `stream` is a
 * cudaStream_t passed as void* (NULL = default stream).  Every function returns 0
 * on success and a non-zero code otherwise; nb200_last_error() describes the last
 * failure of the calling thread.  Per-row numerical failures (spline
 * discriminant < 0, overflow) never trap: the row's outputs are NaN, which the
 * caller drops exactly as flowproposal/flowproposal.py:366-368 drops non-finite
 * log-probabilities.
 */
#ifndef NESSAI_B200_H
#define NESSAI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nb200_flow nb200_flow; /* opaque: one eval-mode flow on one GPU */

int nb200_version(void);
const char* nb200_last_error(void);
/* number of kernel launches issued by this library since load / last reset */
int64_t nb200_launch_count(void);
void nb200_reset_launch_count(void);
/* Enable (1, default) / disable (0) the tcgen05 specialisation for the flow shapes it
 * covers; disabled, every flow runs the generic fp32 kernel.  Returns the previous
 * setting.  (Environment: NB200_DISABLE_TC=1 sets the initial value to 0.) */
int nb200_set_tensor_core_path(int enabled);

/* flows/utils.py:208-246 configure_model -> a device-resident flow object.
 * D = features, H = conditioner width, activation: 0 relu, 1 tanh, 2 silu. */
int nb200_flow_create(nb200_flow** out, int D, int H, int activation);
int nb200_flow_destroy(nb200_flow* flow);

/* Upload the folded eval-mode program for one direction (0 = forward x->z,
 * 1 = inverse z->x): replaces model.eval() + the LU-cache rebuild of
 * flowmodel/base.py:680-690.  h_ops: int32[n_ops*16] (csrc/flow_program.h),
 * h_blob: float32[n_blob]; both are copied. */
int nb200_flow_set_program(nb200_flow* flow, int direction, const int32_t* h_ops, int n_ops,
                           const float* h_blob, int64_t n_blob, int final_buf,
                           double const_logdet);

/* flowmodel/base.py:813-840 FlowModel.inverse and :939-944
 * sample_and_log_prob(z=z):  d_x[n*D], d_logj[n] = log|det dx/dz|,
 * d_logq[n] = log N(z) - logj.  d_logj / d_logq may be NULL. */
int nb200_flow_inverse(nb200_flow* flow, const float* d_z, float* d_x, float* d_logj,
                       float* d_logq, int64_t n, void* stream);

/* flowmodel/base.py:782-811 forward_and_log_prob and :842-864 log_prob:
 * d_z[n*D] (may be NULL), d_logj[n] (may be NULL), d_logp[n] = log N(z) + logj. */
int nb200_flow_forward(nb200_flow* flow, const float* d_x, float* d_z, float* d_logj,
                       float* d_logp, int64_t n, void* stream);

/* Base distribution of the flow: N(0, var I), nessai's MultivariateNormal
 * (flows/distributions.py:17-73, flow_config["distribution"] = "mvn" / "normal",
 * flows/utils.py:35-102); var = 1 (the default) is nflows' StandardNormal.  Enters every
 * log-probability (-0.5 |z|^2 / var - 0.5 D log(2 pi var)) and the draws of
 * nb200_populate_draw (z = sqrt(T var) v). */
int nb200_flow_set_base_variance(nb200_flow* flow, double var);

/* flowmodel/base.py:889-904 sample_latent_distribution: d_z[n*D] ~ N(0, std_dev^2 I),
 * Philox4x32-10 keyed by `seed`, counter = row_offset + row. */
int nb200_sample_latent(float* d_z, int64_t n, int D, uint64_t seed, uint64_t row_offset,
                        double std_dev, void* stream);

/* One turn of the populate() loop, flowproposal/flowproposal.py:431-469, fused:
 * draw z (Philox; * sqrt_temperature) -> latent-radius truncation
 * (truncation.py:358-365: keep |z| <= r_max, r_max <= 0 disables) -> inverse flow
 * -> log_q = log N(z / sqrt T) - D log sqrt T - logj (base.py:401-414) ->
 * diagonal inverse rescale x = x' * scale + shift, log_q -= sum log|scale|
 * (reparameterisations/rescale.py:263-291) -> prior-bounds check
 * (model.py:497-518) -> log_w = log_prior_const - log_q (base.py:1069-1098,
 * uniform box prior) -> running max of log_w over valid rows.
 *
 * d_scale, d_shift, d_lo, d_hi: float64[D] device.  Outputs: d_xp float32[n*D] (the
 * flow output x' BEFORE the rescale; x = (double)x' * scale + shift is formed in
 * float64 for the bounds check here and again, identically, for the accepted
 * rows in nb200_populate_accept, so dropped rows never cost a float64 write),
 * d_logq float64[n], d_logw float64[n] (NaN for dropped rows), d_z float32[n*D]
 * or NULL, d_stats: float64[2] = {max log_w (init -inf by caller), n_valid
 * (accumulated)}.  If log_prior_const is NaN, d_logw holds -log_q and the caller
 * adds its own prior.  min_log_q: rows with log_q <= min_log_q are dropped
 * (MinLogQTruncation.apply_after_backward, proposal/flowproposal/truncation.py:388-394);
 * -inf or NaN keeps every row. */
int nb200_populate_draw(nb200_flow* flow, int64_t n, uint64_t seed, uint64_t row_offset,
                        float r_max, float sqrt_temperature, const double* d_scale,
                        const double* d_shift, const double* d_lo, const double* d_hi,
                        double log_prior_const, double min_log_q, float* d_xp, double* d_logq,
                        double* d_logw, float* d_z, double* d_stats, void* stream);

/* Rejection step + compaction, flowproposal/flowproposal.py:491-498:
 * accept = (log_w - max) > log(u), u ~ U(0,1) (Philox, `seed`, counter =
 * row_offset + row); accepted rows are written IN DRAW ORDER as structured
 * live-point records (x = (double)x' * scale + shift): each record starts as a copy of d_row_template
 * (row_bytes, multiple of 4) and gets the D parameters (float64 at byte offsets
 * h_field_offsets[0..D-1]) and logP (float64 at h_field_offsets[D], skipped if
 * negative) overwritten.  At most `capacity` records are written to d_rows;
 * d_counts: int64[2] = {n_accepted (all), n_written}.  d_max points at the
 * (possibly all-reduced) maximum of log_w.  d_scratch: int64[ceil(n/1024)+1].
 * d_logl (float64[n] or NULL): per-row log-likelihood evaluated inside the loop
 * (LikelihoodThresholdTruncation, flowproposal.py:456-460), written to the
 * record's logL field at byte offset logl_offset (skipped if NULL / negative).
 * One single-pass kernel: decoupled look-back scan over 1024-row chunks. */
int nb200_populate_accept(int64_t n, int D, const float* d_xp, const double* d_scale,
                          const double* d_shift, const double* d_logw, const double* d_logl,
                          const double* d_max, uint64_t seed, uint64_t row_offset,
                          double log_p_value, const uint8_t* d_row_template, int row_bytes,
                          const int32_t* h_field_offsets, int logl_offset, uint8_t* d_rows,
                          int64_t capacity, int64_t write_offset, int64_t* d_counts,
                          int64_t* d_scratch, void* stream);

/* The same rejection step + compaction for a reparameterisation that is not a diagonal
 * affine: the physical rows x (float64[n*D]) were formed by nb200_reparam_tail and are copied
 * into the records as they are.  Everything else as nb200_populate_accept. */
int nb200_populate_accept_x64(int64_t n, int D, const double* d_x64, const double* d_logw,
                              const double* d_logl, const double* d_max, uint64_t seed,
                              uint64_t row_offset, double log_p_value,
                              const uint8_t* d_row_template, int row_bytes,
                              const int32_t* h_field_offsets, int logl_offset, uint8_t* d_rows,
                              int64_t capacity, int64_t write_offset, int64_t* d_counts,
                              int64_t* d_scratch, void* stream);

/* Populate tail for per-parameter maps x = h(a x' + b) * scale + shift with h = identity
 * (kind 0), sigmoid (1), |.| (2), exp (3), log (4), the standard normal CDF (5) or its
 * quantile function (6): the inverse direction of reparameterisations/rescale.py:263-291
 * (ScaleAndShift) and :635-660 (RescaleToBounds.inverse_reparameterise) with the rescaling
 * functions of utils/rescaling.py:290-417 as post-rescaling ("logit" -> sigmoid, log|J| =
 * log h + log1p(-h); "log" -> exp, log|J| = u; "exp" -> log; "gaussian_cdf" -> quantile;
 * "inv_gaussian_cdf" -> CDF), as pre-rescaling (then the affine map comes first: d_pre_scale /
 * d_pre_shift = a / b, log|J| += log|a|, scale = 1, shift = 0) or boundary inversion
 * (rescale.py:570-590: |x'|, an "upper" edge folded into a negative scale), then the affine
 * map to the prior bounds (rescale.py:544-553, log|J| = log|scale|).  Runs AFTER
 * nb200_populate_draw called with scale = 1, shift = 0, lo = -inf, hi = +inf,
 * min_log_q = -inf: d_xp float32[n*D] is the flow output, d_logq float64[n] the flow's own
 * log q (NaN: dropped).  Rewrites d_logq (-= log|J|), d_logw (= log_prior_const - log_q; NaN
 * for rows outside d_lo/d_hi, with non-finite log_q or log_q <= min_log_q), writes
 * d_x64 float64[n*D], and accumulates d_stats = {max log_w, n_valid} (reset by the caller).
 * Pair kinds, reparameterisations/angle.py:17-186 (Angle.inverse_reparameterise): output slot d
 * reads TWO flow features (u0, u1) = (x'[d_src[3d]], x'[d_src[3d+1]]): kind 7 the angle
 * atan2(u1, u0) * scale + shift (scale = 1 / Angle.scale; no log|J| for that constant, as in the
 * reference), 8 the same modulo 2 pi (prior starting at zero, angle.py:158-166), 9 the radius
 * sqrt(u0^2 + u1^2) with log|J| -= log r (:172), 10 the same for an AUXILIARY radius whose chi(2)
 * prior log r - r^2/2 (:183-185) is added to log_w.
 * Kind 11 (single feature): floor(u), no log|J| -- Dequantise's pre-rescaling inverse
 * (reparameterisations/discrete.py:66-78).  Kind 12, ToCartesian (angle.py:189-232):
 * |atan2(u1, u0) * a| * scale + shift with a = d_pre_scale[d] = 1 / ToCartesian.scale and
 * scale = hi - lo, shift = lo of the prior bounds, log|J| += log(hi - lo).
 * Triple kinds, AnglePair (angle.py:235-538): (u0, u1, u2) = the Cartesian (x, y, z); the horizontal
 * angle is kind 7 / 8 of (u0, u1); 13 the zenith atan2(sqrt(u0^2 + u1^2), u2) with
 * log|J| -= log sin(.), 14 the declination atan2(u2, sqrt(u0^2 + u1^2)) with log|J| -= log cos(.),
 * 15 the radius sqrt(u0^2 + u1^2 + u2^2) with log|J| -= 2 log r, 16 the same for an auxiliary
 * radius whose chi(3) prior (:529-537) is added to log_w.
 * A single-feature kind OR-ed with 0x100: floor(.) of the final value, x = floor(h(a x' + b) * scale
 * + shift), no log|J| -- Dequantise with a post-rescaling ("dequantise-logit").
 * Kind 17 (single feature): an augment parameter of AugmentedFlowProposal (proposal/augmented.py:
 * 150-178), passed through unchanged, whose N(0, 1) log-density is added to log_w.
 * d_src int32[3*D] or NULL (slot d reads feature d): the flow features output slot d reads; the
 * single-feature kinds use d_src[3d] only, the pair kinds d_src[3d], d_src[3d+1].
 * d_kind int32[D], d_scale/d_shift/d_lo/d_hi float64[D] on the device; d_pre_scale /
 * d_pre_shift float64[D] or both NULL (a = 1, b = 0); D <= 64. */
int nb200_reparam_tail(int64_t n, int D, const float* d_xp, const int32_t* d_kind,
                       const int32_t* d_src, const double* d_pre_scale, const double* d_pre_shift,
                       const double* d_scale, const double* d_shift, const double* d_lo,
                       const double* d_hi, double log_prior_const, double min_log_q,
                       double* d_logq, double* d_logw, double* d_x64, double* d_stats,
                       void* stream);

/* accumulate_weights variant, flowproposal/flowproposal.py:471-490: the expected pool size is
 * exp(logsumexp(log_weights - log_constant)).  d_partials[b] (float64[n_partials]) receives
 * block b's share of sum_i exp(d_logw[i] - *d_max) over the non-NaN rows; the caller adds them
 * in index order (no atomics: the sum is reproducible). */
int nb200_sum_exp(const double* d_logw, int64_t n, const double* d_max, double* d_partials,
                  int n_partials, void* stream);

/* glasflow.nflows AffineCouplingTransform._coupling_transform_forward / _inverse (the
 * element-wise stage of flows/realnvp.py:110-112 with the conditioner output supplied):
 * d_params[n][2*d_tr] = shift | unconstrained scale (nflows layout; [n][d_tr] when additive),
 * scale = sigmoid(u + 2) + 1e-3; transformed features y = x*scale + shift (inverse:
 * (x - shift)/scale), identity features copied, d_logdet[n] = +-sum log scale.
 * h_transform_features: the d_tr feature indices (host int32), D <= 64. */
int nb200_coupling_transform(const float* d_x, const float* d_params, float* d_y, float* d_logdet,
                             int64_t n, int D, const int32_t* h_transform_features, int d_tr,
                             int additive, int inverse, void* stream);

/* ------------------------------------------------------------------ peer-memory exchange
 * The scalar exchanges of a populate turn sharded over the GPUs of ONE node --
 * the max log-weight of the turn before the rejection step (flowproposal.py:491-494 normalises by
 * the maximum over the whole turn) and every rank's {accepted, written} counts after it -- as
 * direct NVLink stores into the peers' slot buffers (CUDA IPC), polled locally: no NCCL
 * collective, no host round trip.  create: this rank's slot buffer + its 64-byte IPC handle
 * (exchanged by the caller); open: a peer's buffer mapped into this process.
 * allgather: publish n_words (<= 3) 64-bit words from d_src to every rank, collect every rank's
 * into d_gathered[world][n_words] (may be NULL); d_max_out (may be NULL) = max over the ranks of
 * word 0 read as a double.  kind: 0 max, 1 counts; seq: 1, 2, ... the same on every rank for
 * the same exchange.  d_err is set to 1 if a peer did not answer within the spin limit. */
int nb200_xchg_create(void** d_buf, unsigned char* handle64);
int nb200_xchg_open(const unsigned char* handle64, void** d_peer);
int nb200_xchg_close(void* d_peer);
int nb200_xchg_destroy(void* d_buf);
int nb200_xchg_allgather(const void* const* d_peers, int world, int rank, int kind, uint64_t seq,
                         const void* d_src, int n_words, void* d_gathered, double* d_max_out,
                         int* d_err, void* stream);

/* ------------------------------------------------------------------ training
 * flowmodel/base.py:365-452 FlowModel._train: one EPOCH of optimisation steps on a
 * RealNVP flow, fused: train-mode forward (batch-statistics BatchNorm with running-
 * statistics EMA, uncached LU), loss = -mean(log_prob) (or the weighted loss of
 * base.py:404-407 when d_w is given), backward, torch.nn.utils.clip_grad_norm_
 * (base.py:439-443) and the optimiser update (base.py:104-135: adamw / adam / sgd).
 *
 * A trainer is built from a packed description of the architecture and of the flat
 * parameter layout (nessai_b200/train_plan.py; struct TrPlan in csrc/train.cuh):
 * h_plan int32[n_plan_ints], h_itab int32[n_itab] (permutations, mask index lists),
 * h_reduce_idx int32[n_reduce] (parameters whose gradient is a plain sum over rows). */
typedef struct nb200_trainer nb200_trainer;
int nb200_trainer_create(nb200_trainer** out, const int32_t* h_plan, int n_plan_ints,
                         const int32_t* h_itab, int n_itab, const int32_t* h_reduce_idx,
                         int n_reduce);
int nb200_trainer_destroy(nb200_trainer* trainer);
/* Replace the index tables (permutations, mask index lists) of a trainer: a NEW flow of the same
 * architecture (flows/utils.py:249-292 reset_permutations; every level of the importance sampler,
 * flowmodel/importance.py:80-99) reuses the trainer's plan and workspaces. */
int nb200_trainer_set_itab(nb200_trainer* trainer, const int32_t* h_itab, int n_itab, void* stream);
/* MADE masks (flows/maf.py -> nflows MaskedLinear): h_mask float[n_params], the mask value for
 * masked-linear weights and 1 elsewhere (NULL: none).  Masked weights are held at exactly zero
 * (theta_p *= mask at the start of every epoch) and their gradients are masked. */
int nb200_trainer_set_param_mask(nb200_trainer* trainer, const float* h_mask);
/* Copy the gradient of the last step (n_params floats, after clipping) to d_out. */
int nb200_trainer_copy_grad(nb200_trainer* trainer, float* d_out, void* stream);

/* d_theta_p: trainable parameters (flat fp32, reference state_dict order), d_theta_b: float
 * buffers (BatchNorm running statistics), d_m / d_v: Adam moments (same size as d_theta_p).
 * Rows of batch k are d_x[d_perm[k*batch_size + i]] (d_perm NULL: in order), d_x row-major
 * [n_rows][D].  opt_kind: 0 AdamW, 1 Adam (L2 decay), 2 SGD, -1 gradient only (no update).
 * clip <= 0 disables clipping.  step0 = optimiser steps taken before this call (bias
 * correction).  d_loss_sum[0] += loss of every batch (caller zeroes it); d_step_info
 * (may be NULL): float[2 * n_batches] = {loss, gradient norm} per batch. */
int nb200_train_epoch(nb200_trainer* trainer, float* d_theta_p, float* d_theta_b, float* d_m,
                      float* d_v, const float* d_x, const float* d_w, const int64_t* d_perm,
                      int64_t n_rows, int batch_size, int opt_kind, double lr, double beta1,
                      double beta2, double eps, double weight_decay, double clip, int64_t step0,
                      float* d_loss_sum, float* d_step_info, void* stream);

/* flowmodel/base.py:620-680, the epoch loop of FlowModel.train, for up to 64 epochs in ONE
 * cooperative launch: every optimisation step of every epoch (as nb200_train_epoch; epoch e takes
 * its rows in the order d_perm[e * n_rows ...] and the learning rate h_lr[e]), the validation
 * loss of the epoch on d_xv (base.py:454-523; n_val == 0: NaN) and the reference's loop control
 * on the device: when `validate`, an epoch whose validation loss beats the best so far snapshots
 * (theta_p, theta_b) into (d_best_p, d_best_b) (base.py:652-656), and the run stops once
 * epoch - best_epoch > patience (base.py:658-660).  d_hist: float[2 * n_epochs] = {sum of the
 * batch losses, validation loss} per epoch.  d_ctl: 4 words {float best_val, int best_epoch,
 * int epochs_done, int stop}, read at the start (the caller initialises {+inf, 0, 0, 0}) and
 * written at the end; epoch0 = epochs finished before this call. */
int nb200_train_run(nb200_trainer* trainer, float* d_theta_p, float* d_theta_b, int64_t n_theta_b,
                    float* d_m, float* d_v, const float* d_x, const float* d_w,
                    const int64_t* d_perm, int64_t n_rows, int batch_size, const float* d_xv,
                    const float* d_wv, int64_t n_val, int n_epochs, int epoch0, int validate,
                    int patience, int opt_kind, const double* h_lr, double beta1, double beta2,
                    double eps, double weight_decay, double clip, int64_t step0, float* d_hist,
                    void* d_ctl, float* d_best_p, float* d_best_b, void* stream);

/* flowmodel/base.py:454-523 FlowModel._validate: eval-mode (running statistics) loss of
 * the UNFOLDED parameters, d_loss[0] = -sum(w log_prob)/sum(w); d_logp (may be NULL):
 * per-row log_prob. */
int nb200_eval_loss(nb200_trainer* trainer, float* d_theta_p, float* d_theta_b, const float* d_x,
                    const float* d_w, int64_t n, float* d_loss, float* d_logp, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NESSAI_B200_H */
