// TEST HARNESS ONLY: the generic fp32 flow interpreter (nessai_b200/csrc/flow_interp.cuh: weight
// staging, linear / affine-coupling / spline-coupling ops, MADE flags -- the CUDA source, unchanged)
// under the CPU SIMT shim.  The thin kernel around it mirrors flow_apply_kernel of
// nessai_b200.cu (row in -> column buffers -> run_program -> row out, log|J|, log-prob), with the
// dynamic shared memory replaced by a static block.  Compile with -I tests/_hostcheck/fake_cuda.
#define nb200 nb200_simt_interp
#include "simt_shim.h"

inline float __int_as_float(int v) {
  float f;
  std::memcpy(&f, &v, sizeof f);
  return f;
}

#include "../../nessai_b200/csrc/flow_interp.cuh"

namespace {
constexpr float LOG_2PI = 1.8378770664093453f;
alignas(16) float g_smem[64 * 1024];  // 256 KB >= the 226 KB the launcher allows

template <int ACT>
void apply_kernel(nb200::FlowProgramDev P, const float* in, float* out, float* out_logj, float* out_lp, int64_t n,
                  int lp_mode, float base_inv_var, float base_log_z) {
  using namespace nb200;
  float* Ws;
  float* bufs[4];
  const int BS = blockDim.x;
  carve_buffers(g_smem, P, BS, Ws, bufs);
  const int64_t ntiles = (n + BS - 1) / BS;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row = tile * BS + threadIdx.x;
    const bool valid = row < n;
    float ss_in = 0.f;
    for (int d = 0; d < P.D; ++d) {
      const float v = valid ? in[row * P.D + d] : 0.f;
      bufs[BUF_X0][d * BS] = v;
      ss_in = fmaf(v, v, ss_in);
    }
    const float ld = run_program<ACT>(P, Ws, bufs, BS) + P.const_logdet;
    const float* fin = bufs[P.final_buf];
    float ss_out = 0.f;
    for (int d = 0; d < P.D; ++d) {
      const float o = fin[d * BS];
      ss_out = fmaf(o, o, ss_out);
      if (valid && out) out[row * P.D + d] = o;
    }
    if (valid) {
      if (out_logj) out_logj[row] = ld;
      if (out_lp) {
        const float c = base_log_z, hv = 0.5f * base_inv_var;
        out_lp[row] = (lp_mode == 1) ? (-hv * ss_in - c) - ld : (-hv * ss_out - c) + ld;
      }
    }
  }
}
}  // namespace

// ops: int32[n_ops][16] and blob: float32[n_blob] exactly as nb200_flow_set_program receives them.
extern "C" int simt_flow_apply(int grid, const int32_t* ops, int n_ops, const float* blob, int D, int H,
                               int activation, int final_buf, double const_logdet, const float* in, float* out,
                               float* logj, float* lp, int64_t n, int lp_mode, double base_var) {
  using namespace nb200;
  // as launch_apply of nessai_b200.cu forms the base-distribution constants
  const float base_inv_var = (float)(1.0 / base_var);
  const float base_log_z = (float)(0.5 * D * std::log(2.0 * M_PI * base_var));
  (void)LOG_2PI;
  FlowProgramDev P;
  P.ops = reinterpret_cast<const FlowOp*>(ops);
  P.blob = blob;
  P.n_ops = n_ops;
  P.D = D;
  P.Dpad = (D + 7) / 8 * 8;
  P.Hpad = (H + 7) / 8 * 8;
  P.activation = activation;
  P.final_buf = final_buf;
  P.wmax = 0;
  for (int i = 0; i < n_ops; ++i) P.wmax = std::max(P.wmax, P.ops[i].K * P.ops[i].Npad + P.ops[i].Npad);
  P.const_logdet = (float)const_logdet;
  int BS = 0;
  for (int bs : {128, 64, 32})
    if (!BS && interp_smem_bytes(P, bs) <= 226 * 1024) BS = bs;
  if (!BS || interp_smem_bytes(P, BS) > sizeof(g_smem)) return 5;
  if (activation == ACT_RELU) simt_launch(apply_kernel<ACT_RELU>, (unsigned)grid, (unsigned)BS, P, in, out, logj, lp, n, lp_mode, base_inv_var, base_log_z);
  else if (activation == ACT_TANH) simt_launch(apply_kernel<ACT_TANH>, (unsigned)grid, (unsigned)BS, P, in, out, logj, lp, n, lp_mode, base_inv_var, base_log_z);
  else simt_launch(apply_kernel<ACT_SILU>, (unsigned)grid, (unsigned)BS, P, in, out, logj, lp, n, lp_mode, base_inv_var, base_log_z);
  return 0;
}

extern "C" double nb200_host_erfcinv(double) { return 0.0; }  // declared by the shim; unused here
