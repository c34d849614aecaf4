// TEST HARNESS ONLY: the fused populate turn of the generic kernel -- Philox draw (philox.cuh),
// latent-radius truncation, inverse flow (flow_interp.cuh), float64 rescale / bounds / weights /
// statistics (populate_common.cuh): the CUDA sources, unchanged -- under the CPU SIMT shim.  The
// kernel body mirrors populate_draw_kernel of nessai_b200.cu with the dynamic shared memory
// replaced by a static block.  Compile with -I tests/_hostcheck/fake_cuda.
#define nb200 nb200_simt_populate
#define SIMT_HAVE_POPULATE_COMMON 1
#include "simt_shim.h"

inline float __int_as_float(int v) {
  float f;
  std::memcpy(&f, &v, sizeof f);
  return f;
}

#include "../../nessai_b200/csrc/flow_interp.cuh"
#include "../../nessai_b200/csrc/philox.cuh"
#include "../../nessai_b200/csrc/populate_common.cuh"

namespace {
constexpr float LOG_2PI = 1.8378770664093453f;
alignas(16) float g_smem[64 * 1024];

template <int ACT>
void draw_kernel(nb200::FlowProgramDev P, nb200::PopulateArgs A) {
  using namespace nb200;
  float* Ws;
  float* bufs[4];
  const int BS = blockDim.x;
  carve_buffers(g_smem, P, BS, Ws, bufs);
  const int64_t ntiles = (A.n + BS - 1) / BS;
  const int D = P.D;
  double vmax = -INFINITY, vcount = 0.0;
  __shared__ double s_log_const;
  if (threadIdx.x == 0) s_log_const = populate_log_const(A, D);
  __syncthreads();
  const double log_const = s_log_const;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row = tile * BS + threadIdx.x;
    float ss = 0.f;
    for (int d0 = 0; d0 < D; d0 += 4) {
      const Philox4 r = philox4x32_10(A.seed, A.row_offset + row, d0 / 4, 0);
      float v[4];
      box_muller(r.x, r.y, v[0], v[1]);
      box_muller(r.z, r.w, v[2], v[3]);
      for (int j = 0; j < 4 && d0 + j < D; ++j) {
        ss = fmaf(v[j], v[j], ss);
        const float zz = v[j] * A.sqrt_t;
        bufs[BUF_X0][(d0 + j) * BS] = zz;
        if (A.z && row < A.n) A.z[row * D + d0 + j] = zz;
      }
    }
    const float rad = sqrtf(ss) * A.sqrt_t;
    const bool alive = !(A.r_max > 0.f) || (rad <= A.r_max);
    const float logj = run_program<ACT>(P, Ws, bufs, BS) + P.const_logdet;
    const float base_lp = -0.5f * ss - 0.5f * D * LOG_2PI;
    const float* fin = bufs[P.final_buf];
    populate_row(A, D, [&](int d) { return fin[d * BS]; }, row, alive, base_lp, logj, vmax, vcount, A.scale, A.shift,
                 A.lo, A.hi, log_const);
  }
  populate_publish(A, vmax, vcount);
}
}  // namespace

extern "C" int simt_populate_draw(int grid, const int32_t* ops, int n_ops, const float* blob, int D, int H,
                                  int activation, int final_buf, double const_logdet, int64_t n, uint64_t seed,
                                  uint64_t row_offset, float r_max, float sqrt_t, const double* scale,
                                  const double* shift, const double* lo, const double* hi, double log_prior_const,
                                  double min_log_q, float* xp, double* logq, double* logw, float* z, double* stats) {
  using namespace nb200;
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
  PopulateArgs A;
  A.n = n, A.seed = seed, A.row_offset = row_offset, A.r_max = r_max, A.sqrt_t = sqrt_t > 0.f ? sqrt_t : 1.f;
  A.scale = scale, A.shift = shift, A.lo = lo, A.hi = hi, A.log_prior_const = log_prior_const;
  A.min_log_q = std::isnan(min_log_q) ? -INFINITY : min_log_q;
  A.xp = xp, A.logq = logq, A.logw = logw, A.z = z, A.stats = stats;
  int BS = 0;
  for (int bs : {128, 64, 32})
    if (!BS && interp_smem_bytes(P, bs) <= 226 * 1024) BS = bs;
  if (!BS || interp_smem_bytes(P, BS) > sizeof(g_smem)) return 5;
  if (activation == ACT_RELU) simt_launch(draw_kernel<ACT_RELU>, (unsigned)grid, (unsigned)BS, P, A);
  else if (activation == ACT_TANH) simt_launch(draw_kernel<ACT_TANH>, (unsigned)grid, (unsigned)BS, P, A);
  else simt_launch(draw_kernel<ACT_SILU>, (unsigned)grid, (unsigned)BS, P, A);
  return 0;
}

extern "C" double nb200_host_erfcinv(double) { return 0.0; }  // declared by the shim; unused here
