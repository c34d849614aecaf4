// C ABI + generic fp32 kernels of nessai_b200 (see include/nessai_b200.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <new>
#include <cstdlib>
#include <vector>

#include "../../include/nessai_b200.h"
#include "flow_interp.cuh"
#include "philox.cuh"
#include "populate_common.cuh"
#include "reparam_tail.cuh"
#include "flow_tc.cuh"
#include "flow_tc_res.cuh"
#include "flow_tc_nsf.cuh"
#include "coupling.cuh"

using namespace nb200;

// ----------------------------------------------------------------------------- errors
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CUDA_OK(expr)                                                                  \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess)                                                             \
      return fail(2, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                  __LINE__);                                                           \
  } while (0)

extern "C" int nb200_version(void) { return 100; }
extern "C" const char* nb200_last_error(void) { return g_err; }
extern "C" int64_t nb200_launch_count(void) { return g_launches.load(); }
extern "C" void nb200_reset_launch_count(void) { g_launches.store(0); }
extern "C" int nb200_set_tensor_core_path(int enabled) {
  const int prev = tc_enabled() ? 1 : 0;
  tc_enabled_flag() = enabled ? 1 : 0;
  return prev;
}

// ----------------------------------------------------------------------------- flow object
struct DirProgram {
  FlowOp* d_ops = nullptr;
  float* d_blob = nullptr;
  FlowOp* h_ops = nullptr;
  int n_ops = 0;
  int64_t n_blob = 0;
  int final_buf = 0;
  int wmax = 0;
  double const_logdet = 0.0;
  TcProgram tc;  // tensor-core specialisation (valid == false when not applicable)
  RsProgram rs;  // tensor-core specialisation for the ResidualNet conditioner
  NsProgram ns;  // tensor-core specialisation for the neural spline flow
};

struct nb200_flow {
  int D, H, activation;
  int device;
  int num_sms;
  double base_var = 1.0;  // base distribution N(0, base_var I)
  DirProgram dir[2];
};

static void free_dir(DirProgram& p) {
  if (p.d_ops) cudaFree(p.d_ops);
  if (p.d_blob) cudaFree(p.d_blob);
  delete[] p.h_ops;
  tc_free(p.tc);
  rs_free(p.rs);
  ns_free(p.ns);
  p = DirProgram();
}

// Compute capability and SM count of a device.  (cudaGetDeviceProperties fills ~100 fields, some of
// them through slow driver queries: tens of milliseconds per call -- paid by every flow and trainer
// created; the three attributes below are microseconds.)
struct DeviceInfo {
  int major = 0, minor = 0, sms = 0;
};
static int device_info(int dev, DeviceInfo& d) {
  CUDA_OK(cudaDeviceGetAttribute(&d.major, cudaDevAttrComputeCapabilityMajor, dev));
  CUDA_OK(cudaDeviceGetAttribute(&d.minor, cudaDevAttrComputeCapabilityMinor, dev));
  CUDA_OK(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
  return 0;
}

extern "C" int nb200_flow_create(nb200_flow** out, int D, int H, int activation) {
  if (!out) return fail(1, "nb200_flow_create: out is NULL");
  if (D < 1 || D > 1024 || H < 1 || H > 4096) return fail(1, "nb200_flow_create: bad sizes D=%d H=%d", D, H);
  if (activation < 0 || activation > 2) return fail(1, "nb200_flow_create: bad activation %d", activation);
  int dev = 0;
  CUDA_OK(cudaGetDevice(&dev));
  DeviceInfo prop;
  if (int rc = device_info(dev, prop)) return rc;
  if (prop.major != 10)
    return fail(3, "nessai_b200 kernels are built for sm_100a only; device %d is sm_%d%d", dev,
                prop.major, prop.minor);
  nb200_flow* f = new (std::nothrow) nb200_flow();
  if (!f) return fail(4, "out of host memory");
  f->D = D;
  f->H = H;
  f->activation = activation;
  f->device = dev;
  f->num_sms = prop.sms;
  *out = f;
  return 0;
}

extern "C" int nb200_flow_destroy(nb200_flow* f) {
  if (!f) return 0;
  free_dir(f->dir[0]);
  free_dir(f->dir[1]);
  delete f;
  return 0;
}

extern "C" int nb200_flow_set_program(nb200_flow* f, int direction, const int32_t* h_ops,
                                      int n_ops, const float* h_blob, int64_t n_blob,
                                      int final_buf, double const_logdet) {
  if (!f || direction < 0 || direction > 1 || !h_ops || !h_blob || n_ops < 1)
    return fail(1, "nb200_flow_set_program: bad arguments");
  DirProgram& p = f->dir[direction];
  free_dir(p);
  p.h_ops = new FlowOp[n_ops];
  memcpy(p.h_ops, h_ops, sizeof(FlowOp) * n_ops);
  int wmax = 0;
  for (int i = 0; i < n_ops; ++i) {
    const FlowOp& op = p.h_ops[i];
    if (op.Npad % 8 || op.K < 1 || op.w_off % 4 || op.b_off % 4 ||
        (int64_t)op.w_off + (int64_t)op.K * op.Npad > n_blob || op.b_off + op.Npad > n_blob)
      return fail(1, "nb200_flow_set_program: malformed op %d", i);
    if (op.type == OP_COUPLING_SPLINE && (op.e0 < 2 || op.e0 > 16))
      return fail(1, "nb200_flow_set_program: spline bins must be in [2, 16], got %d", op.e0);
    const int w = op.K * op.Npad + op.Npad;
    if (w > wmax) wmax = w;
  }
  CUDA_OK(cudaMalloc(&p.d_ops, sizeof(FlowOp) * n_ops));
  CUDA_OK(cudaMalloc(&p.d_blob, sizeof(float) * n_blob));
  CUDA_OK(cudaMemcpy(p.d_ops, h_ops, sizeof(FlowOp) * n_ops, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(p.d_blob, h_blob, sizeof(float) * n_blob, cudaMemcpyHostToDevice));
  p.n_ops = n_ops;
  p.n_blob = n_blob;
  p.final_buf = final_buf;
  p.wmax = wmax;
  p.const_logdet = const_logdet;
  // try the tcgen05 specialisation (RealNVP + MLP conditioner shapes it covers)
  if (int rc = tc_build(p.tc, p.h_ops, n_ops, h_blob, f->D, f->H, f->activation, final_buf))
    return fail(rc, "tc_build failed: %s", cudaGetErrorString(cudaGetLastError()));
  p.tc.const_logdet = (float)const_logdet;
  if (!p.tc.valid) {
    if (int rc = rs_build(p.rs, p.h_ops, n_ops, h_blob, f->D, f->H, f->activation))
      return fail(rc, "rs_build failed: %s", cudaGetErrorString(cudaGetLastError()));
    p.rs.const_logdet = (float)const_logdet;
    if (!p.rs.valid) {
      if (int rc = ns_build(p.ns, p.h_ops, n_ops, h_blob, f->D, f->H, f->activation))
        return fail(rc, "ns_build failed: %s", cudaGetErrorString(cudaGetLastError()));
      if (!p.ns.valid)  // RealNVP with the ResidualNet conditioner at 17 .. 32 features
        if (int rc = ac_build(p.ns, p.h_ops, n_ops, h_blob, f->D, f->H, f->activation))
          return fail(rc, "ac_build failed: %s", cudaGetErrorString(cudaGetLastError()));
      p.ns.const_logdet = (float)const_logdet;
    }
  }
  return 0;
}

static FlowProgramDev make_dev(const nb200_flow* f, const DirProgram& p) {
  FlowProgramDev P;
  P.ops = p.d_ops;
  P.blob = p.d_blob;
  P.n_ops = p.n_ops;
  P.D = f->D;
  P.Dpad = (f->D + 7) / 8 * 8;
  P.Hpad = (f->H + 7) / 8 * 8;
  P.activation = f->activation;
  P.final_buf = p.final_buf;
  P.wmax = p.wmax;
  P.const_logdet = (float)p.const_logdet;
  return P;
}

// ----------------------------------------------------------------------------- kernels
#define LOG_2PI 1.8378770664093453f

template <int ACT>
__global__ void __launch_bounds__(128)
flow_apply_kernel(FlowProgramDev P, const float* __restrict__ in, float* __restrict__ out,
                  float* __restrict__ out_logj, float* __restrict__ out_lp, int64_t n,
                  int lp_mode, float base_inv_var, float base_log_z) {
  extern __shared__ float4 smem4[];
  float* Ws;
  float* bufs[4];
  const int BS = blockDim.x;
  carve_buffers(reinterpret_cast<float*>(smem4), P, BS, Ws, bufs);
  const int64_t ntiles = (n + BS - 1) / BS;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row = tile * BS + threadIdx.x;
    const bool valid = row < n;
    float ss_in = 0.f;
    for (int d = 0; d < P.D; ++d) {
      const float v = valid ? __ldg(in + row * P.D + d) : 0.f;
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
        // log N(z; 0, var I) = -0.5 |z|^2 / var - 0.5 D log(2 pi var)  (flows/distributions.py:45-56)
        const float c = base_log_z, hv = 0.5f * base_inv_var;
        out_lp[row] = (lp_mode == 1) ? (-hv * ss_in - c) - ld : (-hv * ss_out - c) + ld;
      }
    }
  }
}

__global__ void sample_latent_kernel(float* __restrict__ z, int64_t n, int D, uint64_t seed,
                                     uint64_t row_offset, float std) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  for (int d0 = 0; d0 < D; d0 += 4) {
    const Philox4 r = philox4x32_10(seed, row_offset + row, d0 / 4, 0);
    float v[4];
    box_muller(r.x, r.y, v[0], v[1]);
    box_muller(r.z, r.w, v[2], v[3]);
    for (int j = 0; j < 4 && d0 + j < D; ++j) z[row * D + d0 + j] = v[j] * std;
  }
}

template <int ACT>
__global__ void __launch_bounds__(128)
populate_draw_kernel(FlowProgramDev P, PopulateArgs A) {
  extern __shared__ float4 smem4[];
  float* Ws;
  float* bufs[4];
  const int BS = blockDim.x;
  carve_buffers(reinterpret_cast<float*>(smem4), P, BS, Ws, bufs);
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
    // latent-radius truncation on |z| (after temperature scaling)
    const float rad = sqrtf(ss) * A.sqrt_t;
    const bool alive = !(A.r_max > 0.f) || (rad <= A.r_max);
    const float logj = run_program<ACT>(P, Ws, bufs, BS) + P.const_logdet;
    const float base_lp = -0.5f * ss - 0.5f * D * LOG_2PI;
    const float* fin = bufs[P.final_buf];
    populate_row(A, D, [&](int d) { return fin[d * BS]; }, row, alive, base_lp, logj, vmax, vcount,
                 A.scale, A.shift, A.lo, A.hi, log_const);
  }
  populate_publish(A, vmax, vcount);
}

// ----------------------------------------------------------------------------- accept + compact
#include "accept.cuh"

// ----------------------------------------------------------------------------- launch helpers
template <typename K>
static int prep_kernel(K kernel, size_t smem) {
  // opt in to > 48 KB dynamic shared memory (static shared memory counts
  // against the 227 KB limit, so ask for what the launch needs, rounded up)
  // (the attribute is per device: the cache is keyed on kernel AND current device)
  static thread_local const void* done[64];
  static thread_local int done_dev[64];
  static thread_local size_t done_smem[64];
  static thread_local int ndone = 0;
  int dev = 0;
  CUDA_OK(cudaGetDevice(&dev));
  int slot = -1;
  for (int i = 0; i < ndone; ++i)
    if (done[i] == (const void*)kernel && done_dev[i] == dev) slot = i;
  if (slot >= 0 && done_smem[slot] >= smem) return 0;
  CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (slot < 0 && ndone < 64) slot = ndone++;
  if (slot >= 0) {
    done[slot] = (const void*)kernel;
    done_dev[slot] = dev;
    done_smem[slot] = smem;
  }
  return 0;
}

static int pick_block(const FlowProgramDev& P, int& BS, size_t& smem) {
  for (int bs : {128, 64, 32}) {
    const size_t s = interp_smem_bytes(P, bs);
    if (s <= 226 * 1024) {
      BS = bs;
      smem = s;
      return 0;
    }
  }
  return fail(5, "flow too large for the generic kernel (D=%d H=%d needs > 227 KB shared memory)",
              P.D, P.Hpad);
}

static int launch_apply(nb200_flow* f, int direction, const float* in, float* out, float* logj,
                        float* lp, int64_t n, cudaStream_t st) {
  DirProgram& p = f->dir[direction];
  if (!p.d_ops) return fail(6, "program for direction %d not set", direction);
  if (n <= 0) return 0;
  TcIO io{in, out, logj, lp, direction == 1 ? 1 : 2, n};
  io.base_inv_var = (float)(1.0 / f->base_var);
  io.base_log_z = (float)(0.5 * f->D * std::log(2.0 * M_PI * f->base_var));
  if (p.tc.valid && tc_enabled()) {
    g_launches += 1;
    return tc_launch_apply(p.tc, io, f->num_sms, st)
               ? fail(2, "tc kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()))
               : 0;
  }
  if (p.rs.valid && tc_enabled()) {
    const int nl = rs_launch<0>(p.rs, io, PopulateArgs(), n, f->num_sms, st);
    if (!nl) return fail(2, "tc resnet kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    g_launches += nl;
    return 0;
  }
  if (p.ns.valid && tc_enabled()) {
    const int nl = ns_launch<0>(p.ns, io, PopulateArgs(), n, f->num_sms, st);
    if (!nl) return fail(2, "tc nsf kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    g_launches += nl;
    return 0;
  }
  FlowProgramDev P = make_dev(f, p);
  int BS;
  size_t smem;
  if (int rc = pick_block(P, BS, smem)) return rc;
  const int64_t ntiles = (n + BS - 1) / BS;
  const int grid = (int)std::min<int64_t>(ntiles, (int64_t)f->num_sms * 16);
  const int lp_mode = direction == 1 ? 1 : 2;
#define LAUNCH_APPLY(ACT)                                                               \
  {                                                                                     \
    if (int rc = prep_kernel(flow_apply_kernel<ACT>, smem)) return rc;                  \
    flow_apply_kernel<ACT><<<grid, BS, smem, st>>>(P, in, out, logj, lp, n, lp_mode,    \
                                                   io.base_inv_var, io.base_log_z);      \
  }
  if (f->activation == ACT_RELU) LAUNCH_APPLY(ACT_RELU)
  else if (f->activation == ACT_TANH) LAUNCH_APPLY(ACT_TANH)
  else LAUNCH_APPLY(ACT_SILU)
#undef LAUNCH_APPLY
  g_launches += 1;
  CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int nb200_flow_inverse(nb200_flow* f, const float* d_z, float* d_x, float* d_logj,
                                  float* d_logq, int64_t n, void* stream) {
  if (n <= 0) return 0;
  if (!f || !d_z) return fail(1, "nb200_flow_inverse: bad arguments");
  return launch_apply(f, 1, d_z, d_x, d_logj, d_logq, n, (cudaStream_t)stream);
}

extern "C" int nb200_flow_forward(nb200_flow* f, const float* d_x, float* d_z, float* d_logj,
                                  float* d_logp, int64_t n, void* stream) {
  if (n <= 0) return 0;
  if (!f || !d_x) return fail(1, "nb200_flow_forward: bad arguments");
  return launch_apply(f, 0, d_x, d_z, d_logj, d_logp, n, (cudaStream_t)stream);
}

extern "C" int nb200_flow_set_base_variance(nb200_flow* f, double var) {
  if (!f || !(var > 0.0) || !std::isfinite(var)) return fail(1, "nb200_flow_set_base_variance: bad arguments");
  f->base_var = var;
  return 0;
}

extern "C" int nb200_sample_latent(float* d_z, int64_t n, int D, uint64_t seed,
                                   uint64_t row_offset, double std_dev, void* stream) {
  if (n <= 0) return 0;
  if (!d_z || D < 1 || !(std_dev > 0.0)) return fail(1, "nb200_sample_latent: bad arguments");
  const int bs = 256;
  sample_latent_kernel<<<(unsigned)((n + bs - 1) / bs), bs, 0, (cudaStream_t)stream>>>(
      d_z, n, D, seed, row_offset, (float)std_dev);
  g_launches += 1;
  CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int nb200_populate_draw(nb200_flow* f, int64_t n, uint64_t seed, uint64_t row_offset,
                                   float r_max, float sqrt_temperature, const double* d_scale,
                                   const double* d_shift, const double* d_lo, const double* d_hi,
                                   double log_prior_const, double min_log_q, float* d_xp,
                                   double* d_logq, double* d_logw, float* d_z, double* d_stats,
                                   void* stream) {
  if (!f || !d_scale || !d_shift || !d_lo || !d_hi || !d_xp || !d_logq || !d_logw || !d_stats)
    return fail(1, "nb200_populate_draw: bad arguments");
  DirProgram& p = f->dir[1];
  if (!p.d_ops) return fail(6, "inverse program not set");
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  PopulateArgs A;
  A.n = n;
  A.seed = seed;
  A.row_offset = row_offset;
  A.r_max = r_max;
  // z = sqrt(T var) v, v ~ N(0, I): a N(0, var I) base distribution (flows/distributions.py:17-73)
  // enters the turn exactly like a latent temperature (populate_common.cuh: populate_log_const)
  A.sqrt_t = (sqrt_temperature > 0.f ? sqrt_temperature : 1.f) * (float)std::sqrt(f->base_var);
  A.scale = d_scale;
  A.shift = d_shift;
  A.lo = d_lo;
  A.hi = d_hi;
  A.log_prior_const = log_prior_const;
  A.min_log_q = isnan(min_log_q) ? -INFINITY : min_log_q;
  A.xp = d_xp;
  A.logq = d_logq;
  A.logw = d_logw;
  A.z = d_z;
  A.stats = d_stats;
  if (p.tc.valid && tc_enabled()) {
    g_launches += 1;
    return tc_launch_populate(p.tc, A, f->num_sms, st)
               ? fail(2, "tc populate launch failed: %s", cudaGetErrorString(cudaGetLastError()))
               : 0;
  }
  if (p.rs.valid && tc_enabled()) {
    const int nl = rs_launch<1>(p.rs, TcIO(), A, n, f->num_sms, st);
    if (!nl) return fail(2, "tc resnet populate launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    g_launches += nl;
    return 0;
  }
  if (p.ns.valid && tc_enabled()) {
    const int nl = ns_launch<1>(p.ns, TcIO(), A, n, f->num_sms, st);
    if (!nl) return fail(2, "tc nsf populate launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    g_launches += nl;
    return 0;
  }
  FlowProgramDev P = make_dev(f, p);
  int BS;
  size_t smem;
  if (int rc = pick_block(P, BS, smem)) return rc;
  const int64_t ntiles = (n + BS - 1) / BS;
  const int grid = (int)std::min<int64_t>(ntiles, (int64_t)f->num_sms * 16);
#define LAUNCH_POP(ACT)                                                    \
  {                                                                        \
    if (int rc = prep_kernel(populate_draw_kernel<ACT>, smem)) return rc;  \
    populate_draw_kernel<ACT><<<grid, BS, smem, st>>>(P, A);               \
  }
  if (f->activation == ACT_RELU) LAUNCH_POP(ACT_RELU)
  else if (f->activation == ACT_TANH) LAUNCH_POP(ACT_TANH)
  else LAUNCH_POP(ACT_SILU)
#undef LAUNCH_POP
  g_launches += 1;
  CUDA_OK(cudaGetLastError());
  return 0;
}

static int populate_accept_impl(int64_t n, int D, const float* d_xp, const double* d_x64,
                                const double* d_scale, const double* d_shift, const double* d_logw,
                                const double* d_logl, const double* d_max, uint64_t seed,
                                uint64_t row_offset, double log_p_value,
                                const uint8_t* d_row_template, int row_bytes,
                                const int32_t* h_field_offsets, int logl_offset, uint8_t* d_rows,
                                int64_t capacity, int64_t write_offset, int64_t* d_counts,
                                int64_t* d_scratch, void* stream) {
  if ((!d_xp && !d_x64) || !d_scale || !d_shift || !d_logw || !d_max || !d_row_template || !h_field_offsets || !d_rows || !d_counts ||
      !d_scratch)
    return fail(1, "nb200_populate_accept: bad arguments");
  if (row_bytes % 4 || row_bytes <= 0) return fail(1, "row_bytes must be a positive multiple of 4");
  if (D < 1 || D > 256) return fail(1, "nb200_populate_accept: D out of range");
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0) {
    // an empty shard accepts nothing: {accepted, written} = {0, 0}, not the previous turn's
    CUDA_OK(cudaMemsetAsync(d_counts, 0, 2 * sizeof(int64_t), st));
    return 0;
  }
  RowFormat F;
  F.row_words = row_bytes / 4;
  F.D = D;
  for (int d = 0; d < D; ++d) {
    if (h_field_offsets[d] % 4 || h_field_offsets[d] < 0 || h_field_offsets[d] + 8 > row_bytes)
      return fail(1, "bad field offset %d", h_field_offsets[d]);
    F.off[d] = h_field_offsets[d];
  }
  F.logp_off = h_field_offsets[D];
  if (F.logp_off >= 0 && (F.logp_off % 4 || F.logp_off + 8 > row_bytes))
    return fail(1, "bad logP offset");
  if (d_logl && logl_offset >= 0 && (logl_offset % 4 || logl_offset + 8 > row_bytes))
    return fail(1, "bad logL offset");
  const int64_t nchunks = (n + ACC_CHUNK - 1) / ACC_CHUNK;
  const size_t smem = (size_t)D * 16 + (size_t)F.row_words * 6 + 16;
  if (smem > 40 * 1024) return fail(1, "nb200_populate_accept: record too large (%d bytes)", row_bytes);
  CUDA_OK(cudaMemsetAsync(d_scratch, 0, (size_t)(nchunks + 1) * sizeof(int64_t), st));
  accept_fused_kernel<<<(unsigned)nchunks, ACC_THREADS, smem, st>>>(
      d_xp, d_x64, d_scale, d_shift, d_logw, d_logl, d_logl ? logl_offset : -1, d_max, n, seed, row_offset,
      reinterpret_cast<unsigned long long*>(d_scratch), nchunks, log_p_value,
      reinterpret_cast<const uint32_t*>(d_row_template), F, reinterpret_cast<uint32_t*>(d_rows), capacity,
      write_offset, d_counts);
  g_launches += 1;
  CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int nb200_populate_accept(int64_t n, int D, const float* d_xp, const double* d_scale,
                                     const double* d_shift, const double* d_logw,
                                     const double* d_logl, const double* d_max, uint64_t seed,
                                     uint64_t row_offset, double log_p_value,
                                     const uint8_t* d_row_template, int row_bytes,
                                     const int32_t* h_field_offsets, int logl_offset, uint8_t* d_rows,
                                     int64_t capacity, int64_t write_offset, int64_t* d_counts,
                                     int64_t* d_scratch, void* stream) {
  if (!d_xp) return fail(1, "nb200_populate_accept: bad arguments");
  return populate_accept_impl(n, D, d_xp, nullptr, d_scale, d_shift, d_logw, d_logl, d_max, seed,
                              row_offset, log_p_value, d_row_template, row_bytes, h_field_offsets,
                              logl_offset, d_rows, capacity, write_offset, d_counts, d_scratch, stream);
}

extern "C" int nb200_populate_accept_x64(int64_t n, int D, const double* d_x64, const double* d_logw,
                                         const double* d_logl, const double* d_max, uint64_t seed,
                                         uint64_t row_offset, double log_p_value,
                                         const uint8_t* d_row_template, int row_bytes,
                                         const int32_t* h_field_offsets, int logl_offset,
                                         uint8_t* d_rows, int64_t capacity, int64_t write_offset,
                                         int64_t* d_counts, int64_t* d_scratch, void* stream) {
  if (!d_x64) return fail(1, "nb200_populate_accept_x64: bad arguments");
  // scale / shift are unused when x64 is given; any valid device pointer of >= D doubles serves
  return populate_accept_impl(n, D, nullptr, d_x64, d_x64, d_x64, d_logw, d_logl, d_max, seed,
                              row_offset, log_p_value, d_row_template, row_bytes, h_field_offsets,
                              logl_offset, d_rows, capacity, write_offset, d_counts, d_scratch, stream);
}

extern "C" int nb200_reparam_tail(int64_t n, int D, const float* d_xp, const int32_t* d_kind,
                                  const int32_t* d_src, const double* d_pre_scale,
                                  const double* d_pre_shift,
                                  const double* d_scale, const double* d_shift, const double* d_lo,
                                  const double* d_hi, double log_prior_const, double min_log_q,
                                  double* d_logq, double* d_logw, double* d_x64, double* d_stats,
                                  void* stream) {
  if (!d_xp || !d_kind || !d_scale || !d_shift || !d_lo || !d_hi || !d_logq || !d_logw || !d_x64 || !d_stats)
    return fail(1, "nb200_reparam_tail: bad arguments");
  if (D < 1 || D > TAIL_MAXD) return fail(1, "nb200_reparam_tail: D=%d out of range (1..%d)", D, TAIL_MAXD);
  if (!d_pre_scale != !d_pre_shift) return fail(1, "nb200_reparam_tail: pre_scale and pre_shift go together");
  if (n <= 0) return 0;
  int dev = 0, sms = 148;
  CUDA_OK(cudaGetDevice(&dev));
  CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t smem = tail_smem_bytes(D);
  if (int rc = prep_kernel(reparam_tail_kernel, smem)) return rc;
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / (smem + 4096)));
  const int grid = (int)std::min<int64_t>((n + TAIL_THREADS - 1) / TAIL_THREADS, (int64_t)sms * per_sm);
  reparam_tail_kernel<<<grid, TAIL_THREADS, smem, (cudaStream_t)stream>>>(
      n, D, d_xp, d_kind, d_src, d_pre_scale, d_pre_shift, d_scale, d_shift, d_lo, d_hi,
      isnan(log_prior_const) ? 0.0 : log_prior_const, isnan(min_log_q) ? -INFINITY : min_log_q,
      d_logq, d_logw, d_x64, d_stats);
  g_launches += 1;
  CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int nb200_sum_exp(const double* d_logw, int64_t n, const double* d_max,
                             double* d_partials, int n_partials, void* stream) {
  if (!d_logw || !d_max || !d_partials || n_partials < 1 || n_partials > 65535)
    return fail(1, "nb200_sum_exp: bad arguments");
  if (n < 0) n = 0;  // every partial is still written (zeros)
  sum_exp_kernel<<<n_partials, SUMEXP_THREADS, 0, (cudaStream_t)stream>>>(d_logw, n, d_max, d_partials);
  g_launches += 1;
  CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int nb200_coupling_transform(const float* d_x, const float* d_params, float* d_y,
                                        float* d_logdet, int64_t n, int D,
                                        const int32_t* h_transform_features, int d_tr, int additive,
                                        int inverse, void* stream) {
  if (n <= 0) return 0;
  if (!d_x || !d_params || !d_y || !d_logdet || !h_transform_features)
    return fail(1, "nb200_coupling_transform: bad arguments");
  if (D < 1 || D > CP_MAXD || d_tr < 1 || d_tr > D)
    return fail(1, "nb200_coupling_transform: D=%d d_tr=%d out of range", D, d_tr);
  CouplingMap map;
  for (int f = 0; f < CP_MAXD; ++f) map.rank[f] = -1;
  for (int i = 0; i < d_tr; ++i) {
    const int f = h_transform_features[i];
    if (f < 0 || f >= D || map.rank[f] >= 0)
      return fail(1, "nb200_coupling_transform: bad transform feature list");
    map.rank[f] = (int8_t)i;
  }
  int dev = 0, sms = 148;
  CUDA_OK(cudaGetDevice(&dev));
  CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  cudaStream_t st = (cudaStream_t)stream;
  const int lpr = D / 4;
  const bool vec = D % 4 == 0 && (lpr & (lpr - 1)) == 0 && lpr <= 16 &&
                   ((uintptr_t)d_x % 16 == 0) && ((uintptr_t)d_y % 16 == 0);
  if (vec) {
    const int64_t total = n * lpr;
    const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sms * 32);
#define CP_LAUNCH(L) coupling_vec_kernel<L><<<grid, 256, 0, st>>>(d_x, d_params, d_y, d_logdet, n, d_tr, additive, inverse, map)
    if (lpr == 1) CP_LAUNCH(1);
    else if (lpr == 2) CP_LAUNCH(2);
    else if (lpr == 4) CP_LAUNCH(4);
    else if (lpr == 8) CP_LAUNCH(8);
    else CP_LAUNCH(16);
#undef CP_LAUNCH
  } else {
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sms * 32);
    coupling_row_kernel<<<grid, 256, 0, st>>>(d_x, d_params, d_y, d_logdet, n, D, d_tr, additive, inverse, map);
  }
  g_launches += 1;
  CUDA_OK(cudaGetLastError());
  return 0;
}

#ifdef NB200_TC_TRACE
extern "C" int nb200_debug_trace(long long* h_out, int* h_n) {
  CUDA_OK(cudaDeviceSynchronize());
  CUDA_OK(cudaMemcpyFromSymbol(h_out, nb200::tc_trace_buf, sizeof(long long) * 2 * 4096));
  CUDA_OK(cudaMemcpyFromSymbol(h_n, nb200::tc_trace_n, sizeof(int) * 2));
  int zero[2] = {0, 0};
  CUDA_OK(cudaMemcpyToSymbol(nb200::tc_trace_n, zero, sizeof(zero)));
  return 0;
}
#endif

// ============================================================================= peer-memory exchange
#include "xchg.cuh"

extern "C" int nb200_xchg_create(void** d_buf, unsigned char* handle64) {
  if (!d_buf || !handle64) return fail(1, "nb200_xchg_create: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  void* p = nullptr;
  CUDA_OK(cudaMalloc(&p, XCHG_BYTES));
  CUDA_OK(cudaMemset(p, 0, XCHG_BYTES));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(2, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  }
  memcpy(handle64, &h, 64);
  *d_buf = p;
  return 0;
}

extern "C" int nb200_xchg_open(const unsigned char* handle64, void** d_peer) {
  if (!handle64 || !d_peer) return fail(1, "nb200_xchg_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(d_peer, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return fail(2, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int nb200_xchg_close(void* d_peer) {
  if (d_peer) CUDA_OK(cudaIpcCloseMemHandle(d_peer));
  return 0;
}

extern "C" int nb200_xchg_destroy(void* d_buf) {
  if (d_buf) CUDA_OK(cudaFree(d_buf));
  return 0;
}

extern "C" int nb200_xchg_allgather(const void* const* d_peers, int world, int rank, int kind, uint64_t seq,
                                    const void* d_src, int n_words, void* d_gathered, double* d_max_out,
                                    int* d_err, void* stream) {
  if (!d_peers || !d_src || !d_err || world < 1 || world > XCHG_MAXW || rank < 0 || rank >= world ||
      kind < 0 || kind >= XCHG_KINDS || n_words < 1 || n_words > XCHG_WORDS || seq == 0)
    return fail(1, "nb200_xchg_allgather: bad arguments");
  xchg_allgather_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(
      (void* const*)d_peers, world, rank, kind, (unsigned long long)seq, (const unsigned long long*)d_src, n_words,
      (unsigned long long*)d_gathered, d_max_out, d_err);
  g_launches += 1;
  CUDA_OK(cudaGetLastError());
  return 0;
}

// ============================================================================= training
#include "train.cuh"

struct nb200_trainer {
  TrPlan h_plan;
  TrPlan* d_plan = nullptr;
  int* d_itab = nullptr;
  int* d_reduce = nullptr;
  int num_sms = 0;
  int cap_tiles = 0;
  float *ws = nullptr, *dout0 = nullptr, *dout1 = nullptr, *ldrow = nullptr, *crow = nullptr;
  float *stat_part = nullptr, *stats = nullptr, *s_part0 = nullptr, *s_part1 = nullptr;
  float *wsum_part = nullptr, *loss_part = nullptr, *part = nullptr, *grad = nullptr;
  float *gn_part = nullptr, *eval_part = nullptr, *pmask = nullptr;
  float* stat_n = nullptr;
  float* run_scalars = nullptr;  // [0] loss accumulator of nb200_train_run
  unsigned* bar = nullptr;       // grid barrier counter of the persistent kernel
  long long* trace = nullptr;    // NB200_TR_TRACE: phase timeline of CTA 0
  size_t smem_fwd = 0, smem_bwd = 0;
};
constexpr int TR_TRACE_CAP = 16384;

static void trainer_free_rows(nb200_trainer* t) {
  cudaFree(t->ws), cudaFree(t->dout0), cudaFree(t->dout1), cudaFree(t->ldrow), cudaFree(t->crow);
  t->ws = t->dout0 = t->dout1 = t->ldrow = t->crow = nullptr;
  t->cap_tiles = 0;
}

extern "C" int nb200_trainer_destroy(nb200_trainer* t) {
  if (!t) return 0;
  trainer_free_rows(t);
  cudaFree(t->d_plan), cudaFree(t->d_itab), cudaFree(t->d_reduce);
  cudaFree(t->stat_part), cudaFree(t->stats), cudaFree(t->s_part0), cudaFree(t->s_part1);
  cudaFree(t->wsum_part), cudaFree(t->loss_part), cudaFree(t->part), cudaFree(t->grad);
  cudaFree(t->gn_part), cudaFree(t->eval_part), cudaFree(t->stat_n), cudaFree(t->pmask);
  cudaFree(t->run_scalars), cudaFree(t->bar), cudaFree(t->trace);
  delete t;
  return 0;
}

extern "C" int nb200_trainer_create(nb200_trainer** out, const int32_t* h_plan, int n_plan_ints,
                                    const int32_t* h_itab, int n_itab,
                                    const int32_t* h_reduce_idx, int n_reduce) {
  if (!out || !h_plan || !h_itab || !h_reduce_idx)
    return fail(1, "nb200_trainer_create: bad arguments");
  if (n_plan_ints != TR_PLAN_INTS)
    return fail(1, "nb200_trainer_create: plan has %d ints, expected %d", n_plan_ints, TR_PLAN_INTS);
  int dev = 0;
  CUDA_OK(cudaGetDevice(&dev));
  DeviceInfo prop;
  if (int rc = device_info(dev, prop)) return rc;
  if (prop.major != 10)
    return fail(3, "nessai_b200 kernels are built for sm_100a only; device %d is sm_%d%d", dev,
                prop.major, prop.minor);
  nb200_trainer* t = new (std::nothrow) nb200_trainer();
  if (!t) return fail(4, "out of host memory");
  memcpy(&t->h_plan, h_plan, sizeof(TrPlan));
  const TrPlan& P = t->h_plan;
  if (P.D < 2 || P.D > TR_MAXD || P.L < 1 || P.L > TR_MAXL || P.n_itab != n_itab ||
      P.n_reduce != n_reduce || P.max_dim < P.D || P.max_in < P.D) {
    delete t;
    return fail(1, "nb200_trainer_create: plan out of range (D=%d L=%d)", P.D, P.L);
  }
  for (int l = 0; l < P.L; ++l) {
    const TrLayer& ly = P.layer[l];
    if (ly.n_lin < 1 || ly.n_lin > TR_MAXLIN || ly.n_buf < 2 || ly.n_buf > TR_MAXBUF) {
      delete t;
      return fail(1, "nb200_trainer_create: layer %d has %d linears / %d buffers", l, ly.n_lin, ly.n_buf);
    }
  }
  t->num_sms = prop.sms;
  t->smem_fwd = 4 * tr_smem_floats(P.D, P.vals_floats, P.max_in, P.wmax, P.n_itab, false);
  t->smem_bwd = 4 * tr_smem_floats(P.D, P.vals_floats, P.max_in, P.wmax, P.n_itab, true);
  if (t->smem_bwd > 226 * 1024) {
    delete t;
    return fail(5, "flow too large for the training kernels (%zu KB shared memory)", t->smem_bwd / 1024);
  }
  const int G = TR_MAXG;
#define TR_ALLOC(ptr, count) CUDA_OK(cudaMalloc(&(ptr), sizeof(*(ptr)) * (size_t)(count)))
  TR_ALLOC(t->d_plan, 1);
  TR_ALLOC(t->d_itab, n_itab > 0 ? n_itab : 1);
  TR_ALLOC(t->d_reduce, n_reduce > 0 ? n_reduce : 1);
  TR_ALLOC(t->stat_part, (size_t)P.L * G * 2 * P.D);
  TR_ALLOC(t->stats, (size_t)P.L * 2 * P.D);
  TR_ALLOC(t->s_part0, (size_t)G * 2 * P.D);
  TR_ALLOC(t->s_part1, (size_t)G * 2 * P.D);
  TR_ALLOC(t->wsum_part, G);
  TR_ALLOC(t->loss_part, G);
  const size_t part_stride = ((size_t)P.n_part + 3) & ~(size_t)3, n_grad = ((size_t)P.n_params + 3) & ~(size_t)3;
  TR_ALLOC(t->part, (size_t)G * part_stride);
  TR_ALLOC(t->grad, n_grad);
  TR_ALLOC(t->gn_part, G);
  CUDA_OK(cudaMemset(t->part, 0, sizeof(float) * G * part_stride));
  TR_ALLOC(t->eval_part, 2 * 2 * prop.sms);
  TR_ALLOC(t->stat_n, G);
  TR_ALLOC(t->run_scalars, 4);
  TR_ALLOC(t->bar, 4);
  CUDA_OK(cudaMemset(t->run_scalars, 0, 4 * sizeof(float)));
  if (getenv("NB200_TR_TRACE")) TR_ALLOC(t->trace, 2 * TR_TRACE_CAP + 1);
  CUDA_OK(cudaMemcpy(t->d_plan, &t->h_plan, sizeof(TrPlan), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(t->d_itab, h_itab, sizeof(int) * n_itab, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(t->d_reduce, h_reduce_idx, sizeof(int) * n_reduce, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemset(t->grad, 0, sizeof(float) * n_grad));
  *out = t;
  return 0;
}

static int trainer_reserve(nb200_trainer* t, int n_tiles) {
  if (n_tiles <= t->cap_tiles) return 0;
  CUDA_OK(cudaDeviceSynchronize());
  trainer_free_rows(t);
  const TrPlan& P = t->h_plan;
  TR_ALLOC(t->ws, (size_t)n_tiles * P.rec_total * TR_R);
  TR_ALLOC(t->dout0, (size_t)n_tiles * P.D * TR_R);
  TR_ALLOC(t->dout1, (size_t)n_tiles * P.D * TR_R);
  TR_ALLOC(t->ldrow, (size_t)n_tiles * TR_R);
  TR_ALLOC(t->crow, (size_t)n_tiles * TR_R);
  t->cap_tiles = n_tiles;
  return 0;
}

static TrBuffers trainer_buffers(nb200_trainer* t, float* theta_p, float* theta_b, int G) {
  TrBuffers B;
  B.itab = t->d_itab;
  B.reduce_idx = t->d_reduce;
  B.theta_p = theta_p;
  B.theta_b = theta_b;
  B.ws = t->ws;
  B.dout[0] = t->dout0;
  B.dout[1] = t->dout1;
  B.ldrow = t->ldrow;
  B.crow = t->crow;
  B.stat_part = t->stat_part;
  B.stat_n = t->stat_n;
  B.stats = t->stats;
  B.s_part[0] = t->s_part0;
  B.s_part[1] = t->s_part1;
  B.wsum_part = t->wsum_part;
  B.loss_part = t->loss_part;
  B.part = t->part;
  B.part_stride = (t->h_plan.n_part + 3) & ~3;
  B.grad = t->grad;
  B.gn_part = t->gn_part;
  B.G = G;
  B.pmask = t->pmask;
  B.n_reduce_blocks = G;  // the REDUCE phase runs on the row kernels' grid
  return B;
}

extern "C" int nb200_trainer_set_itab(nb200_trainer* t, const int32_t* h_itab, int n_itab, void* stream) {
  if (!t || !h_itab || n_itab != t->h_plan.n_itab) return fail(1, "nb200_trainer_set_itab: bad arguments");
  // (stream-ordered against the trainer's previous launches; the host array is pageable: the copy
  // has left it when the call returns)
  CUDA_OK(cudaMemcpyAsync(t->d_itab, h_itab, sizeof(int) * n_itab, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return 0;
}

extern "C" int nb200_trainer_set_param_mask(nb200_trainer* t, const float* h_mask) {
  if (!t) return fail(1, "nb200_trainer_set_param_mask: bad arguments");
  if (!h_mask) {
    cudaFree(t->pmask);
    t->pmask = nullptr;
    return 0;
  }
  if (!t->pmask) TR_ALLOC(t->pmask, t->h_plan.n_params);
  CUDA_OK(cudaMemcpy(t->pmask, h_mask, sizeof(float) * t->h_plan.n_params, cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int nb200_trainer_copy_grad(nb200_trainer* t, float* d_out, void* stream) {
  if (!t || !d_out) return fail(1, "nb200_trainer_copy_grad: bad arguments");
  CUDA_OK(cudaMemcpyAsync(d_out, t->grad, sizeof(float) * t->h_plan.n_params, cudaMemcpyDeviceToDevice,
                          (cudaStream_t)stream));
  return 0;
}

// Cooperative launch of the persistent kernel over `n_epochs` epochs (or one gradient step).
static int trainer_launch_run(nb200_trainer* t, float* d_theta_p, float* d_theta_b, TrRun& R, cudaStream_t st) {
  const TrPlan& P = t->h_plan;
  const int max_b = (int)std::min<int64_t>(R.batch_size, R.n_rows);
  int max_tiles = (max_b + TR_R - 1) / TR_R;
  if (int rc = trainer_reserve(t, max_tiles)) return rc;
  if (int rc = prep_kernel(tr_train_kernel, t->smem_bwd)) return rc;
  if (R.n_val > 0) max_tiles = std::max<int>(max_tiles, (int)((R.n_val + TR_R - 1) / TR_R));
  const int G = std::max(1, std::min(max_tiles, std::min(TR_MAXG, t->num_sms)));
  TrBuffers B = trainer_buffers(t, d_theta_p, d_theta_b, G);
  R.eval_part = t->eval_part;
  R.bar = t->bar;
  B.trace = t->trace;
  B.trace_cap = t->trace ? TR_TRACE_CAP : 0;
  CUDA_OK(cudaMemsetAsync(t->bar, 0, sizeof(unsigned), st));
  if (t->trace) CUDA_OK(cudaMemsetAsync(t->trace, 0, sizeof(long long) * (2 * TR_TRACE_CAP + 1), st));
  void* args[] = {(void*)&t->h_plan, (void*)&B, (void*)&R};
  CUDA_OK(cudaLaunchCooperativeKernel((const void*)tr_train_kernel, dim3(G), dim3(TR_THREADS), args, t->smem_bwd, st));
  g_launches += 1;
  if (t->trace) {
    // debugging aid: phase timeline of CTA 0, "tag ns-since-start" per line
    static std::vector<long long> h(2 * TR_TRACE_CAP);
    CUDA_OK(cudaStreamSynchronize(st));
    CUDA_OK(cudaMemcpy(h.data(), t->trace, sizeof(long long) * 2 * TR_TRACE_CAP, cudaMemcpyDeviceToHost));
    if (FILE* f = fopen(getenv("NB200_TR_TRACE"), "a")) {
      fprintf(f, "# run G=%d epochs=%d rows=%lld batch=%d\n", G, R.n_epochs, (long long)R.n_rows, R.batch_size);
      for (int i = 0; i < TR_TRACE_CAP && (i == 0 || h[2 * i + 1]); ++i)
        fprintf(f, "%lld %lld\n", h[2 * i], h[2 * i + 1] - h[1]);
      fclose(f);
    }
  }
  return 0;
}

static int trainer_check_args(nb200_trainer* t, float* d_theta_p, float* d_theta_b, float* d_m, float* d_v,
                              const float* d_x, int64_t n_rows, int batch_size, int opt_kind, const char* who) {
  if (!t || !d_theta_p || !d_x || n_rows < 1 || batch_size < 1) return fail(1, "%s: bad arguments", who);
  if (opt_kind < -1 || opt_kind > 2) return fail(1, "%s: unknown optimiser %d", who, opt_kind);
  if (opt_kind >= 0 && opt_kind <= 1 && (!d_m || !d_v)) return fail(1, "%s: Adam needs its moment buffers", who);
  const TrPlan& P = t->h_plan;
  bool any_bn = false;
  for (int l = 0; l < P.L; ++l) any_bn |= P.layer[l].bn_uw >= 0;
  if (any_bn && !d_theta_b) return fail(1, "%s: BatchNorm buffers missing", who);
  const int64_t last = n_rows % batch_size;
  if (any_bn && (std::min<int64_t>(batch_size, n_rows) < 2 || last == 1))
    return fail(1, "%s: a batch of one row has no batch variance", who);
  return 0;
}

extern "C" int nb200_train_epoch(nb200_trainer* t, float* d_theta_p, float* d_theta_b, float* d_m,
                                 float* d_v, const float* d_x, const float* d_w,
                                 const int64_t* d_perm, int64_t n_rows, int batch_size,
                                 int opt_kind, double lr, double beta1, double beta2, double eps,
                                 double weight_decay, double clip, int64_t step0,
                                 float* d_loss_sum, float* d_step_info, void* stream) {
  if (int rc = trainer_check_args(t, d_theta_p, d_theta_b, d_m, d_v, d_x, n_rows, batch_size, opt_kind,
                                  "nb200_train_epoch"))
    return rc;
  const TrPlan& P = t->h_plan;
  cudaStream_t st = (cudaStream_t)stream;
  // NB200_TR_CHAIN=1: one launch per phase (the pre-persistent-kernel path; debugging aid)
  static const bool chain = getenv("NB200_TR_CHAIN") != nullptr;
  if (!chain) {
    TrRun R{};
    R.x = d_x, R.w = d_w, R.perm = d_perm, R.n_rows = n_rows, R.batch_size = batch_size;
    R.n_epochs = 1, R.kind = opt_kind;
    R.beta1 = (float)beta1, R.beta2 = (float)beta2, R.eps = (float)eps, R.weight_decay = (float)weight_decay;
    R.clip = (float)clip, R.beta1d = beta1, R.beta2d = beta2, R.step0 = step0, R.lr[0] = (float)lr;
    R.b1pow0 = std::pow(beta1, (double)step0), R.b2pow0 = std::pow(beta2, (double)step0);
    R.m = d_m, R.v = d_v, R.step_info = d_step_info, R.loss_accum = d_loss_sum;
    return trainer_launch_run(t, d_theta_p, d_theta_b, R, st);
  }
  const int max_b = (int)std::min<int64_t>(batch_size, n_rows);
  if (int rc = trainer_reserve(t, (max_b + TR_R - 1) / TR_R)) return rc;
  if (int rc = prep_kernel(tr_fwd_kernel, t->smem_fwd)) return rc;
  if (int rc = prep_kernel(tr_loss_kernel, t->smem_fwd)) return rc;
  if (int rc = prep_kernel(tr_bwd_kernel, t->smem_bwd)) return rc;
  if (t->pmask) {
    tr_mask_params_kernel<<<std::min((P.n_params + 255) / 256, 2 * t->num_sms), 256, 0, st>>>(d_theta_p, t->pmask,
                                                                                                 P.n_params);
    g_launches += 1;
  }
  int64_t step = step0;
  int ib = 0;
  for (int64_t i0 = 0; i0 < n_rows; i0 += batch_size, ++ib) {
    TrBatch bt;
    bt.x = d_x;
    bt.perm = d_perm;
    bt.w = d_w;
    bt.i0 = i0;
    bt.B = (int)std::min<int64_t>(batch_size, n_rows - i0);
    bt.n_tiles = (bt.B + TR_R - 1) / TR_R;
    const int G = std::min(bt.n_tiles, std::min(TR_MAXG, t->num_sms));
    TrBuffers B = trainer_buffers(t, d_theta_p, d_theta_b, G);
    for (int l = 0; l < P.L; ++l) tr_fwd_kernel<<<G, TR_THREADS, t->smem_fwd, st>>>(P, B, bt, l);
    tr_loss_kernel<<<G, TR_THREADS, t->smem_fwd, st>>>(P, B, bt);
    for (int l = P.L - 1; l >= 0; --l) tr_bwd_kernel<<<G, TR_THREADS, t->smem_bwd, st>>>(P, B, bt, l);
    tr_reduce_kernel<<<G, TR_THREADS, 0, st>>>(P, B);
    ++step;
    TrOptim o;
    o.kind = opt_kind;
    o.lr = (float)lr;
    o.beta1 = (float)beta1;
    o.beta2 = (float)beta2;
    o.eps = (float)eps;
    o.weight_decay = (float)weight_decay;
    o.clip = (float)clip;
    o.bc1 = (float)(1.0 - std::pow(beta1, (double)step));
    o.bc2 = (float)(1.0 - std::pow(beta2, (double)step));
    const int ablocks = std::min((P.n_params + 255) / 256, 2 * t->num_sms);
    tr_adam_kernel<<<ablocks, 256, 0, st>>>(B, P.n_params, o, d_m, d_v, d_step_info ? d_step_info + 2 * ib : nullptr,
                                            d_loss_sum);
    g_launches += 2 * P.L + 3;
  }
  CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int nb200_train_run(nb200_trainer* t, float* d_theta_p, float* d_theta_b, int64_t n_theta_b,
                               float* d_m, float* d_v, const float* d_x, const float* d_w,
                               const int64_t* d_perm, int64_t n_rows, int batch_size, const float* d_xv,
                               const float* d_wv, int64_t n_val, int n_epochs, int epoch0, int validate,
                               int patience, int opt_kind, const double* h_lr, double beta1, double beta2,
                               double eps, double weight_decay, double clip, int64_t step0, float* d_hist,
                               void* d_ctl, float* d_best_p, float* d_best_b, void* stream) {
  if (int rc = trainer_check_args(t, d_theta_p, d_theta_b, d_m, d_v, d_x, n_rows, batch_size, opt_kind,
                                  "nb200_train_run"))
    return rc;
  if (n_epochs < 1 || n_epochs > TR_MAX_CHUNK) return fail(1, "nb200_train_run: 1..%d epochs per call", TR_MAX_CHUNK);
  if (opt_kind < 0 || !h_lr || !d_hist || !d_ctl) return fail(1, "nb200_train_run: bad arguments");
  if (validate && (!d_best_p || (n_theta_b > 0 && !d_best_b))) return fail(1, "nb200_train_run: snapshot buffers missing");
  if (n_val > 0 && !d_xv) return fail(1, "nb200_train_run: validation rows missing");
  if (n_val > (int64_t)1 << 30) return fail(1, "nb200_train_run: too many validation rows");
  TrRun R{};
  R.x = d_x, R.w = d_w, R.perm = d_perm, R.n_rows = n_rows, R.batch_size = batch_size;
  R.xv = d_xv, R.wv = d_wv, R.n_val = n_val;
  R.n_epochs = n_epochs, R.epoch0 = epoch0, R.validate = validate, R.patience = patience, R.kind = opt_kind;
  R.beta1 = (float)beta1, R.beta2 = (float)beta2, R.eps = (float)eps, R.weight_decay = (float)weight_decay;
  R.clip = (float)clip, R.beta1d = beta1, R.beta2d = beta2, R.step0 = step0;
  for (int e = 0; e < n_epochs; ++e) R.lr[e] = (float)h_lr[e];
  R.b1pow0 = std::pow(beta1, (double)step0), R.b2pow0 = std::pow(beta2, (double)step0);
  R.m = d_m, R.v = d_v, R.hist = d_hist, R.loss_accum = t->run_scalars;
  R.ctl = (TrCtl*)d_ctl, R.best_p = d_best_p, R.best_b = d_best_b, R.n_b = (int)n_theta_b;
  return trainer_launch_run(t, d_theta_p, d_theta_b, R, (cudaStream_t)stream);
}

extern "C" int nb200_eval_loss(nb200_trainer* t, float* d_theta_p, float* d_theta_b, const float* d_x,
                               const float* d_w, int64_t n, float* d_loss, float* d_logp,
                               void* stream) {
  if (!t || !d_theta_p || !d_x || n < 1 || (!d_loss && !d_logp))
    return fail(1, "nb200_eval_loss: bad arguments");
  if (n > (int64_t)1 << 30) return fail(1, "nb200_eval_loss: too many rows");
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = prep_kernel(tr_eval_kernel, t->smem_fwd)) return rc;
  TrBatch bt;
  bt.x = d_x;
  bt.perm = nullptr;
  bt.w = d_w;
  bt.i0 = 0;
  bt.B = (int)n;
  bt.n_tiles = (int)((n + TR_R - 1) / TR_R);
  const int G = std::min(bt.n_tiles, 2 * t->num_sms);
  TrBuffers B = trainer_buffers(t, d_theta_p, d_theta_b, G);
  if (t->pmask) {
    tr_mask_params_kernel<<<std::min((t->h_plan.n_params + 255) / 256, 2 * t->num_sms), 256, 0, st>>>(
        d_theta_p, t->pmask, t->h_plan.n_params);
    g_launches += 1;
  }
  tr_eval_kernel<<<G, TR_THREADS, t->smem_fwd, st>>>(t->h_plan, B, bt, t->eval_part, d_logp);
  if (d_loss) tr_eval_final_kernel<<<1, 32, 0, st>>>(t->eval_part, G, d_loss);
  g_launches += d_loss ? 2 : 1;
  CUDA_OK(cudaGetLastError());
  return 0;
}
