// C ABI + generic fp32 kernels of nessai_b200 (see include/nessai_b200.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <new>

#include "../../include/nessai_b200.h"
#include "flow_interp.cuh"
#include "philox.cuh"
#include "populate_common.cuh"
#include "flow_tc.cuh"

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
};

struct nb200_flow {
  int D, H, activation;
  int device;
  int num_sms;
  DirProgram dir[2];
};

static void free_dir(DirProgram& p) {
  if (p.d_ops) cudaFree(p.d_ops);
  if (p.d_blob) cudaFree(p.d_blob);
  delete[] p.h_ops;
  tc_free(p.tc);
  p = DirProgram();
}

extern "C" int nb200_flow_create(nb200_flow** out, int D, int H, int activation) {
  if (!out) return fail(1, "nb200_flow_create: out is NULL");
  if (D < 1 || D > 1024 || H < 1 || H > 4096) return fail(1, "nb200_flow_create: bad sizes D=%d H=%d", D, H);
  if (activation < 0 || activation > 2) return fail(1, "nb200_flow_create: bad activation %d", activation);
  int dev = 0;
  CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return fail(3, "nessai_b200 kernels are built for sm_100a only; device %d is sm_%d%d", dev,
                prop.major, prop.minor);
  nb200_flow* f = new (std::nothrow) nb200_flow();
  if (!f) return fail(4, "out of host memory");
  f->D = D;
  f->H = H;
  f->activation = activation;
  f->device = dev;
  f->num_sms = prop.multiProcessorCount;
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
                  int lp_mode) {
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
        const float c = 0.5f * P.D * LOG_2PI;
        out_lp[row] = (lp_mode == 1) ? (-0.5f * ss_in - c) - ld : (-0.5f * ss_out - c) + ld;
      }
    }
  }
}

__global__ void sample_latent_kernel(float* __restrict__ z, int64_t n, int D, uint64_t seed,
                                     uint64_t row_offset) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  for (int d0 = 0; d0 < D; d0 += 4) {
    const Philox4 r = philox4x32_10(seed, row_offset + row, d0 / 4, 0);
    float v[4];
    box_muller(r.x, r.y, v[0], v[1]);
    box_muller(r.z, r.w, v[2], v[3]);
    for (int j = 0; j < 4 && d0 + j < D; ++j) z[row * D + d0 + j] = v[j];
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
#define ACC_CHUNK 1024
#define ACC_THREADS 256

__device__ __forceinline__ bool accept_row(const double* __restrict__ logw, double mx,
                                           uint64_t seed, uint64_t grow, int64_t row, int64_t n) {
  if (row >= n) return false;
  const double lw = logw[row];
  if (isnan(lw)) return false;
  const Philox4 r = philox4x32_10(seed, grow, 0, 1);
  const double u = ((double)r.x + 0.5) * 2.3283064365386963e-10;
  return (lw - mx) > log(u);
}

__global__ void __launch_bounds__(ACC_THREADS)
accept_count_kernel(const double* __restrict__ logw, const double* __restrict__ d_max, int64_t n,
                    uint64_t seed, uint64_t row_offset, int64_t* __restrict__ scratch) {
  __shared__ int wsum[ACC_THREADS / 32];
  const double mx = *d_max;
  const int64_t base = (int64_t)blockIdx.x * ACC_CHUNK + threadIdx.x * 4;
  int c = 0;
  for (int j = 0; j < 4; ++j) c += accept_row(logw, mx, seed, row_offset + base + j, base + j, n);
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < ACC_THREADS / 32; ++i) t += wsum[i];
    scratch[blockIdx.x] = t;
  }
}

// exclusive scan of the chunk counts (single block)
__global__ void __launch_bounds__(1024)
accept_scan_kernel(int64_t* __restrict__ scratch, int64_t nchunks, int64_t capacity,
                   int64_t* __restrict__ counts) {
  __shared__ int64_t wsum[32];
  __shared__ int64_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < nchunks; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int64_t v = i < nchunks ? scratch[i] : 0;
    int64_t incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      int64_t w = wsum[threadIdx.x];
      int64_t wi = w;
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, wi, o);
        if (threadIdx.x >= o) wi += t;
      }
      wsum[threadIdx.x] = wi - w;  // exclusive warp offsets
    }
    __syncthreads();
    const int64_t carry = carry_s;
    const int64_t excl = carry + wsum[threadIdx.x >> 5] + incl - v;
    if (i < nchunks) scratch[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int64_t total = carry_s;
    counts[0] = total;
    counts[1] = total < capacity ? total : capacity;
  }
}

struct RowFormat {
  int row_words;
  int D;
  int logp_off;  // bytes, < 0: skip
  int off[256];  // byte offsets of the D parameters
};

__global__ void __launch_bounds__(ACC_THREADS)
accept_write_kernel(const float* __restrict__ xp, const double* __restrict__ scale,
                    const double* __restrict__ shift, const double* __restrict__ logw,
                    const double* __restrict__ d_max, int64_t n, uint64_t seed,
                    uint64_t row_offset, const int64_t* __restrict__ scratch, double logp,
                    const uint32_t* __restrict__ tmpl, RowFormat F, uint32_t* __restrict__ rows,
                    int64_t capacity, int64_t write_offset) {
  __shared__ int wsum[ACC_THREADS / 32];
  const double mx = *d_max;
  const int64_t base = (int64_t)blockIdx.x * ACC_CHUNK + threadIdx.x * 4;
  bool acc[4];
  int c = 0;
  for (int j = 0; j < 4; ++j) {
    acc[j] = accept_row(logw, mx, seed, row_offset + base + j, base + j, n);
    c += acc[j];
  }
  int incl = c;
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += t;
  }
  if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
  __syncthreads();
  int woff = 0;
  for (int i = 0; i < (threadIdx.x >> 5); ++i) woff += wsum[i];
  int64_t idx = scratch[blockIdx.x] + woff + incl - c;
  for (int j = 0; j < 4; ++j) {
    if (!acc[j]) continue;
    if (idx < capacity) {
      uint32_t* dst = rows + (write_offset + idx) * F.row_words;
      for (int w = 0; w < F.row_words; ++w) dst[w] = tmpl[w];
      const float* xr = xp + (base + j) * F.D;
      for (int d = 0; d < F.D; ++d) {
        // same float64 arithmetic as the bounds check of the draw kernel
        const unsigned long long b = __double_as_longlong((double)xr[d] * scale[d] + shift[d]);
        dst[F.off[d] / 4] = (uint32_t)b;
        dst[F.off[d] / 4 + 1] = (uint32_t)(b >> 32);
      }
      if (F.logp_off >= 0) {
        const unsigned long long b = __double_as_longlong(logp);
        dst[F.logp_off / 4] = (uint32_t)b;
        dst[F.logp_off / 4 + 1] = (uint32_t)(b >> 32);
      }
    }
    ++idx;
  }
}

// ----------------------------------------------------------------------------- launch helpers
template <typename K>
static int prep_kernel(K kernel, size_t smem) {
  // opt in to > 48 KB dynamic shared memory (static shared memory counts
  // against the 227 KB limit, so ask for what the launch needs, rounded up)
  static thread_local const void* done[16];
  static thread_local size_t done_smem[16];
  static thread_local int ndone = 0;
  int slot = -1;
  for (int i = 0; i < ndone; ++i)
    if (done[i] == (const void*)kernel) slot = i;
  if (slot >= 0 && done_smem[slot] >= smem) return 0;
  CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (slot < 0 && ndone < 16) slot = ndone++;
  if (slot >= 0) {
    done[slot] = (const void*)kernel;
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
  if (p.tc.valid && tc_enabled()) {
    g_launches += 1;
    return tc_launch_apply(p.tc, in, out, logj, lp, n, direction == 1 ? 1 : 2, f->num_sms, st)
               ? fail(2, "tc kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()))
               : 0;
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
    flow_apply_kernel<ACT><<<grid, BS, smem, st>>>(P, in, out, logj, lp, n, lp_mode);   \
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

extern "C" int nb200_sample_latent(float* d_z, int64_t n, int D, uint64_t seed,
                                   uint64_t row_offset, void* stream) {
  if (n <= 0) return 0;
  if (!d_z || D < 1) return fail(1, "nb200_sample_latent: bad arguments");
  const int bs = 256;
  sample_latent_kernel<<<(unsigned)((n + bs - 1) / bs), bs, 0, (cudaStream_t)stream>>>(
      d_z, n, D, seed, row_offset);
  g_launches += 1;
  CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int nb200_populate_draw(nb200_flow* f, int64_t n, uint64_t seed, uint64_t row_offset,
                                   float r_max, float sqrt_temperature, const double* d_scale,
                                   const double* d_shift, const double* d_lo, const double* d_hi,
                                   double log_prior_const, float* d_xp, double* d_logq,
                                   double* d_logw, float* d_z, double* d_stats, void* stream) {
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
  A.sqrt_t = sqrt_temperature > 0.f ? sqrt_temperature : 1.f;
  A.scale = d_scale;
  A.shift = d_shift;
  A.lo = d_lo;
  A.hi = d_hi;
  A.log_prior_const = log_prior_const;
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

extern "C" int nb200_populate_accept(int64_t n, int D, const float* d_xp, const double* d_scale,
                                     const double* d_shift, const double* d_logw,
                                     const double* d_max, uint64_t seed, uint64_t row_offset,
                                     double log_p_value, const uint8_t* d_row_template,
                                     int row_bytes, const int32_t* h_field_offsets,
                                     uint8_t* d_rows, int64_t capacity, int64_t write_offset,
                                     int64_t* d_counts, int64_t* d_scratch, void* stream) {
  if (!d_xp || !d_scale || !d_shift || !d_logw || !d_max || !d_row_template || !h_field_offsets || !d_rows || !d_counts ||
      !d_scratch)
    return fail(1, "nb200_populate_accept: bad arguments");
  if (row_bytes % 4 || row_bytes <= 0) return fail(1, "row_bytes must be a positive multiple of 4");
  if (D < 1 || D > 256) return fail(1, "nb200_populate_accept: D out of range");
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
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
  const int64_t nchunks = (n + ACC_CHUNK - 1) / ACC_CHUNK;
  accept_count_kernel<<<(unsigned)nchunks, ACC_THREADS, 0, st>>>(d_logw, d_max, n, seed,
                                                                 row_offset, d_scratch);
  accept_scan_kernel<<<1, 1024, 0, st>>>(d_scratch, nchunks, capacity, d_counts);
  accept_write_kernel<<<(unsigned)nchunks, ACC_THREADS, 0, st>>>(
      d_xp, d_scale, d_shift, d_logw, d_max, n, seed, row_offset, d_scratch, log_p_value,
      reinterpret_cast<const uint32_t*>(d_row_template), F, reinterpret_cast<uint32_t*>(d_rows),
      capacity, write_offset);
  g_launches += 3;
  CUDA_OK(cudaGetLastError());
  return 0;
}
