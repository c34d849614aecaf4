// Rejection step + in-order compaction of one populate turn (flowproposal.py:491-506), one
// single-pass kernel.  Lives in a header (included once, by nessai_b200.cu) so that the very same
// source can also be compiled by g++ against the CPU SIMT shim of tests/_hostcheck, which defines
// NB200_SIMT_SHIM: the two spots below that cannot compile for the host (the relaxed
// gpu-scope PTX load / store of the look-back words, the dynamic shared-memory declaration) have
// a host flavour there; under nvcc the text is what it always was.
#pragma once
#include <cstdint>

#include "philox.cuh"

#ifdef NB200_SIMT_SHIM
#define NB200_DYNAMIC_SMEM(name) unsigned char* name = simt::dynamic_smem
#else
#define NB200_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

using namespace nb200;

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

struct RowFormat {
  int row_words;
  int D;
  int logp_off;  // bytes, < 0: skip
  int off[256];  // byte offsets of the D parameters
};

// Chunk status word of the single-pass scan: flag (2 bits) | count (62 bits).
constexpr unsigned long long ACC_FLAG_AGG = 1ull << 62, ACC_FLAG_INCL = 2ull << 62,
                             ACC_VALUE_MASK = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long acc_ld(const unsigned long long* p) {
#ifdef NB200_SIMT_SHIM
  return *reinterpret_cast<const volatile unsigned long long*>(p);
#else
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
#endif
}
__device__ __forceinline__ void acc_st(unsigned long long* p, unsigned long long v) {
#ifdef NB200_SIMT_SHIM
  *reinterpret_cast<volatile unsigned long long*>(p) = v;
#else
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#endif
}

// Rejection step + in-order compaction in ONE pass (flowproposal.py:492-506): a chunk of
// ACC_CHUNK rows per block; the exclusive offset of a chunk comes from a decoupled look-back over
// the status words of its predecessors (chunks are taken in ticket order, so a predecessor is
// always running or done); accepted rows are written as live-point records by the whole warp
// (one coalesced 4-byte word per lane) in draw order.  status[0 .. nchunks-1] and the ticket
// status[nchunks] are zeroed by the launcher.
__global__ void __launch_bounds__(ACC_THREADS)
accept_fused_kernel(const float* __restrict__ xp, const double* __restrict__ x64,
                    const double* __restrict__ scale,
                    const double* __restrict__ shift, const double* __restrict__ logw,
                    const double* __restrict__ logl, int logl_off,
                    const double* __restrict__ d_max, int64_t n, uint64_t seed, uint64_t row_offset,
                    unsigned long long* __restrict__ status, int64_t nchunks, double logp,
                    const uint32_t* __restrict__ tmpl, RowFormat F, uint32_t* __restrict__ rows,
                    int64_t capacity, int64_t write_offset, int64_t* __restrict__ counts) {
  NB200_DYNAMIC_SMEM(acc_smem);
  double* scale_s = reinterpret_cast<double*>(acc_smem);          // [D]
  double* shift_s = scale_s + F.D;                                 // [D]
  uint32_t* tmpl_s = reinterpret_cast<uint32_t*>(shift_s + F.D);   // [row_words]
  short* src_s = reinterpret_cast<short*>(tmpl_s + F.row_words);   // [row_words]: 2 d + half, -1: template,
                                                                   // -2 / -3: low / high word of logL
  __shared__ int wsum[ACC_THREADS / 32];
  __shared__ int64_t chunk_s, prefix_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) chunk_s = (int64_t)atomicAdd(reinterpret_cast<unsigned int*>(status + nchunks), 1u);
  for (int w = tid; w < F.row_words; w += ACC_THREADS) {
    tmpl_s[w] = tmpl[w];
    src_s[w] = -1;
  }
  for (int d = tid; d < F.D; d += ACC_THREADS) scale_s[d] = scale[d], shift_s[d] = shift[d];
  __syncthreads();
  for (int d = tid; d < F.D; d += ACC_THREADS) {
    src_s[F.off[d] / 4] = (short)(2 * d);
    src_s[F.off[d] / 4 + 1] = (short)(2 * d + 1);
  }
  if (tid == 0 && F.logp_off >= 0) {
    const unsigned long long b = __double_as_longlong(logp);
    tmpl_s[F.logp_off / 4] = (uint32_t)b;
    tmpl_s[F.logp_off / 4 + 1] = (uint32_t)(b >> 32);
  }
  if (tid == 0 && logl && logl_off >= 0) {
    src_s[logl_off / 4] = -2;
    src_s[logl_off / 4 + 1] = -3;
  }
  const int64_t chunk = chunk_s;
  const double mx = *d_max;
  const int64_t base = chunk * ACC_CHUNK + tid * 4;
  unsigned accbits = 0;
  int c = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const bool a = accept_row(logw, mx, seed, row_offset + base + j, base + j, n);
    accbits |= (unsigned)a << j;
    c += a;
  }
  int incl = c;
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int tot = 0;
    for (int i = 0; i < ACC_THREADS / 32; ++i) tot += wsum[i];
    if (lane == 0) acc_st(status + chunk, (chunk == 0 ? ACC_FLAG_INCL : ACC_FLAG_AGG) | (unsigned long long)tot);
    // look back: lanes read the 32 predecessors below i, nearest first
    int64_t excl = 0;
    int64_t i = chunk - 1;
    while (i >= 0) {
      const int64_t idx = i - lane;
      const unsigned long long sw = idx >= 0 ? acc_ld(status + idx) : ACC_FLAG_INCL;
      const unsigned flag = (unsigned)(sw >> 62);
      const unsigned m_incl = __ballot_sync(0xffffffffu, flag == 2);
      const unsigned m_none = __ballot_sync(0xffffffffu, flag == 0);
      const int first = m_incl ? __ffs(m_incl) - 1 : 32;          // nearest inclusive prefix
      const unsigned need = first >= 31 ? 0xffffffffu : ((2u << first) - 1u);
      if (m_none & need) continue;                                // a predecessor has not published yet
      long long v = (lane <= first) ? (long long)(sw & ACC_VALUE_MASK) : 0ll;
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      excl += v;
      if (m_incl) break;
      i -= 32;
    }
    if (lane == 0) {
      if (chunk > 0) acc_st(status + chunk, ACC_FLAG_INCL | (unsigned long long)(excl + tot));
      prefix_s = excl;
      if (chunk == nchunks - 1) {
        const int64_t total = excl + tot;
        counts[0] = total;
        counts[1] = total < capacity ? total : capacity;
      }
    }
  }
  __syncthreads();
  int woff = 0;
  for (int i = 0; i < warp; ++i) woff += wsum[i];
  const int64_t idx0 = prefix_s + woff + incl - c;  // first record of this thread
  unsigned todo = __ballot_sync(0xffffffffu, c > 0);
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const unsigned bits = __shfl_sync(0xffffffffu, accbits, src);
    int64_t idx = __shfl_sync(0xffffffffu, idx0, src);
    const int64_t row0 = chunk * ACC_CHUNK + (warp * 32 + src) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (!((bits >> j) & 1u)) continue;
      if (idx < capacity) {
        uint32_t* dst = rows + (write_offset + idx) * F.row_words;
        const float* xr = xp + (row0 + j) * F.D;
        for (int w = lane; w < F.row_words; w += 32) {
          const int sc = src_s[w];
          uint32_t word = tmpl_s[w];
          if (sc >= 0) {
            // same float64 arithmetic as the bounds check of the draw kernel
            const int d = sc >> 1;
            // (x64: physical x already formed by reparam_tail_kernel for a non-affine rescaling)
            const unsigned long long b = __double_as_longlong(
                x64 ? x64[(row0 + j) * F.D + d] : (double)xr[d] * scale_s[d] + shift_s[d]);
            word = (sc & 1) ? (uint32_t)(b >> 32) : (uint32_t)b;
          } else if (sc < -1) {
            const unsigned long long b = __double_as_longlong(logl[row0 + j]);
            word = sc == -3 ? (uint32_t)(b >> 32) : (uint32_t)b;
          }
          dst[w] = word;
        }
      }
      ++idx;
    }
  }
}
