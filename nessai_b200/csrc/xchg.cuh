// Small all-gathers between the GPUs of a node over peer memory (NVLink), without NCCL and
// without the host: what the ranks of a sharded populate turn exchange is a handful of bytes --
// the shard's max log-weight before the rejection step (flowproposal.py:492 needs the maximum
// over the WHOLE turn) and every rank's {accepted, written} counts after it -- so a collective's
// launch and protocol latency (tens of microseconds each, on the critical path of a 300 us turn)
// is all cost.  Here every rank owns a small buffer of slots that its peers write directly
// (st.release.sys through the CUDA-IPC mapping) and that it polls locally (ld.acquire.sys):
// one warp, lane r talks to rank r.
//
//   slot(kind, parity, src) = { seq, payload[3] }     kind: 0 max, 1 counts; parity = seq & 1
//
// A rank cannot run two exchanges ahead of a peer (its next exchange needs the peer's), so two
// parities per kind are enough.  Spins are bounded: on a timeout *err is set and the caller raises.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace nb200 {

constexpr int XCHG_MAXW = 32;       // ranks (one lane each)
constexpr int XCHG_KINDS = 2;
constexpr int XCHG_WORDS = 3;       // payload words per slot
struct XchgSlot {
  unsigned long long seq;
  unsigned long long payload[XCHG_WORDS];
};
constexpr size_t XCHG_BYTES = sizeof(XchgSlot) * XCHG_KINDS * 2 * XCHG_MAXW;

__device__ __forceinline__ XchgSlot* xchg_slot(void* base, int kind, unsigned long long seq, int src) {
  return reinterpret_cast<XchgSlot*>(base) + ((kind * 2 + (int)(seq & 1ull)) * XCHG_MAXW + src);
}

// d_peers[r]: rank r's slot buffer as mapped into this process (d_peers[rank]: our own).
// d_src: n_words 64-bit words to publish.  d_gathered ([world][n_words], may be NULL): what every
// rank published, rank-major.  d_max_out (may be NULL): max over the ranks of word 0 read as a double.
__global__ void __launch_bounds__(32)
xchg_allgather_kernel(void* const* __restrict__ d_peers, int world, int rank, int kind, unsigned long long seq,
                      const unsigned long long* __restrict__ d_src, int n_words,
                      unsigned long long* __restrict__ d_gathered, double* __restrict__ d_max_out, int* d_err) {
  const int r = threadIdx.x;
  unsigned long long w[XCHG_WORDS] = {0ull, 0ull, 0ull};
  bool ok = true;
  if (r < world) {
    // publish: our payload into rank r's buffer, then the sequence number (release: after the payload)
    XchgSlot* dst = xchg_slot(d_peers[r], kind, seq, rank);
    for (int k = 0; k < n_words; ++k)
      asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(&dst->payload[k]), "l"(d_src[k]) : "memory");
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&dst->seq), "l"(seq) : "memory");
    // collect: rank r's payload from our own buffer
    XchgSlot* mine = xchg_slot(d_peers[rank], kind, seq, r);
    unsigned long long seen = 0ull;
    ok = false;
    for (long long spin = 0; spin < 20000000ll; ++spin) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(&mine->seq) : "memory");
      if (seen == seq) {
        ok = true;
        break;
      }
      if (spin > 64) __nanosleep(100);
    }
    for (int k = 0; k < n_words; ++k) {
      asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w[k]) : "l"(&mine->payload[k]) : "memory");
      if (d_gathered) d_gathered[r * n_words + k] = w[k];
    }
  }
  if (!ok) atomicExch(d_err, 1);
  if (d_max_out) {
    double v = r < world ? __longlong_as_double((long long)w[0]) : -INFINITY;
    // fmax would drop a NaN: a shard without a valid row publishes -inf, never NaN
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (r == 0) d_max_out[0] = v;
  }
}

}  // namespace nb200
