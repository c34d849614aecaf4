// TEST HARNESS ONLY: a minimal SIMT shim that lets g++ compile and RUN the CUDA kernels of
// nessai_b200/csrc/reparam_tail.cuh unchanged on the CPU.  One OS thread per CUDA thread; the
// blocks of a grid run one after the other, so a function-local `__shared__` variable becomes a
// `static` shared by the threads of the running block.  __syncthreads() is a 256-thread barrier,
// __shfl_xor_sync() an exchange through a per-warp slot array with two 32-thread barriers (every
// lane of a warp must call it, as in the kernels), atomics are a mutex.  Faithful for kernels
// whose warp-level primitives sit outside divergent code -- which is what is being checked.
#pragma once
#define __CUDACC__ 1
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __host__
#define __device__
#define __forceinline__ inline
#define __global__
#define __shared__ static
#define __launch_bounds__(...)
#define __grid_constant__

using std::isnan;
using std::max;  // CUDA's global integer min / max
using std::min;

struct simt_dim3 {
  unsigned x = 1, y = 1, z = 1;
};
static thread_local simt_dim3 threadIdx, blockIdx;
static simt_dim3 blockDim, gridDim;

namespace simt {
constexpr int kWarp = 32;
alignas(16) inline unsigned char dynamic_smem[232 * 1024];  // what a launch's dynamic shared memory points at
inline std::unique_ptr<std::barrier<>> block_barrier;
inline std::vector<std::unique_ptr<std::barrier<>>> warp_barrier;
inline std::vector<double> shfl_slot;
inline std::mutex atomic_mutex;
}  // namespace simt

inline void __syncthreads() { simt::block_barrier->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { simt::warp_barrier[threadIdx.x / simt::kWarp]->arrive_and_wait(); }

inline double __shfl_xor_sync(unsigned, double v, int lane_mask) {
  const unsigned tid = threadIdx.x, warp = tid / simt::kWarp;
  simt::shfl_slot[tid] = v;
  simt::warp_barrier[warp]->arrive_and_wait();
  const double r = simt::shfl_slot[tid ^ (unsigned)lane_mask];
  simt::warp_barrier[warp]->arrive_and_wait();
  return r;
}
inline float __shfl_xor_sync(unsigned m, float v, int lane_mask) {
  return (float)__shfl_xor_sync(m, (double)v, lane_mask);  // a float survives the round trip exactly
}

// generic warp exchange: every lane publishes a 64-bit word, then reads what it needs
namespace simt {
inline std::vector<unsigned long long> warp_word;
template <typename F>
inline auto warp_collective(unsigned long long mine, F read) {
  const unsigned tid = threadIdx.x, warp = tid / kWarp;
  warp_word[tid] = mine;
  warp_barrier[warp]->arrive_and_wait();
  auto r = read(&warp_word[warp * kWarp], tid % kWarp);
  warp_barrier[warp]->arrive_and_wait();
  return r;
}
}  // namespace simt

inline unsigned __ballot_sync(unsigned, int pred) {
  return simt::warp_collective(pred ? 1ull : 0ull, [](const unsigned long long* w, unsigned) {
    unsigned m = 0;
    for (int l = 0; l < simt::kWarp; ++l) m |= (unsigned)(w[l] & 1ull) << l;
    return m;
  });
}
template <typename T>
inline T __shfl_sync(unsigned, T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shim shuffles move at most 64 bits");
  unsigned long long bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  const unsigned long long got = simt::warp_collective(
      bits, [src_lane](const unsigned long long* w, unsigned) { return w[src_lane & (simt::kWarp - 1)]; });
  T r;
  std::memcpy(&r, &got, sizeof(T));
  return r;
}
template <typename T>
inline T __shfl_up_sync(unsigned, T v, unsigned delta) {
  unsigned long long bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  const unsigned long long got = simt::warp_collective(bits, [delta](const unsigned long long* w, unsigned lane) {
    return lane >= delta ? w[lane - delta] : w[lane];  // lanes below delta keep their own value
  });
  T r;
  std::memcpy(&r, &got, sizeof(T));
  return r;
}
inline long long __shfl_xor_sync(unsigned, long long v, int lane_mask) {
  return (long long)simt::warp_collective((unsigned long long)v, [lane_mask](const unsigned long long* w, unsigned lane) {
    return w[lane ^ (unsigned)lane_mask];
  });
}
inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
inline unsigned atomicAdd(unsigned* p, unsigned v) {
  std::lock_guard<std::mutex> g(simt::atomic_mutex);
  const unsigned old = *p;
  *p = old + v;
  return old;
}
#define __align__(n) alignas(n)

// cache-hinted loads / stores and the fast-math intrinsics of the device build
template <typename T>
inline T __ldg(const T* p) { return *p; }
template <typename T>
inline T __ldcs(const T* p) { return *p; }
template <typename T>
inline void __stcs(T* p, T v) { *p = v; }
inline float __fdividef(float a, float b) { return a / b; }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline float __expf(float x) { return expf(x); }
inline float __logf(float x) { return logf(x); }

inline double atomicAdd(double* p, double v) {
  std::lock_guard<std::mutex> g(simt::atomic_mutex);
  const double old = *p;
  *p = old + v;
  return old;
}

// CUDA's erfcinv for the device build of the header (supplied by the harness)
extern "C" double nb200_host_erfcinv(double);
inline double erfcinv(double y) { return nb200_host_erfcinv(y); }

inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long expected, unsigned long long v) {
  std::lock_guard<std::mutex> g(simt::atomic_mutex);
  const unsigned long long old = *p;
  if (old == expected) *p = v;
  return old;
}
inline long long __double_as_longlong(double v) {
  long long r;
  std::memcpy(&r, &v, sizeof r);
  return r;
}
inline double __longlong_as_double(long long v) {
  double r;
  std::memcpy(&r, &v, sizeof r);
  return r;
}
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
#define __cosf(x) cosf(x)  // (glibc's <math.h> declares these two names itself)
#define __sinf(x) sinf(x)
using std::isfinite;

#ifndef SIMT_HAVE_POPULATE_COMMON
namespace nb200 {
// populate_common.cuh's atomic maximum, for kernels compiled without that header
inline void atomic_max_double(double* addr, double v) {
  std::lock_guard<std::mutex> g(simt::atomic_mutex);
  if (v > *addr) *addr = v;
}
}  // namespace nb200
#endif

// Run `kernel(args...)` over a 1-D grid of 1-D blocks.  `block` OS threads are created once per
// launch and walk the blocks of the grid together (a barrier between blocks keeps the block-shared
// statics of one block from leaking into the next).
namespace simt {
inline bool launch_failed = false;  // the machine cannot run `block` OS threads at once
// Can `block` threads exist at the same time here?  (A thread that cannot be created in the middle of
// a launch would leave the others waiting at a barrier for ever, so this is probed beforehand.)
inline bool can_run(unsigned block) {
  static unsigned proven = 0;
  if (block <= proven) return true;
  std::atomic<bool> go{false};
  std::vector<std::thread> probe;
  bool ok = true;
  try {
    probe.reserve(block);
    for (unsigned t = 0; t < block; ++t)
      probe.emplace_back([&go]() {
        while (!go.load()) std::this_thread::yield();
      });
  } catch (...) {
    ok = false;
  }
  go.store(true);
  for (auto& th : probe) th.join();
  if (ok) proven = block;
  return ok;
}
}  // namespace simt

template <typename K, typename... A>
void simt_launch(K kernel, unsigned grid, unsigned block, A... args) {
  if (simt::launch_failed || !simt::can_run(block)) {
    simt::launch_failed = true;
    return;
  }
  blockDim.x = block;
  gridDim.x = grid;
  simt::shfl_slot.assign(block, 0.0);
  simt::warp_word.assign((block + simt::kWarp - 1) / simt::kWarp * simt::kWarp, 0ull);
  simt::block_barrier = std::make_unique<std::barrier<>>(block);
  simt::warp_barrier.clear();
  for (unsigned w = 0; w < (block + simt::kWarp - 1) / simt::kWarp; ++w) {
    const unsigned lanes = std::min<unsigned>(simt::kWarp, block - w * simt::kWarp);
    simt::warp_barrier.push_back(std::make_unique<std::barrier<>>(lanes));
  }
  std::barrier<> next_block(block);
  std::vector<std::thread> threads;
  threads.reserve(block);
  for (unsigned t = 0; t < block; ++t)
    threads.emplace_back([&, t]() {
      for (unsigned b = 0; b < grid; ++b) {
        threadIdx.x = t;
        blockIdx.x = b;
        kernel(args...);
        next_block.arrive_and_wait();
      }
    });
  for (auto& th : threads) th.join();
}

// Harness entry points report this when a launch could not run (tests skip on it).
extern "C" int simt_launch_failed() { return simt::launch_failed ? 1 : 0; }
// 0 when `block` OS threads can exist at once on this machine (fixtures skip otherwise).
extern "C" int simt_probe(unsigned block) { return simt::can_run(block) ? 0 : 1; }
