// tcgen05 specialisation of the folded flow (stub until the kernel lands).
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>
#include "flow_program.h"

struct PopulateArgs;

namespace nb200 {
struct TcProgram {
  bool valid = false;
};
inline void tc_free(TcProgram&) {}
inline int tc_build(TcProgram& t, const FlowOp*, int, const float*, int, int, int, int) {
  t.valid = false;
  return 0;
}
inline bool tc_enabled() { return false; }
inline int tc_launch_apply(TcProgram&, const float*, float*, float*, float*, int64_t, int, int, cudaStream_t) { return 1; }
template <typename A>
inline int tc_launch_populate(TcProgram&, const A&, int, cudaStream_t) { return 1; }
}  // namespace nb200
