// TEST HARNESS ONLY: accept_fused_kernel (nessai_b200/csrc/accept.cuh: rejection step, single-pass
// look-back scan, warp-cooperative record writes -- the CUDA source, unchanged apart from the two
// NB200_SIMT_SHIM spots) under the CPU SIMT shim.  Blocks run one after the other in ticket order,
// so the look-back always finds its predecessors published.  Compile with
// -I tests/_hostcheck/fake_cuda.
#define NB200_SIMT_SHIM 1
#include "simt_shim.h"

#include "../../nessai_b200/csrc/accept.cuh"

// mirrors populate_accept_impl of nessai_b200.cu (record format, scratch zeroing, launch shape)
extern "C" int simt_populate_accept(int64_t n, int D, const float* xp, const double* x64, const double* scale,
                                    const double* shift, const double* logw, const double* logl, const double* d_max,
                                    uint64_t seed, uint64_t row_offset, double log_p_value, const uint8_t* row_template,
                                    int row_bytes, const int32_t* field_offsets, int logl_offset, uint8_t* rows,
                                    int64_t capacity, int64_t write_offset, int64_t* counts, int64_t* scratch) {
  if (n <= 0) return 0;
  RowFormat F;
  F.row_words = row_bytes / 4;
  F.D = D;
  for (int d = 0; d < D; ++d) F.off[d] = field_offsets[d];
  F.logp_off = field_offsets[D];
  const int64_t nchunks = (n + ACC_CHUNK - 1) / ACC_CHUNK;
  const size_t smem = (size_t)D * 16 + (size_t)F.row_words * 6 + 16;
  if (smem > sizeof(simt::dynamic_smem)) return 1;
  std::memset(scratch, 0, (size_t)(nchunks + 1) * sizeof(int64_t));
  simt_launch(accept_fused_kernel, (unsigned)nchunks, ACC_THREADS, xp, x64, scale, shift, logw, logl,
              logl ? logl_offset : -1, d_max, n, seed, row_offset, reinterpret_cast<unsigned long long*>(scratch),
              nchunks, log_p_value, reinterpret_cast<const uint32_t*>(row_template), F,
              reinterpret_cast<uint32_t*>(rows), capacity, write_offset, counts);
  return 0;
}

extern "C" double nb200_host_erfcinv(double) { return 0.0; }  // declared by the shim; unused here
