// msb_niw_tc.cuh -- tensor-core (tcgen05) path for the NIW Mahalanobis term.
// Placeholder until the tcgen05 kernel lands: reports "not done" so the caller
// runs the CUDA-core kernel.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include "../../include/mscope_b200.h"

namespace msb {
static inline int niw_tc_init(size_t, std::string &) { return MSB_OK; }
static inline size_t niw_tc_operand_bytes(size_t, uint32_t) { return 16; }
static inline int niw_tc_score(cudaStream_t, uint64_t *, const float *, uint32_t, const float *, const float *,
                               const float *, float *, size_t, float *, size_t, size_t, size_t, int, bool *done,
                               std::string &) {
  *done = false;
  return MSB_OK;
}
}  // namespace msb
