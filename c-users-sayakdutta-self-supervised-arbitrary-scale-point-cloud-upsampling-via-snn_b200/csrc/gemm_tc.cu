// tcgen05 tensor-core GEMM engine -- placeholder until the engine lands: reports "unsupported" so that
// SAPCU_MODE_TC transparently uses the SIMT engine for every contraction.
#include "gemm_tc.h"

namespace sapcu {
bool gemm_tc_supported(const GemmArgs&, int) { return false; }
int launch_gemm_tc(const GemmArgs&, int, cudaStream_t) {
  set_error("gemm_tc: not built");
  return -1;
}
}  // namespace sapcu
