// tcgen05 (5th-gen tensor core) 3xTF32 GEMM -- placeholder until the kernel lands: the dispatcher
// falls through to the FFMA kernel while gemm_tc_supported() returns false.
#include "kgb_common.cuh"

namespace kgb {
bool gemm_tc_supported(int, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t) { return false; }
size_t gemm_tc_workspace_bytes(int, int64_t, int64_t, int64_t) { return 0; }
int gemm_tc(int, const float*, int64_t, const float*, int64_t, float*, int64_t, int64_t, int64_t, int64_t, float, float,
            const float*, int, void*, size_t, cudaStream_t) {
  set_error("gemm_tc: not built");
  return KGB_ERR_UNSUPPORTED;
}
}  // namespace kgb
