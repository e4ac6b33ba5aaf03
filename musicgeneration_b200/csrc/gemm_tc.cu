// tcgen05 GEMM -- placeholder until the tensor-core kernel lands (next commit).
#include "ops.cuh"
namespace mt {
bool gemm_tc_supported(int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int, int, int, int, int,
                       const void*, const void*, const void*) { return false; }
size_t gemm_tc_workspace_bytes(int64_t, int64_t, int64_t) { return 0; }
int gemm_tc(const void*, const void*, void*, const float*, const float*, const void*, int64_t, int64_t,
            int64_t, int64_t, int64_t, int64_t, int, int, int, int, int, void*, size_t, cudaStream_t) {
  set_error("gemm_tc: not built");
  return MT_E_UNSUPPORTED;
}
}  // namespace mt
