// tcgen05 relative attention -- placeholder until the tensor-core kernel lands.
#include "ops.cuh"
namespace mt {
bool rga_tc_supported(const RgaArgs&, int, int, bool) { return false; }
int rga_fwd_tc(const RgaArgs&, int, int, cudaStream_t) { set_error("rga_fwd_tc: not built"); return MT_E_UNSUPPORTED; }
int rga_bwd_tc(const RgaArgs&, int, int, cudaStream_t) { set_error("rga_bwd_tc: not built"); return MT_E_UNSUPPORTED; }
}  // namespace mt
