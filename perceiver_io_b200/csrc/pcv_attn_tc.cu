// placeholder until the tcgen05 kernel lands
#include "pcv_common.cuh"
namespace pcv {
bool attn_tc_supported(const pcv_attn_params& p, const char** why) { *why = "tcgen05 kernel not built yet"; return false; }
int launch_attn_tc(const pcv_attn_params& p, cudaStream_t stream) { set_error("tcgen05 kernel not built yet"); return PCV_ERR_UNSUPPORTED; }
int attn_tc_workspace_bytes(const pcv_attn_params& p, size_t* bytes) { *bytes = 0; return PCV_OK; }
}
