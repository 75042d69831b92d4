// placeholder until fused_bwd.cu lands
#include "common.cuh"
namespace gcnb {
bool fused_bwd_supported(const LayerShape&, bool) { return false; }
size_t fused_cheb_workspace(const LayerShape&, bool, bool) { return 256; }
int fused_cheb_bwd(const float*, const int32_t*, int, const float*, const uint8_t*, const float*, const gcnb_csr&, const gcnb_csr*, const float*,
                   float*, float*, float*, const LayerShape&, int, int, Workspace&, cudaStream_t) { set_error("fused backward not built"); return GCNB_ERR_INVALID; }
}
