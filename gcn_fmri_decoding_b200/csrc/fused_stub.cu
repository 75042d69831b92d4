// placeholder until fused_fwd.cu / fused_bwd.cu land
#include "common.cuh"
namespace gcnb {
bool fused_fwd_supported(const LayerShape&) { return false; }
bool fused_bwd_supported(const LayerShape&, bool) { return false; }
size_t fused_cheb_workspace(const LayerShape&, bool, bool) { return 0; }
int fused_cheb_fwd(const float*, const int32_t*, int, const gcnb_csr&, const float*, const float*, float*, uint8_t*,
                   const LayerShape&, int, int, Workspace&, cudaStream_t) { set_error("fused path not built"); return GCNB_ERR_INVALID; }
int fused_cheb_bwd(const float*, const int32_t*, int, const float*, const uint8_t*, const float*, const gcnb_csr&, const gcnb_csr*, const float*,
                   float*, float*, float*, const LayerShape&, int, int, Workspace&, cudaStream_t) { set_error("fused path not built"); return GCNB_ERR_INVALID; }
}
