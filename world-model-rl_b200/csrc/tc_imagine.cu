// placeholder - replaced by the tcgen05 kernel
#include "common.cuh"
namespace wmrl {
bool tc_supported(const Dims&) { return false; }
size_t tc_packed_bytes(const Dims&) { return 0; }
size_t tc_workspace_bytes(const Dims&) { return 0; }
int tc_pack_weights(const Dims&, const wmrl_rssm_params*, const wmrl_actor_params*, void*, cudaStream_t) { set_error("tc path not built"); return 9; }
int tc_imagine_fwd(const Dims&, const wmrl_rssm_params*, const wmrl_actor_params*, const void*, const float*, const float*, const float*, float*, float*, float*, float*, float*, int32_t*, void*, cudaStream_t) { set_error("tc path not built"); return 9; }
}
