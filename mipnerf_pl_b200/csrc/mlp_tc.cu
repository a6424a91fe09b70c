// placeholder until the tcgen05 kernels land: reports "unsupported" for every shape.
#include "mlp_tc.h"
namespace mipnerf {
bool tc_supported(const mipnerf_b200_config*, int) { return false; }
bool tc_mlp_supported(const mipnerf_b200_config*, int, int) { return false; }
size_t tc_packed_bytes(const mipnerf_b200_config*, int) { return 0; }
size_t tc_workspace_bytes(const mipnerf_b200_config*, int64_t, int) { return 0; }
cudaError_t tc_pack_weights(const mipnerf_b200_config*, const mipnerf_b200_weights*, int, void*, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t tc_forward(const mipnerf_b200_config*, const mipnerf_b200_weights*, const mipnerf_b200_rays*, int, const float*, const float*, int, int, mipnerf_b200_level_out*, void*, size_t, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t tc_mlp_forward(const mipnerf_b200_config*, const mipnerf_b200_weights*, const float*, const float*, int64_t, int, float*, float*, cudaStream_t) { return cudaErrorNotSupported; }
}
