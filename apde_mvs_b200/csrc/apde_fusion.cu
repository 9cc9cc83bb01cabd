// apde_fusion.cu -- placeholder until the fusion kernels land (filled in next commit)
#include "apde_fusion.h"
namespace apde {
cudaError_t fusion_weak_vis_filter(const std::vector<FusionView> &, int, int, int, int, uint8_t *, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t fusion_run(const std::vector<FusionView> &, int, int, const uint8_t *, float *, float *, int64_t, int64_t *, cudaStream_t, uint64_t *) { return cudaErrorNotSupported; }
}
