// apde_fusion.h -- GPU fusion: WeakVisFilter + RunFusion (APD.cpp:962-1227).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/apde.h"

namespace apde {

struct FusionView {
    apde_camera cam;       // rescaled to the map size (RescaleImageAndCamera, APD.cpp:844-864)
    const float *depth;    // [P]
    const float *normal;   // [P][3]
    const uint8_t *weak;   // [P]
    const uint8_t *conf;   // [P]
    const uint8_t *bgr;    // full-resolution colour image [Hfull][Wfull][3] or nullptr
    std::vector<int> src;  // neighbour view indices
};

// skip: [V][P] device, rows [first_view, first_view + num_views) zero-initialised by the caller and filled here
cudaError_t fusion_weak_vis_filter(const std::vector<FusionView> &views, int w, int h, int first_view, int num_views, uint8_t *skip,
                                   cudaStream_t st);
// host copy of a fused cloud kept by the context between a count-only apde_fuse_variant call and apde_fuse_take_points
struct FusedPoints {
    std::vector<float> xyz, bgr;
    bool valid = false;
};

// greedy fusion in the reference order; xyz / bgr are HOST buffers (may be null); keep (may be null) receives ALL points
cudaError_t fusion_run(const std::vector<FusionView> &views, int w, int h, const uint8_t *skip, float *xyz, float *bgr,
                       int64_t max_points, int64_t *num_points, cudaStream_t st, uint64_t *launches, FusedPoints *keep = nullptr);

// Tanks-and-Temples variants: 1 = RunFusion_TAT_I (APD.cpp:1229-1431), 2 = RunFusion_TAT_A (APD.cpp:1433-1608)
cudaError_t fusion_run_tat(const std::vector<FusionView> &views, int w, int h, int variant, const uint8_t *skip, float *xyz,
                           float *bgr, int64_t max_points, int64_t *num_points, cudaStream_t st, uint64_t *launches,
                           FusedPoints *keep = nullptr);

// frees the scratch blocks the fusion entry points keep for re-use
void fusion_release_cache();

}  // namespace apde
