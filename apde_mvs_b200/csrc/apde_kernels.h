// apde_kernels.h -- launch interface between the host driver (apde_api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>

#include "apde_device.cuh"

namespace apde {

// run one kernel of the pass (stage ids of include/apde.h).  curve: optional [P][61] export of DepthToWeak.
cudaError_t launch_stage(const PassK &K, int stage, int iter, int color, cudaStream_t st, float *curve);
// quad-cooperative versions of the heavy stages (apde_quad.cu); *handled = false when the stage has none
cudaError_t launch_stage_quad(const PassK &K, int stage, int iter, int color, cudaStream_t st, float *curve, bool *handled);
// weak-texture (APD) stages: nearest strong, anchors, RANSAC fit, deformable propagation
cudaError_t launch_stage_apd(const PassK &K, int stage, int iter, int color, cudaStream_t st);
cudaError_t launch_eval_costs(const PassK &K, int n, const int *tuples, const float4 *planes, int mode, float *out,
                              cudaStream_t st);
int prop_block_threads(int N);
// compacted (colour, strong | weak) pixel lists for the checkerboard kernels: lists = 4 x cap ints, counts = 4 ints
cudaError_t launch_build_lists(const PassK &K, int *lists, int *counts, int cap, cudaStream_t st);

// scene / map kernels (apde_maps.cu)
// OpenCV INTER_LINEAR resize of a u8 image to float (APD.cpp:574): dst[h][w] from src[H][W]
cudaError_t launch_resize_linear_u8(const uint8_t *src, int W, int H, float *dst, int w, int h, cudaStream_t st);
// OpenCV INTER_NEAREST resize (APD.cpp:607, 621, 625, 671-672); elem = bytes per element (1, 4 or 12)
cudaError_t launch_resize_nearest(const void *src, int W, int H, void *dst, int w, int h, int elem, cudaStream_t st);
// planes <- (normal, depth) maps (APD.cpp:674-682); weak/conf defaults (APD.cpp:656-657)
cudaError_t launch_planes_from_maps(const float *depth, const float *normal, float4 *planes, int P, cudaStream_t st);
cudaError_t launch_fill_u8(uint8_t *dst, uint8_t v, size_t n, cudaStream_t st);
// ProcessProblem tail (main.cpp:168-178): depth range check, normals, states
cudaError_t launch_finish(const float4 *planes, uint8_t *weak, float *depth, float *normal, uint8_t *weak_out, int P,
                          float dmin, float dmax, cudaStream_t st);

// roofline denominators measured in-process (apde_microbench.cu)
cudaError_t microbench_fp32(double *tflops, cudaStream_t st);
cudaError_t microbench_tex(cudaTextureObject_t tex, int layer, int W, int H, double *gsamples, cudaStream_t st);
cudaError_t microbench_tex_pattern(cudaTextureObject_t tex, int layer, int W, int H, int mode, float spread, double *gsamples,
                                   cudaStream_t st);

}  // namespace apde
