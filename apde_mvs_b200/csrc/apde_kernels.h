// apde_kernels.h -- launch interface between the host driver (apde_api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>

#include <atomic>

#include "apde_device.cuh"

namespace apde {

// High-water mark, PER DEVICE, of the dynamic shared memory a kernel has been opted in to (cudaFuncSetAttribute acts on the
// current device's instance of the function, and the contexts of a multi-GPU job -- one host thread per GPU -- share the process).
struct SmemOptIn {
    std::atomic<size_t> mark[64];
    SmemOptIn() { for (auto &m : mark) m.store(0); }
    // true: the caller must (re)configure its kernels for `smem` on the current device, then call done(smem)
    bool needed(size_t smem, int *dev) const {
        cudaGetDevice(dev);
        return smem > mark[*dev & 63].load(std::memory_order_acquire);
    }
    void done(size_t smem, int dev) { mark[dev & 63].store(smem, std::memory_order_release); }
};


// run one kernel of the pass (stage ids of include/apde.h).  curve: optional [P][61] export of DepthToWeak.
cudaError_t launch_stage(const PassK &K, int stage, int iter, int color, cudaStream_t st, float *curve);
// weak-texture (APD) stages: nearest strong, anchors, RANSAC fit, deformable propagation
cudaError_t launch_stage_apd(const PassK &K, int stage, int iter, int color, cudaStream_t st);
cudaError_t launch_eval_costs(const PassK &K, int n, const int *tuples, const float4 *planes, int mode, float *out,
                              cudaStream_t st);
int prop_block_threads(int N);
// compacted (colour, strong | weak) pixel lists for the checkerboard kernels: lists = 4 x cap ints, counts = 4 ints
// scratch of the tile-ordered list build (apde_lists.cu): per-tile class counts, their scan
struct ListScratch {
    int *tile_counts = nullptr, *offsets = nullptr;
    void *scan_tmp = nullptr;
    size_t scan_bytes = 0, cap = 0;
    void release();
};
cudaError_t launch_build_lists(const PassK &K, int *lists, int *counts, int cap, ListScratch &ls, cudaStream_t st);

// red/black propagation as a pipeline of balanced column kernels (apde_prop.cu)
struct PropK {  // device view of the workspace, passed by value
    const int *list;      // compacted pixel list of the (colour, class) being swept
    const int *count;
    int cap;              // row stride of every per-pixel array below
    int *cand_pos;        // [8][cap]
    uint32_t *cand_flags; // [cap]  bits 0-7: candidate usable, bits 8-15: anchor present (weak)
    float *cost1;         // [9][N][cap] phase-1 costs (8 candidates + current plane)
    float4 *plane_now;    // [cap] state after the candidate choice
    float *depth_now, *cost_now, *cost_written, *wnorm;
    uint32_t *wmask;
    uint8_t *nh;          // [cap] number of refinement hypotheses (0 / 5 / 11)
    float4 *hyp;          // [11][cap]
    float *refpatch;      // [38][cap] reference patch texels + mean + variance
    float *anchorref;     // [88][cap] weak pixels: reference side of the 8 anchor patches (72 texels, 8 means, 8 variances)
    int *anchor_xy;       // [8][cap]  anchor coordinates packed (x & 0xffff) | (y << 16); x == -1: empty slot
    uint8_t *base3;       // [cap] weak pixels: first refinement slot after the fit-plane test (1 or 6)
    int *flags3, *colidx3, *colmap3;  // (pixel, selected view) columns: flags [N][cap], their prefix sum, column -> flat
    float *cost3;         // [11][N * cap]
    // weak pixels, anchor-sorted pipeline: the (pixel slot, anchor slot) pairs of the half-sweep ordered by the 8x8 image tile
    // their anchor lies in, so that a warp gathers anchor patches that are neighbours in the image
    unsigned *akey_in, *akey;  // [8 * cap] tile number of the anchor (0xffffffff: empty slot), unsorted / sorted
    int *aitem_in, *aitem;     // [8 * cap] pix * 8 + k, unsorted / sorted
    float *acost1;             // [9][8][N][cap]   anchor costs of the phase-1 hypotheses (< 0: the anchor does not take part)
    float *acost3;             // [5][8][N * cap]  anchor costs of the refinement hypotheses, by column
};
struct PropWorkspace {
    int cap = 0, views = 0;
    int *cand_pos = nullptr, *flags3 = nullptr, *colidx3 = nullptr, *colmap3 = nullptr;
    uint32_t *cand_flags = nullptr, *wmask = nullptr;
    float *cost1 = nullptr, *depth_now = nullptr, *cost_now = nullptr, *cost_written = nullptr, *wnorm = nullptr, *cost3 = nullptr;
    float4 *plane_now = nullptr, *hyp = nullptr;
    float *refpatch = nullptr, *anchorref = nullptr;
    int *anchor_xy = nullptr;
    uint8_t *nh = nullptr, *base3 = nullptr;
    void *scan_tmp = nullptr;
    size_t scan_bytes = 0;
    // anchor-sorted weak pipeline (allocated on first use, sized by the weak lists actually seen)
    unsigned *akey_in = nullptr, *akey = nullptr;
    int *aitem_in = nullptr, *aitem = nullptr;
    float *acost1 = nullptr, *acost3 = nullptr;
    void *sort_tmp = nullptr;
    size_t sort_bytes = 0, aitem_cap = 0, acost_cap = 0;
    cudaError_t reserve(int cap, int N);
    cudaError_t reserve_anchor_pipeline(size_t stride, int N);
    void release();
};
cudaError_t prop_half_sweep(const PassK &K, PropWorkspace &ws, bool weak, const int *list, const int *count, int max_pixels, int iter,
                            cudaStream_t st, uint64_t *launches);

// DepthToWeak + LocalRefine as a balanced (pixel, view) column workload (apde_sweep.cu)
struct SweepWorkspace {
    int *flags = nullptr, *colidx = nullptr;  // [N][P] view-major column flags and their exclusive prefix sum
    int *colmap = nullptr;                     // [ncols] column -> flat (view, pixel)
    size_t map_cap = 0;
    void *scan_tmp = nullptr;
    size_t scan_bytes = 0, flag_cap = 0, col_cap = 0, geo_cap = 0;
    float *ncc = nullptr;                      // [62][ncols] column costs (slot 61 = the current depth); geometric passes: FFMA(geom_factor, geom, ncc)
    float *geo = nullptr;                      // geometric passes: [22][ncols] = raw NCC [11] and raw geometric cost [11] of the LocalRefine steps
    int ncols = 0;
    int y0 = 0, rows = 0, Pb = 0;              // the band of rows the stored columns cover; Pb = slots (8x4 tiles, padded)
    size_t budget_mb = 0;                      // column storage budget (0 = APDE_SWEEP_BUDGET_MB or 48 GB)
    bool valid = false;                        // columns cover the whole image and match the problem state (set by DepthToWeak, used by LocalRefine)
    cudaError_t reserve_flags(size_t n);
    cudaError_t reserve_columns(size_t ncols, bool geom);
    void release();
};
cudaError_t sweep_depth_to_weak(const PassK &K, SweepWorkspace &ws, float *curve, cudaStream_t st, uint64_t *launches);
cudaError_t sweep_local_refine(const PassK &K, SweepWorkspace &ws, cudaStream_t st, uint64_t *launches);

// scene / map kernels (apde_maps.cu)
// OpenCV INTER_LINEAR resize of a u8 image to float (APD.cpp:574): dst[h][w] from src[H][W]
cudaError_t launch_resize_linear_u8(const uint8_t *src, int W, int H, float *dst, int w, int h, cudaStream_t st);
// OpenCV INTER_NEAREST resize (APD.cpp:607, 621, 625, 671-672); elem = bytes per element (1, 4 or 12)
cudaError_t launch_resize_nearest(const void *src, int W, int H, void *dst, int w, int h, int elem, cudaStream_t st);
// planes <- (normal, depth) maps (APD.cpp:674-682); weak/conf defaults (APD.cpp:656-657)
cudaError_t launch_planes_from_maps(const float *depth, const float *normal, float4 *planes, int P, cudaStream_t st);
cudaError_t launch_fill_u8(uint8_t *dst, uint8_t v, size_t n, cudaStream_t st);
cudaError_t launch_float_to_half(const float *src, void *dst, size_t n, cudaStream_t st);
cudaError_t launch_float_to_u16x4(const float *src, void *dst, size_t n, cudaStream_t st);
// ProcessProblem tail (main.cpp:168-178): depth range check, normals, states
cudaError_t launch_finish(const float4 *planes, uint8_t *weak, float *depth, float *normal, uint8_t *weak_out, int P,
                          float dmin, float dmax, cudaStream_t st);

// roofline denominators measured in-process (apde_microbench.cu)
cudaError_t microbench_fp32(double *tflops, cudaStream_t st);
cudaError_t microbench_tex(cudaTextureObject_t tex, int layer, int W, int H, double *gsamples, cudaStream_t st);
cudaError_t microbench_tex_pattern(cudaTextureObject_t tex, int layer, int W, int H, int mode, float spread, double *gsamples,
                                   cudaStream_t st);

}  // namespace apde
