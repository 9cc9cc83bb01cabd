// apde_context.h -- the context behind the C ABI of include/apde.h, shared by the host-side translation units
// (apde_api.cu: scene / problem / schedule; apde_comm.cu: multi-GPU exchange).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/apde.h"
#include "apde_fusion.h"
#include "apde_kernels.h"

using namespace apde;  // internal header of the host-side translation units

static const int kStages = 11;
static const int kCounterWords = (kStages + 1) * 4;
std::string &apde_error_slot();  // thread-local last error (apde_last_error)
static inline int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    apde_error_slot() = buf;
    return code;
}
#ifndef CU
#define CU(call)                                                                                           \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) return fail(APDE_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)
#endif

struct ViewStore {
    apde_camera cam;
    std::vector<int> src;
    uint8_t *d_gray = nullptr, *d_bgr = nullptr;
    uint8_t *d_sa = nullptr;  // segment labels as read from sa_masks/<id>.bin (own size saw x sah), nullptr = none
    int saw = 0, sah = 0;
    int mw = 0, mh = 0;  // size of the stored maps (0 = none yet)
    int dw = 0, dh = 0;  // size of this view's depth map in the READABLE pool d_depth_pool[cur]: equal to mw x mh, except in
                         // Jacobi mode between a view's finish and the end of the pass (the new map went to the other pool)
    float *d_normal = nullptr;
    uint8_t *d_weak = nullptr, *d_conf = nullptr;
    bool has_conf = false;
};

struct apde_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int V = 0, W = 0, H = 0;
    bool committed = false;
    std::vector<ViewStore> views;
    float *d_depth_pool[2] = {nullptr, nullptr};  // [V][W*H] ; [1] only allocated in Jacobi mode
    // normal / weak / confidence maps of all views, contiguous per field so that a multi-GPU job can all-gather them in
    // place before fusion (apde_map_pool); ViewStore::d_normal / d_weak / d_conf point into these
    float *d_normal_pool = nullptr;               // [V][W*H*3]
    uint8_t *d_weak_pool = nullptr, *d_conf_pool = nullptr;  // [V][W*H]
    int cur = 0;
    // pyramid level (one alive at a time)
    // pyramid levels: built on first use, kept for the whole scene (images are immutable after commit)
    struct Level {
        int scale = 0, w = 0, h = 0;
        cudaArray_t arr = nullptr;
        cudaTextureObject_t tex = 0;
        float *lin = nullptr;  // [V][h*w]
        void *half = nullptr;  // [V][h*w] __half staging when the texture is stored as fp16
        bool fp16 = false;
        bool u8 = false;       // 8-bit UNORM texels (scale 1 only: the images are 8-bit integers)
        bool u16 = false;      // 16-bit UNORM texels holding 4 x value (power-of-two scales of divisible sizes: 2x2 means)
        bool stale = true;     // contents must be re-derived from the views' images
    };
    std::vector<Level> levels;
    int level_scale = 0, lw = 0, lh = 0;  // the level of the current problem
    float level_unorm = 0.0f, level_inv = 1.0f;
    cudaArray_t level_arr = nullptr;
    cudaTextureObject_t level_tex = 0;
    float *d_level_lin = nullptr;  // [V][lh*lw]
    // problem working set (allocated once at full resolution)
    bool ws_alloc = false;
    float4 *d_planes = nullptr, *d_fit = nullptr;
    uint8_t *d_sa = nullptr;  // segment labels of the active problem's reference view at the working size
    ListScratch list_scratch;
    float *d_costs = nullptr, *d_depthws = nullptr, *d_scratch_depth = nullptr, *d_scratch_normal = nullptr;
    uint32_t *d_sel = nullptr;
    uint4 *d_vw = nullptr;
    uint8_t *d_weak = nullptr, *d_conf = nullptr, *d_reliable = nullptr;
    short2 *d_nearest = nullptr, *d_anchors = nullptr;
    uint16_t *d_ns_tiles = nullptr;
    float *d_curve = nullptr;   // [P][61] DepthToWeak cost curves while capture_curve is on
    size_t curve_cap = 0;
    bool capture_curve = false;
    int *d_lists = nullptr, *d_list_counts = nullptr;  // compacted checkerboard pixel lists (4 x list_cap)
    int list_cap = 0;
    int h_list_counts[4] = {0, 0, 0, 0};  // host copy of the list lengths: sizes (or skips) the list-driven launches
    bool lists_dirty = true;
    unsigned long long *d_counters = nullptr;  // [kStages + 1][4]: per-stage NCC-Old / NCC-New / geom evaluation counts
    uint64_t launches = 0;
    // per-stage profiling (CUDA events on the launching stream)
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool, pm_events;
    std::vector<int> ev_stage;
    double stage_ms[16] = {0};
    uint64_t stage_launches[16] = {0};
    // current problem
    bool problem_active = false;
    int ref_view = -1;
    bool jacobi_write = false;
    apde_params params;
    std::vector<apde_camera> cams;
    PassK K;
    SweepWorkspace sweep;  // DepthToWeak / LocalRefine column costs
    PropWorkspace prop;    // propagation pipeline buffers
    // fusion state
    uint8_t *d_skip = nullptr;
    FusedPoints fused;  // cloud of the last count-only fusion call, until apde_fuse_take_points
    // multi-GPU job (apde_comm.cu): NCCL communicator of the job, this rank's block of views, the exchange stream
    struct Comm *comm = nullptr;
};

// view block of `rank` when V views are dealt out in contiguous blocks over `world` ranks (the first V % world ranks hold one more)
static inline void apde_view_block(int V, int world, int rank, int *first, int *count) {
    const int base = V / world, extra = V % world;
    *count = base + (rank < extra ? 1 : 0);
    *first = rank * base + (rank < extra ? rank : extra);
}

// apde_comm.cu: hooks the schedule calls when the context belongs to a multi-GPU job
int apde_comm_attached(const apde_context *c);
int apde_comm_block(const apde_context *c, int *first_view, int *num_views);
// queue the broadcast of view `view`'s freshly finished depth row (pool `which_pool` of the double buffer) from its owner;
// collective in program order; runs on the exchange stream after everything queued on c->stream so far
int apde_comm_share_depth_row(apde_context *c, int k /* k-th view of every rank's block */, int which_pool, int w, int h);
// make c->stream wait for the queued exchanges; adds the exposed wait to *exposed_ms when timing is on
int apde_comm_join(apde_context *c, double *exposed_ms, uint64_t *bytes);

