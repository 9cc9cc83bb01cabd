// apde_api.cu -- host driver behind the C ABI of include/apde.h: scene-resident images / cameras / maps, pyramid
// levels as one layered texture, per-problem set-up (the reference's InuputInitialization + CudaSpaceInitialization,
// APD.cpp:501-814, without any host round trip), the stage launcher and the multi-scale schedule of main.cpp:303-367.
#include <cuda_profiler_api.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>

#include "apde_context.h"
#include "apde_fusion.h"

std::string &apde_error_slot() {
    static thread_local std::string e;
    return e;
}

extern "C" {

const char *apde_last_error(void) { return apde_error_slot().c_str(); }
const char *apde_version(void) { return "apde-b200 0.1 sm_100a"; }

void apde_params_default(apde_params *p) {  // main.h:80-100
    memset(p, 0, sizeof(*p));
    p->max_iterations = 3;
    p->num_images = 5;
    p->top_k = 4;
    p->depth_min = 0.0f;
    p->depth_max = 1.0f;
    p->geom_consistency = 0;
    p->use_impetus = 1;
    p->strong_radius = 5;
    p->strong_increment = 2;
    p->weak_radius = 5;
    p->weak_increment = 5;
    p->use_APD = 1;
    p->use_sa = 1;
    p->weak_peak_radius = 2;
    p->rotate_time = 4;
    p->ransac_threshold = 0.005f;
    p->geom_factor = 0.2f;
    p->state = APDE_FIRST_INIT;
}

void apde_schedule_default(apde_schedule *s) {
    memset(s, 0, sizeof(*s));
    s->rounds = 0;
    s->geom_iterations = 3;
    s->jacobi = 0;
    s->use_impetus = 1;
    s->geom_factor = 0.2f;
    s->seed = 1;
    s->use_sa = 1;
}

int apde_create(int device, apde_context **out) {
    if (!out) return fail(APDE_ERR_ARG, "apde_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return fail(APDE_ERR_CUDA, "apde_create: no CUDA device (%s); libapde has no CPU fallback",
                    e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(APDE_ERR_ARG, "apde_create: device %d out of range (%d devices)", device, n);
    CU(cudaSetDevice(device));
    apde_context *c = new apde_context();
    c->device = device;
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&c->ev0));
    CU(cudaEventCreate(&c->ev1));
    CU(cudaMalloc(&c->d_counters, kCounterWords * sizeof(unsigned long long)));
    CU(cudaMemset(c->d_counters, 0, kCounterWords * sizeof(unsigned long long)));
    *out = c;
    return APDE_OK;
}

static void free_level(apde_context *c) {
    for (auto &L : c->levels) {
        if (L.tex) cudaDestroyTextureObject(L.tex);
        if (L.arr) cudaFreeArray(L.arr);
        if (L.lin) cudaFree(L.lin);
        if (L.half) cudaFree(L.half);
    }
    c->levels.clear();
    c->level_tex = 0; c->level_arr = nullptr; c->d_level_lin = nullptr; c->level_scale = 0;
}
static void free_scene(apde_context *c) {
    for (auto &v : c->views) {
        cudaFree(v.d_gray); cudaFree(v.d_bgr); cudaFree(v.d_sa);
    }
    c->views.clear();
    cudaFree(c->d_normal_pool); cudaFree(c->d_weak_pool); cudaFree(c->d_conf_pool);
    c->d_normal_pool = nullptr; c->d_weak_pool = c->d_conf_pool = nullptr;
    cudaFree(c->d_depth_pool[0]); cudaFree(c->d_depth_pool[1]);
    c->d_depth_pool[0] = c->d_depth_pool[1] = nullptr;
    free_level(c);
    if (c->ws_alloc) {
        cudaFree(c->d_planes); cudaFree(c->d_fit); cudaFree(c->d_costs); cudaFree(c->d_depthws);
        cudaFree(c->d_scratch_depth); cudaFree(c->d_scratch_normal); cudaFree(c->d_sel); cudaFree(c->d_vw);
        cudaFree(c->d_weak); cudaFree(c->d_conf); cudaFree(c->d_reliable); cudaFree(c->d_nearest); cudaFree(c->d_anchors); cudaFree(c->d_ns_tiles); cudaFree(c->d_curve); c->d_curve = nullptr; c->curve_cap = 0;
        cudaFree(c->d_lists); cudaFree(c->d_list_counts);
        c->ws_alloc = false;
    }
    cudaFree(c->d_skip);
    c->d_skip = nullptr;
    std::vector<float>().swap(c->fused.xyz);
    std::vector<float>().swap(c->fused.bgr);
    c->fused.valid = false;
    cudaFree(c->d_sa);
    c->d_sa = nullptr;
    c->list_scratch.release();
    fusion_release_cache();
    c->sweep.release();
    c->prop.release();
    c->committed = false;
    c->problem_active = false;
}

void apde_destroy(apde_context *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    apde_comm_destroy(c);
    free_scene(c);
    cudaFree(c->d_counters);
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    for (auto e : c->pm_events) cudaEventDestroy(e);
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
    cudaStreamDestroy(c->stream);
    delete c;
}

// ------------------------------------------------------------------------------------------------ scene
int apde_scene_begin(apde_context *c, int num_views, int width, int height) {
    if (!c) return fail(APDE_ERR_ARG, "null context");
    if (num_views <= 0 || width <= 0 || height <= 0) return fail(APDE_ERR_ARG, "scene_begin: bad dimensions");
    if (width > 32768 || height > 32768 || num_views > 2048)
        return fail(APDE_ERR_ARG, "scene_begin: %dx%d x %d views exceeds the layered-texture limits", width, height, num_views);
    CU(cudaSetDevice(c->device));
    if (c->committed && c->V == num_views && c->W == width && c->H == height) {
        // same shape as the resident scene: keep every allocation, only forget the maps and the pyramid level
        CU(cudaStreamSynchronize(c->stream));
        for (auto &v : c->views) {
            v.mw = v.mh = v.dw = v.dh = 0; v.has_conf = false; v.src.clear();
            CU(cudaMemsetAsync(v.d_conf, 0, (size_t)c->W * c->H, c->stream));
            // nothing of the previous scene may leak into this one: label maps and colour images are per scene
            cudaFree(v.d_sa); v.d_sa = nullptr; v.saw = v.sah = 0;
            cudaFree(v.d_bgr); v.d_bgr = nullptr;
        }
        for (auto &L : c->levels) L.stale = true;
        c->level_scale = 0;
        c->problem_active = false;
        c->committed = false;
        c->cur = 0;
        return APDE_OK;
    }
    free_scene(c);
    c->V = num_views; c->W = width; c->H = height;
    c->views.resize(num_views);
    const size_t P = (size_t)width * height;
    CU(cudaMalloc(&c->d_normal_pool, (size_t)num_views * P * 3 * sizeof(float)));
    CU(cudaMalloc(&c->d_weak_pool, (size_t)num_views * P));
    CU(cudaMalloc(&c->d_conf_pool, (size_t)num_views * P));
    // no confidence map exists before the first geometric / APD pass (main.cpp:187-190).  Everything is ordered on c->stream: it is
    // a non-blocking stream, which the legacy default stream does not synchronise with
    CU(cudaMemsetAsync(c->d_conf_pool, 0, (size_t)num_views * P, c->stream));
    for (int i = 0; i < num_views; ++i) {
        ViewStore &v = c->views[i];
        CU(cudaMalloc(&v.d_gray, P));
        v.d_normal = c->d_normal_pool + (size_t)i * P * 3;
        v.d_weak = c->d_weak_pool + (size_t)i * P;
        v.d_conf = c->d_conf_pool + (size_t)i * P;
        memset(&v.cam, 0, sizeof(v.cam));
    }
    CU(cudaMalloc(&c->d_depth_pool[0], (size_t)num_views * P * sizeof(float)));
    CU(cudaMemsetAsync(c->d_depth_pool[0], 0, (size_t)num_views * P * sizeof(float), c->stream));
    c->cur = 0;
    return APDE_OK;
}

int apde_scene_set_view(apde_context *c, int view, const uint8_t *gray, const uint8_t *bgr, const apde_camera *cam) {
    if (!c || view < 0 || view >= c->V || !gray || !cam) return fail(APDE_ERR_ARG, "scene_set_view: bad argument");
    CU(cudaSetDevice(c->device));
    ViewStore &v = c->views[view];
    const size_t P = (size_t)c->W * c->H;
    CU(cudaMemcpyAsync(v.d_gray, gray, P, cudaMemcpyHostToDevice, c->stream));
    if (bgr) {
        if (!v.d_bgr) CU(cudaMalloc(&v.d_bgr, P * 3));
        CU(cudaMemcpyAsync(v.d_bgr, bgr, P * 3, cudaMemcpyHostToDevice, c->stream));
    } else if (v.d_bgr) {  // a view without colours must not keep an earlier image's
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(v.d_bgr);
        v.d_bgr = nullptr;
    }
    // the pyramid levels derived from the previous image of this view are stale (images may be replaced on a committed scene)
    for (auto &L : c->levels) L.stale = true;
    c->level_scale = 0;
    v.cam = *cam;
    v.cam.width = c->W;
    v.cam.height = c->H;
    CU(cudaStreamSynchronize(c->stream));
    return APDE_OK;
}

int apde_view_set_sa_mask(apde_context *c, int view, const uint8_t *labels, int width, int height) {
    if (!c || view < 0 || view >= c->V) return fail(APDE_ERR_ARG, "view_set_sa_mask: bad argument");
    CU(cudaSetDevice(c->device));
    ViewStore &v = c->views[view];
    cudaFree(v.d_sa);
    v.d_sa = nullptr; v.saw = v.sah = 0;
    if (!labels) return APDE_OK;
    if (width < 1 || height < 1) return fail(APDE_ERR_ARG, "view_set_sa_mask: bad size %d x %d", width, height);
    CU(cudaMalloc(&v.d_sa, (size_t)width * height));
    CU(cudaMemcpyAsync(v.d_sa, labels, (size_t)width * height, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    v.saw = width; v.sah = height;
    return APDE_OK;
}

int apde_scene_set_pairs(apde_context *c, int view, int num_src, const int32_t *src_views) {
    if (!c || view < 0 || view >= c->V || num_src < 0 || (num_src > 0 && !src_views))
        return fail(APDE_ERR_ARG, "scene_set_pairs: bad argument");
    if (num_src > APDE_MAX_IMAGES - 1)
        return fail(APDE_ERR_ARG, "Can't process so much images: %d", num_src + 1);  // APD.cpp:528-531
    for (int i = 0; i < num_src; ++i)
        if (src_views[i] < 0 || src_views[i] >= c->V) return fail(APDE_ERR_ARG, "scene_set_pairs: source %d out of range", src_views[i]);
    c->views[view].src.assign(src_views, src_views + num_src);
    return APDE_OK;
}

static int alloc_workspace(apde_context *c) {
    if (c->ws_alloc) return APDE_OK;
    const size_t P = (size_t)c->W * c->H;
    CU(cudaMalloc(&c->d_planes, P * sizeof(float4)));
    CU(cudaMalloc(&c->d_fit, P * sizeof(float4)));
    CU(cudaMalloc(&c->d_costs, P * sizeof(float)));
    CU(cudaMalloc(&c->d_depthws, P * sizeof(float) * APDE_MAX_IMAGES));
    CU(cudaMalloc(&c->d_scratch_depth, P * sizeof(float)));
    CU(cudaMalloc(&c->d_scratch_normal, P * 3 * sizeof(float)));
    CU(cudaMalloc(&c->d_sel, P * sizeof(uint32_t)));
    CU(cudaMalloc(&c->d_vw, P * sizeof(uint4)));
    CU(cudaMalloc(&c->d_weak, P));
    CU(cudaMalloc(&c->d_conf, P));
    CU(cudaMalloc(&c->d_reliable, P));
    CU(cudaMalloc(&c->d_nearest, P * sizeof(short2)));
    CU(cudaMalloc(&c->d_ns_tiles, (size_t)((c->W + 7) / 8) * ((c->H + 7) / 8) * sizeof(uint16_t)));
    CU(cudaMalloc(&c->d_anchors, P * APDE_ANCHOR_NUM * sizeof(short2)));
    c->list_cap = (int)(P / 2 + 64);
    CU(cudaMalloc(&c->d_lists, (size_t)4 * c->list_cap * sizeof(int)));
    CU(cudaMalloc(&c->d_list_counts, 4 * sizeof(int)));
    CU(cudaMemsetAsync(c->d_vw, 0, P * sizeof(uint4), c->stream));
    CU(cudaMemsetAsync(c->d_sel, 0, P * sizeof(uint32_t), c->stream));
    c->ws_alloc = true;
    return APDE_OK;
}

int apde_scene_commit(apde_context *c) {
    if (!c || c->V <= 0) return fail(APDE_ERR_STATE, "scene_commit: no scene");
    CU(cudaSetDevice(c->device));
    int rc = alloc_workspace(c);
    if (rc) return rc;
    c->committed = true;
    return APDE_OK;
}

// build the pyramid level for scale_size: INTER_LINEAR resize of every view (APD.cpp:564-588) into one layered texture
static int ensure_level(apde_context *c, int scale) {
    if (c->level_scale == scale) return APDE_OK;
    apde_context::Level *L = nullptr;
    for (auto &l : c->levels) if (l.scale == scale) L = &l;
    if (!L) {
        const float factor = 1.0f / (float)scale;
        const int lw = (int)std::round(c->W * factor), lh = (int)std::round(c->H * factor);
        if (lw < 16 || lh < 16) return fail(APDE_ERR_ARG, "pyramid level %dx%d too small", lw, lh);
        apde_context::Level nl;
        nl.scale = scale; nl.w = lw; nl.h = lh;
        const size_t P = (size_t)lw * lh;
        CU(cudaMalloc(&nl.lin, (size_t)c->V * P * sizeof(float)));
        // EXPERIMENT, off by default (APDE_TEX_FP16=1): fp16 texels halve the bytes behind every gather (the binding unit is
        // the L1TEX data pipe): +36 % gather rate at +-4 px scatter, +9 % end to end.  The texels themselves are exact in
        // fp16 (8-bit images and their 2x2 means), but the texture unit then rounds the FILTERED sample to fp16 (measured:
        // max 0.0625 grey levels off, 10 % bit-identical) -> costs move by ~1e-3: outside the 1e-4 parity bar, so not used.
        static const bool want16 = [] { const char *e = getenv("APDE_TEX_FP16"); return e && e[0] == '1'; }();
        nl.fp16 = want16 && (scale == 1 || (scale == 2 && c->W % 2 == 0 && c->H % 2 == 0));
        // 8-bit UNORM texels at full resolution: exact (see fetch() in apde_device.cuh); APDE_TEX_U8=0 keeps float32
        static const bool want8 = [] { const char *e = getenv("APDE_TEX_U8"); return !(e && e[0] == '0'); }();
        nl.u8 = want8 && !nl.fp16 && scale == 1;
        // 16-bit UNORM texels (4 x value) for the coarser levels were measured NOT exact: the unit filters them at fp16
        // precision (max 0.125 grey levels off, tools/diag_u8.py), so coarse levels keep float32 texels.  APDE_TEX_U16=1 forces it.
        static const bool want16u = [] { const char *e = getenv("APDE_TEX_U16"); return e && e[0] == '1'; }();
        nl.u16 = want16u && !nl.fp16 && scale > 1 && (scale & (scale - 1)) == 0 && c->W % scale == 0 && c->H % scale == 0;
        cudaChannelFormatDesc desc = nl.u8 ? cudaCreateChannelDesc(8, 0, 0, 0, cudaChannelFormatKindUnsigned)
                                     : nl.u16 ? cudaCreateChannelDesc(16, 0, 0, 0, cudaChannelFormatKindUnsigned)
                                     : nl.fp16 ? cudaCreateChannelDescHalf() : cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
        if (nl.u16) CU(cudaMalloc(&nl.half, (size_t)c->V * P * 2));
        if (nl.fp16) CU(cudaMalloc(&nl.half, (size_t)c->V * P * 2));
        CU(cudaMalloc3DArray(&nl.arr, &desc, make_cudaExtent(lw, lh, c->V), cudaArrayLayered));
        cudaResourceDesc rd;
        memset(&rd, 0, sizeof(rd));
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = nl.arr;
        cudaTextureDesc td;
        memset(&td, 0, sizeof(td));
        // the reference asks for "wrap" with unnormalised coordinates, which CUDA serves as clamp (APD.cpp:701-705)
        td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModeLinear;
        td.readMode = (nl.u8 || nl.u16) ? cudaReadModeNormalizedFloat : cudaReadModeElementType;
        td.normalizedCoords = 0;
        CU(cudaCreateTextureObject(&nl.tex, &rd, &td, nullptr));
        c->levels.push_back(nl);
        L = &c->levels.back();
    }
    if (L->stale) {
        const size_t P = (size_t)L->w * L->h;
        for (int v = 0; v < c->V; ++v)
            CU(launch_resize_linear_u8(c->views[v].d_gray, c->W, c->H, L->lin + (size_t)v * P, L->w, L->h, c->stream));
        c->launches += c->V;
        cudaMemcpy3DParms cp;
        memset(&cp, 0, sizeof(cp));
        if (L->u8) {
            for (int v = 0; v < c->V; ++v) {  // the 8-bit images are the texels; one layer per view
                cudaMemcpy3DParms cv;
                memset(&cv, 0, sizeof(cv));
                cv.srcPtr = make_cudaPitchedPtr(c->views[v].d_gray, (size_t)L->w, L->w, L->h);
                cv.dstArray = L->arr;
                cv.dstPos = make_cudaPos(0, 0, v);
                cv.extent = make_cudaExtent(L->w, L->h, 1);
                cv.kind = cudaMemcpyDeviceToDevice;
                CU(cudaMemcpy3DAsync(&cv, c->stream));
            }
        } else if (L->u16) {
            CU(launch_float_to_u16x4(L->lin, L->half, (size_t)c->V * P, c->stream));
            cp.srcPtr = make_cudaPitchedPtr(L->half, (size_t)L->w * 2, L->w, L->h);
        } else if (L->fp16) {
            CU(launch_float_to_half(L->lin, L->half, (size_t)c->V * P, c->stream));
            cp.srcPtr = make_cudaPitchedPtr(L->half, (size_t)L->w * 2, L->w, L->h);
        } else {
            cp.srcPtr = make_cudaPitchedPtr(L->lin, (size_t)L->w * sizeof(float), L->w, L->h);
        }
        cp.dstArray = L->arr;
        cp.extent = make_cudaExtent(L->w, L->h, c->V);
        cp.kind = cudaMemcpyDeviceToDevice;
        if (!L->u8) CU(cudaMemcpy3DAsync(&cp, c->stream));
        L->stale = false;
    }
    c->level_scale = scale; c->lw = L->w; c->lh = L->h;
    c->level_arr = L->arr; c->level_tex = L->tex; c->d_level_lin = L->lin;
    c->level_unorm = L->u8 ? 255.0f * 256.0f : (L->u16 ? 65535.0f * 256.0f : 0.0f);
    c->level_inv = L->u8 ? 1.0f / 256.0f : 1.0f / 1024.0f;
    return APDE_OK;
}

// ------------------------------------------------------------------------------------------------ view maps
int apde_view_download(apde_context *c, int view, float *depth, float *normal, uint8_t *weak, uint8_t *conf, int *width,
                       int *height) {
    if (!c || view < 0 || view >= c->V) return fail(APDE_ERR_ARG, "view_download: bad view");
    CU(cudaSetDevice(c->device));
    ViewStore &v = c->views[view];
    if (width) *width = v.mw;
    if (height) *height = v.mh;
    const size_t P = (size_t)v.mw * v.mh, Pfull = (size_t)c->W * c->H;
    if (P == 0) return APDE_OK;
    if (depth) CU(cudaMemcpyAsync(depth, c->d_depth_pool[c->cur] + (size_t)view * Pfull, P * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (normal) CU(cudaMemcpyAsync(normal, v.d_normal, P * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (weak) CU(cudaMemcpyAsync(weak, v.d_weak, P, cudaMemcpyDeviceToHost, c->stream));
    if (conf) CU(cudaMemcpyAsync(conf, v.d_conf, P, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return APDE_OK;
}

int apde_view_upload(apde_context *c, int view, const float *depth, const float *normal, const uint8_t *weak,
                     const uint8_t *conf, int width, int height) {
    if (!c || view < 0 || view >= c->V || width <= 0 || height <= 0 || width > c->W || height > c->H)
        return fail(APDE_ERR_ARG, "view_upload: bad argument");
    CU(cudaSetDevice(c->device));
    ViewStore &v = c->views[view];
    const size_t P = (size_t)width * height, Pfull = (size_t)c->W * c->H;
    // on c->stream (non-blocking: a legacy-stream copy from pageable memory is not ordered against it), then synchronised so that
    // the caller's buffers may be re-used at once
    if (depth) CU(cudaMemcpyAsync(c->d_depth_pool[c->cur] + (size_t)view * Pfull, depth, P * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    if (normal) CU(cudaMemcpyAsync(v.d_normal, normal, P * 3 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    if (weak) CU(cudaMemcpyAsync(v.d_weak, weak, P, cudaMemcpyHostToDevice, c->stream));
    if (conf) { CU(cudaMemcpyAsync(v.d_conf, conf, P, cudaMemcpyHostToDevice, c->stream)); v.has_conf = true; }
    CU(cudaStreamSynchronize(c->stream));
    v.mw = width; v.mh = height;
    if (depth) { v.dw = width; v.dh = height; }
    return APDE_OK;
}

// ------------------------------------------------------------------------------------------------ problem
// a0*b0 + a1*b1 + a2*b2 as the reference build contracts it on the device (FMUL of the middle product, FFMA of the first,
// FFMA of the last; see dot3_ref in apde_device.cuh).  IEEE fp32 multiply / fused multiply-add on the host give the same
// bits as the GPU's FMUL / FFMA (the operands here are far from the denormal range, so .FTZ does not matter).
static inline float dot3_ref_host(float a0, float b0, float a1, float b1, float a2, float b2) {
    volatile float mid = a1 * b1;  // volatile: the product must be rounded to fp32 on its own (no host-side contraction)
    return fmaf(a2, b2, fmaf(a0, b0, mid));
}

// Camera-pair part of ComputeHomography (APD.cu:336-362), computed once per (reference, source) pair instead of once per
// evaluation, in the reference's own fp32 operation order so that every bit of R_rel and t_rel agrees with what the
// reference's kernels compute on the fly:  C = -(R^T t);  R_rel = R_s R_r^T;  t_rel = R_s (C_r - C_s).
static void make_view_k(const apde_camera &rc, const apde_camera &sc, int layer, ViewK &vk) {
    float Sr[3], Ss[3];  // -C of both cameras
    for (int j = 0; j < 3; ++j) {
        Sr[j] = dot3_ref_host(rc.R[0 + j], rc.t[0], rc.R[3 + j], rc.t[1], rc.R[6 + j], rc.t[2]);
        Ss[j] = dot3_ref_host(sc.R[0 + j], sc.t[0], sc.R[3 + j], sc.t[1], sc.R[6 + j], sc.t[2]);
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            vk.Rrel[3 * i + j] = dot3_ref_host(sc.R[3 * i], rc.R[3 * j], sc.R[3 * i + 1], rc.R[3 * j + 1], sc.R[3 * i + 2], rc.R[3 * j + 2]);
    // C_rel = ref_C - src_C = (-Sr) - (-Ss) = Ss - Sr (exact either way)
    const float c0 = Ss[0] - Sr[0], c1 = Ss[1] - Sr[1], c2 = Ss[2] - Sr[2];
    for (int i = 0; i < 3; ++i) vk.trel[i] = dot3_ref_host(sc.R[3 * i], c0, sc.R[3 * i + 1], c1, sc.R[3 * i + 2], c2);
    for (int i = 0; i < 9; ++i) { vk.K[i] = sc.K[i]; vk.R[i] = sc.R[i]; }
    for (int i = 0; i < 3; ++i) { vk.t[i] = sc.t[i]; vk.c[i] = sc.c[i]; }
    vk.layer = layer;
}

static void scale_camera(apde_camera &cam, int W, int H, int lw, int lh) {  // APD.cpp:570-585
    if (lw == W && lh == H) { cam.width = W; cam.height = H; return; }
    const float sx = lw / static_cast<float>(W), sy = lh / static_cast<float>(H);
    cam.K[0] *= sx; cam.K[2] *= sx; cam.K[4] *= sy; cam.K[5] *= sy;
    cam.width = lw; cam.height = lh;
}

int apde_problem_setup(apde_context *c, int ref_view, const apde_params *params, int scale_size, uint32_t seed) {
    if (!c || !params) return fail(APDE_ERR_ARG, "problem_setup: null argument");
    if (!c->committed) return fail(APDE_ERR_STATE, "problem_setup: scene not committed");
    if (ref_view < 0 || ref_view >= c->V) return fail(APDE_ERR_ARG, "problem_setup: bad reference view %d", ref_view);
    if (scale_size < 1) return fail(APDE_ERR_ARG, "problem_setup: bad scale_size");
    if (params->strong_radius != 5 || params->strong_increment != 2 || params->weak_radius != 5 || params->weak_increment != 5)
        return fail(APDE_ERR_ARG, "problem_setup: only the reference patch geometry (radius 5, increments 2/5) is supported");
    CU(cudaSetDevice(c->device));
    ViewStore &rv = c->views[ref_view];
    const int N = (int)rv.src.size();
    if (N < 1) return fail(APDE_ERR_ARG, "problem_setup: view %d has no source views", ref_view);
    int rc = ensure_level(c, scale_size);
    if (rc) return rc;
    const int w = c->lw, h = c->lh;
    const size_t P = (size_t)w * h, Pfull = (size_t)c->W * c->H;
    cudaStream_t st = c->stream;

    c->params = *params;
    c->cams.resize(N + 1);
    for (int i = 0; i <= N; ++i) {
        const int vid = (i == 0) ? ref_view : rv.src[i - 1];
        c->cams[i] = c->views[vid].cam;
        scale_camera(c->cams[i], c->W, c->H, w, h);
    }
    c->params.depth_min = c->cams[0].depth_min * 0.6f;  // APD.cpp:554-556
    c->params.depth_max = c->cams[0].depth_max * 1.2f;
    c->params.num_images = N + 1;

    PassK &K = c->K;
    memset(&K, 0, sizeof(K));
    K.W = w; K.H = h; K.N = N;
    K.state = c->params.state; K.geom = c->params.geom_consistency; K.impetus = c->params.use_impetus;
    K.use_apd = c->params.use_APD; K.top_k = c->params.top_k; K.weak_peak_radius = c->params.weak_peak_radius;
    K.rotate_time = c->params.rotate_time; K.max_iterations = c->params.max_iterations;
    K.depth_min = c->params.depth_min; K.depth_max = c->params.depth_max;
    K.geom_factor = c->params.geom_factor; K.ransac_threshold = c->params.ransac_threshold;
    K.fx = c->cams[0].K[0]; K.fy = c->cams[0].K[4]; K.cx = c->cams[0].K[2]; K.cy = c->cams[0].K[5];
    for (int i = 0; i < 9; ++i) { K.R[i] = c->cams[0].R[i]; K.Kr[i] = c->cams[0].K[i]; }
    for (int i = 0; i < 3; ++i) { K.t[i] = c->cams[0].t[i]; K.c[i] = c->cams[0].c[i]; }
    K.ref_layer = ref_view;
    K.seed = seed;
    K.stream = (uint32_t)ref_view;
    K.tex = c->level_tex;
    {  // experiment knob: shape of the implicit checkerboard tiles (3 = 8x8 default ... 6 = 64x1)
        static const int ts = [] { const char *e = getenv("APDE_TILE_SHIFT"); const int v = e ? atoi(e) : 3; return (v >= 2 && v <= 6) ? v : 3; }();
        K.tile_shift = ts;
    }
    K.tex_unorm = c->level_unorm;
    K.tex_inv = c->level_inv;
    K.planes = c->d_planes; K.costs = c->d_costs; K.sel = c->d_sel; K.vw = c->d_vw; K.weak = c->d_weak; K.conf = c->d_conf;
    K.fit = c->d_fit; K.reliable = c->d_reliable; K.nearest = c->d_nearest; K.ns_tiles = c->d_ns_tiles; K.anchors = c->d_anchors;
    K.depth = c->d_depthws; K.counters = c->d_counters + 4 * kStages;
    for (int i = 0; i < N; ++i) make_view_k(c->cams[0], c->cams[i + 1], rv.src[i], K.v[i]);

    const bool need_depth = c->params.geom_consistency || c->params.use_APD;
    if (need_depth) {  // APD.cpp:592-610
        for (int i = 0; i <= N; ++i) {
            const int vid = (i == 0) ? ref_view : rv.src[i - 1];
            const ViewStore &sv = c->views[vid];
            if (sv.dw == 0) return fail(APDE_ERR_STATE, "problem_setup: view %d has no depth map yet", vid);
            CU(launch_resize_nearest(c->d_depth_pool[c->cur] + (size_t)vid * Pfull, sv.dw, sv.dh, c->d_depthws + (size_t)i * P, w, h, 4, st));
        }
        c->launches += N + 1;
    }
    if (c->params.use_APD) {  // APD.cpp:614-626
        if (rv.mw == 0 || !rv.has_conf) return fail(APDE_ERR_STATE, "problem_setup: use_APD needs weak/confidence maps of view %d", ref_view);
        CU(launch_resize_nearest(rv.d_weak, rv.mw, rv.mh, c->d_weak, w, h, 1, st));
        CU(launch_resize_nearest(rv.d_conf, rv.mw, rv.mh, c->d_conf, w, h, 1, st));
        CU(cudaMemsetAsync(c->d_fit, 0, P * sizeof(float4), st));  // APD.cpp:771
        c->launches += 2;
        if (c->params.use_sa && rv.d_sa) {  // APD.cpp:641-649: the file's label map, nearest-resized to the working size
            if (!c->d_sa) CU(cudaMalloc(&c->d_sa, Pfull));
            CU(launch_resize_nearest(rv.d_sa, rv.saw, rv.sah, c->d_sa, w, h, 1, st));
            c->launches += 1;
            K.sa = c->d_sa;
        }
    } else {  // APD.cpp:656-657
        CU(launch_fill_u8(c->d_weak, APDE_STRONG, P, st));
        CU(launch_fill_u8(c->d_conf, 1, P, st));
    }
    if (c->params.state != APDE_FIRST_INIT) {  // APD.cpp:662-683
        if (rv.mw == 0) return fail(APDE_ERR_STATE, "problem_setup: view %d has no previous depth/normal", ref_view);
        const float *dsrc = c->d_depth_pool[c->cur] + (size_t)ref_view * Pfull;
        const float *nsrc = rv.d_normal;
        if (rv.mw != w || rv.mh != h) {
            CU(launch_resize_nearest(dsrc, rv.dw, rv.dh, c->d_scratch_depth, w, h, 4, st));
            CU(launch_resize_nearest(nsrc, rv.mw, rv.mh, c->d_scratch_normal, w, h, 12, st));
            dsrc = c->d_scratch_depth; nsrc = c->d_scratch_normal;
            c->launches += 2;
        }
        CU(launch_planes_from_maps(dsrc, nsrc, c->d_planes, (int)P, st));
        c->launches += 1;
    } else {
        CU(cudaMemsetAsync(c->d_planes, 0, P * sizeof(float4), st));
    }
    c->ref_view = ref_view;
    c->problem_active = true;
    c->lists_dirty = true;
    c->sweep.valid = false;
    return APDE_OK;
}

int apde_problem_dims(apde_context *c, int *width, int *height, int *num_images) {
    if (!c || !c->problem_active) return fail(APDE_ERR_STATE, "no active problem");
    if (width) *width = c->K.W;
    if (height) *height = c->K.H;
    if (num_images) *num_images = c->K.N + 1;
    return APDE_OK;
}

// (re)build the four (colour, class) pixel lists and fetch their lengths.  The 16-byte read-back costs one stream sync per
// build (twice per pass), and saves launching worst-case grids for the sparse WEAK class: the column pipeline of a weak
// half-sweep is ~10 launches over [N][capacity] arrays, ~1 ms each time even when the frame holds no WEAK pixel at all.
static int build_lists(apde_context *c) {
    CU(launch_build_lists(c->K, c->d_lists, c->d_list_counts, c->list_cap, c->list_scratch, c->stream));
    CU(cudaMemcpyAsync(c->h_list_counts, c->d_list_counts, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->launches += 2;  // k_list_count + k_list_fill (the scan between them is cub's)
    c->lists_dirty = false;
    return APDE_OK;
}

int apde_problem_stage(apde_context *c, int stage, int iter, int color) {
    if (!c || !c->problem_active) return fail(APDE_ERR_STATE, "problem_stage: no active problem");
    if (stage < 0 || stage >= kStages) return fail(APDE_ERR_ARG, "problem_stage: unknown stage %d", stage);
    CU(cudaSetDevice(c->device));
    PassK Kl = c->K;
    Kl.counters = c->d_counters + 4 * stage;
    if (stage == APDE_STAGE_GEN_ANCHORS) c->lists_dirty = true;  // NeigbourUpdate turns unreliable WEAK pixels into UNKNOWN
    static const bool legacy_prop = [] { const char *e = getenv("APDE_LEGACY_PROP"); return e && e[0] == '1'; }();
    const bool prop_stage = stage == APDE_STAGE_PROP_STRONG || stage == APDE_STAGE_PROP_WEAK;
    // measured on B200 (r01, 960x540, 5 src): the column pipeline wins for the sparse WEAK class (81 vs 116 ms per step:
    // N x more parallelism for few pixels); for the dense strong class the fused one-thread-per-pixel kernel is faster (165 vs
    // 179 ms: no intermediate buffers, no second pass over the reference patch).  APDE_PIPELINE_STRONG=1 forces the pipeline.
    static const bool pipe_strong = [] { const char *e = getenv("APDE_PIPELINE_STRONG"); return e && e[0] == '1'; }();
    const bool pipeline = prop_stage && !legacy_prop && (stage == APDE_STAGE_PROP_WEAK || pipe_strong);
    if (prop_stage && (c->params.use_APD || pipeline)) {
        if (c->lists_dirty) { const int rc = build_lists(c); if (rc) return rc; }
        const int cls = (color << 1) | (stage == APDE_STAGE_PROP_WEAK ? 1 : 0);
        // The fused strong kernel walks the list only when enough lanes of the implicit 8x8-tile mapping would idle on WEAK
        // pixels.  Measured on B200 (r01, 1920x1080, 10 src, ~1 % WEAK): tile mapping 1731 ms per step vs 1896 ms through the
        // list (72.8 % vs 66.5 % of the gather peak) -- list slots are appended in atomic order, so neighbouring CTAs work on
        // distant tiles and warps straddle tiles.  APDE_STRONG_LIST_SHARE = share of same-colour pixels below which the list
        // is used (0 = never, 1 = always; default 0.8).
        static const float list_share = [] { const char *e = getenv("APDE_STRONG_LIST_SHARE"); return e ? (float)atof(e) : 0.8f; }();
        const int ylimit = std::min(c->K.H, 32 * (((c->K.H / 2) + 15) / 16));  // rows the half-sweeps cover (quirk 7, APD.cu:2676-2678)
        const bool dense = stage == APDE_STAGE_PROP_STRONG && !pipeline &&
                           (float)c->h_list_counts[cls] >= list_share * 0.5f * (float)c->K.W * (float)ylimit;
        if (!dense) {
            Kl.list = c->d_lists + (size_t)cls * c->list_cap;
            Kl.list_count = c->d_list_counts + cls;
        }
    }
    // WEAK-only full-frame stages (GenAnchors, RANSAC plane fit) walk the two (colour, weak) lists instead of the whole grid.
    // The lists hold rows [0, half_rows_limit): only usable when that covers the frame (quirk 7 shapes fall back).
    const bool weak_lists = (stage == APDE_STAGE_GEN_ANCHORS || stage == APDE_STAGE_RANSAC_FIT) && c->params.use_APD &&
                            !getenv("APDE_NO_WEAK_LISTS") && 32 * (((c->K.H / 2) + 15) / 16) >= c->K.H;
    if (weak_lists && c->lists_dirty) { const int rc = build_lists(c); if (rc) return rc; }
    auto run_stage = [&](const PassK &Kq) -> cudaError_t {
        static const bool legacy = [] { const char *e = getenv("APDE_LEGACY_SWEEP"); return e && e[0] == '1'; }();
        if (weak_lists) {
            // both weak lists (black = class 1, red = class 3) in ONE launch, blockIdx.y = colour: the few long-running threads
            // of a sparse class overlap instead of serialising two launch tails
            PassK Kw = Kq;
            Kw.list = c->d_lists + (size_t)1 * c->list_cap;
            Kw.list_count = c->d_list_counts + 1;
            Kw.list_pair_stride = 2 * c->list_cap;
            const int longest = std::max(c->h_list_counts[1], c->h_list_counts[3]);
            if (longest == 0 && stage == APDE_STAGE_GEN_ANCHORS) return cudaSuccess;  // no WEAK pixel: nothing to do
            if (stage == APDE_STAGE_RANSAC_FIT) c->launches += 1;  // fit copy + fit (+1 counted below)
            return launch_stage(Kw, stage, iter, longest, c->stream, nullptr);  // `color` carries the list length here
        }
        if (pipeline) {
            cudaError_t e = c->prop.reserve(c->list_cap, Kq.N);
            if (e != cudaSuccess) return e;
            const int cls = (color << 1) | (stage == APDE_STAGE_PROP_WEAK ? 1 : 0);
            return prop_half_sweep(Kq, c->prop, stage == APDE_STAGE_PROP_WEAK, Kq.list, Kq.list_count, c->h_list_counts[cls], iter, c->stream,
                                   &c->launches);
        }
        if (!legacy) {
            // (the sweep functions count every launch but the last one; the generic count below adds that)
            if (stage == APDE_STAGE_DEPTH_TO_WEAK) {
                float *curve = nullptr;
                if (c->capture_curve) {
                    const size_t need = (size_t)Kq.W * Kq.H * 61;
                    if (need > c->curve_cap) {
                        cudaFree(c->d_curve);
                        c->d_curve = nullptr; c->curve_cap = 0;
                        cudaError_t e = cudaMalloc(&c->d_curve, need * sizeof(float));
                        if (e != cudaSuccess) return e;
                        c->curve_cap = need;
                    }
                    cudaError_t e = cudaMemsetAsync(c->d_curve, 0, need * sizeof(float), c->stream);
                    if (e != cudaSuccess) return e;
                    curve = c->d_curve;
                }
                return sweep_depth_to_weak(Kq, c->sweep, curve, c->stream, &c->launches);
            }
            if (stage == APDE_STAGE_LOCAL_REFINE) return sweep_local_refine(Kq, c->sweep, c->stream, &c->launches);
        }
        return launch_stage(Kq, stage, iter, color, c->stream, nullptr);
    };
    // the stored column costs stay valid only across ConfidenceCompute (it touches neither planes nor view weights)
    if (stage != APDE_STAGE_DEPTH_TO_WEAK && stage != APDE_STAGE_LOCAL_REFINE && stage != APDE_STAGE_CONFIDENCE) c->sweep.valid = false;
    if (c->profiling) {
        const size_t k = c->ev_stage.size();
        while (c->ev_pool.size() < 2 * (k + 1)) {
            cudaEvent_t e;
            CU(cudaEventCreate(&e));
            c->ev_pool.push_back(e);
        }
        CU(cudaEventRecord(c->ev_pool[2 * k], c->stream));
        CU(run_stage(Kl));
        CU(cudaEventRecord(c->ev_pool[2 * k + 1], c->stream));
        c->ev_stage.push_back(stage);
    } else {
        CU(run_stage(Kl));
    }
    const int nl = ((stage == APDE_STAGE_INIT && Kl.use_apd) || stage == APDE_STAGE_NEAREST_STRONG) ? 2 : 1;  // two-phase init (see k_init); tile summary + search
    c->launches += nl;
    c->stage_launches[stage] += nl;
    if (stage == APDE_STAGE_GEN_ANCHORS) c->lists_dirty = true;  // NeigbourUpdate turned unreliable WEAK pixels into UNKNOWN
    return APDE_OK;
}

static int collect_stage_events(apde_context *c) {
    if (c->ev_stage.empty()) return APDE_OK;
    CU(cudaStreamSynchronize(c->stream));
    for (size_t k = 0; k < c->ev_stage.size(); ++k) {
        float ms = 0.0f;
        CU(cudaEventElapsedTime(&ms, c->ev_pool[2 * k], c->ev_pool[2 * k + 1]));
        c->stage_ms[c->ev_stage[k]] += ms;
    }
    c->ev_stage.clear();
    return APDE_OK;
}

int apde_problem_capture_curve(apde_context *c, int on) {
    if (!c) return fail(APDE_ERR_ARG, "capture_curve: no context");
    c->capture_curve = on != 0;
    if (!on) { CU(cudaSetDevice(c->device)); cudaFree(c->d_curve); c->d_curve = nullptr; c->curve_cap = 0; }
    return APDE_OK;
}

int apde_set_sweep_budget_mb(apde_context *c, size_t megabytes) {
    if (!c) return fail(APDE_ERR_ARG, "set_sweep_budget_mb: no context");
    c->sweep.budget_mb = megabytes;
    c->sweep.valid = false;
    return APDE_OK;
}

int apde_set_profiling(apde_context *c, int on) {
    if (!c) return fail(APDE_ERR_ARG, "set_profiling: null context");
    int rc = collect_stage_events(c);
    if (rc) return rc;
    c->profiling = on != 0;
    return APDE_OK;
}

int apde_get_stage_stats(apde_context *c, double *ms, uint64_t *launches, uint64_t *evals, int reset) {
    if (!c) return fail(APDE_ERR_ARG, "get_stage_stats: null context");
    CU(cudaSetDevice(c->device));
    int rc = collect_stage_events(c);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    unsigned long long h[kCounterWords];
    CU(cudaMemcpy(h, c->d_counters, sizeof(h), cudaMemcpyDeviceToHost));
    for (int s = 0; s < kStages; ++s) {
        if (ms) ms[s] = c->stage_ms[s];
        if (launches) launches[s] = c->stage_launches[s];
        if (evals) for (int k = 0; k < 3; ++k) evals[3 * s + k] = h[4 * s + k];
    }
    if (reset) {
        CU(cudaMemset(c->d_counters, 0, sizeof(h)));
        for (int s = 0; s < 16; ++s) { c->stage_ms[s] = 0; c->stage_launches[s] = 0; }
    }
    return APDE_OK;
}

// APD::RunPatchMatch, APD.cu:2663-2737
int apde_problem_run(apde_context *c) {
    if (!c || !c->problem_active) return fail(APDE_ERR_STATE, "problem_run: no active problem");
    CU(cudaSetDevice(c->device));
    const apde_params &p = c->params;
    int rc;
#define ST(stage, it, col) if ((rc = apde_problem_stage(c, stage, it, col))) return rc;
    if (p.use_APD) { ST(APDE_STAGE_NEAREST_STRONG, 0, 0) ST(APDE_STAGE_GEN_ANCHORS, 0, 0) }
    ST(APDE_STAGE_INIT, 0, 0)
    for (int i = 0; i < p.max_iterations; ++i) {
        ST(APDE_STAGE_PROP_STRONG, i, 0)
        ST(APDE_STAGE_PROP_STRONG, i, 1)
        if (p.use_APD) { ST(APDE_STAGE_RANSAC_FIT, i, 0) ST(APDE_STAGE_PROP_WEAK, i, 0) ST(APDE_STAGE_PROP_WEAK, i, 1) }
    }
    ST(APDE_STAGE_DEPTH_NORMAL, 0, 0)
    ST(APDE_STAGE_MEDIAN, 0, 0)
    ST(APDE_STAGE_MEDIAN, 0, 1)
    ST(APDE_STAGE_DEPTH_TO_WEAK, 0, 0)
    if (p.geom_consistency || p.use_APD) ST(APDE_STAGE_CONFIDENCE, 0, 0)
    ST(APDE_STAGE_LOCAL_REFINE, 0, 0)
#undef ST
    if (c->profiling && c->ev_stage.size() > 512) return collect_stage_events(c);
    return APDE_OK;
}

__global__ void k_unpack_vw(const uint4 *__restrict__ vw, uint8_t *__restrict__ out, int P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const uint4 w = vw[i];
    for (int v = 0; v < 32; ++v) out[(size_t)i * 32 + v] = (uint8_t)vw_get(w, v);
}
__global__ void k_pack_vw(const uint8_t *__restrict__ in, uint4 *__restrict__ vw, int P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    uint32_t w[4] = {0, 0, 0, 0};
    for (int v = 0; v < 32; ++v) w[v >> 3] |= (uint32_t)(in[(size_t)i * 32 + v] & 15u) << ((v & 7) * 4);
    vw[i] = make_uint4(w[0], w[1], w[2], w[3]);
}

static int field_ptr(apde_context *c, int field, void **ptr, size_t *bytes) {
    const size_t P = (size_t)c->K.W * c->K.H;
    switch (field) {
        case APDE_FIELD_PLANES: *ptr = c->d_planes; *bytes = P * 16; break;
        case APDE_FIELD_COSTS: *ptr = c->d_costs; *bytes = P * 4; break;
        case APDE_FIELD_SELECTED_VIEWS: *ptr = c->d_sel; *bytes = P * 4; break;
        case APDE_FIELD_WEAK_INFO: *ptr = c->d_weak; *bytes = P; break;
        case APDE_FIELD_CONFIDENCE: *ptr = c->d_conf; *bytes = P; break;
        case APDE_FIELD_FIT_PLANES: *ptr = c->d_fit; *bytes = P * 16; break;
        case APDE_FIELD_WEAK_RELIABLE: *ptr = c->d_reliable; *bytes = P; break;
        case APDE_FIELD_NEAREST_STRONG: *ptr = c->d_nearest; *bytes = P * 4; break;
        case APDE_FIELD_ANCHORS: *ptr = c->d_anchors; *bytes = P * APDE_ANCHOR_NUM * 4; break;
        case APDE_FIELD_SRC_DEPTH: *ptr = c->d_depthws; *bytes = P * 4 * (c->K.N + 1); break;
        case APDE_FIELD_IMAGE: *ptr = c->d_level_lin + (size_t)c->ref_view * P; *bytes = P * 4; break;
        case APDE_FIELD_RELIABLE_CURVE:
            if (!c->capture_curve || c->curve_cap < P * 61) return fail(APDE_ERR_STATE, "no reliable curve captured (apde_problem_capture_curve)");
            *ptr = c->d_curve; *bytes = P * 61 * 4; break;
        case APDE_FIELD_SA_MASK:
            if (!c->K.sa) return fail(APDE_ERR_STATE, "the active problem has no segment-label map");
            *ptr = c->d_sa; *bytes = P; break;
        default: return fail(APDE_ERR_ARG, "unknown field %d", field);
    }
    return APDE_OK;
}

int apde_problem_get(apde_context *c, int field, void *host, size_t bytes) {
    if (!c || !c->problem_active || !host) return fail(APDE_ERR_STATE, "problem_get: no active problem");
    CU(cudaSetDevice(c->device));
    const size_t P = (size_t)c->K.W * c->K.H;
    CU(cudaStreamSynchronize(c->stream));
    if (field == APDE_FIELD_VIEW_WEIGHT) {
        if (bytes != P * 32) return fail(APDE_ERR_ARG, "problem_get: view_weight needs %zu bytes", P * 32);
        uint8_t *tmp;
        CU(cudaMalloc(&tmp, P * 32));
        k_unpack_vw<<<(unsigned)((P + 255) / 256), 256, 0, c->stream>>>(c->d_vw, tmp, (int)P);
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaMemcpy(host, tmp, P * 32, cudaMemcpyDeviceToHost));
        cudaFree(tmp);
        return APDE_OK;
    }
    void *ptr; size_t need;
    int rc = field_ptr(c, field, &ptr, &need);
    if (rc) return rc;
    if (bytes != need) return fail(APDE_ERR_ARG, "problem_get: field %d needs %zu bytes, got %zu", field, need, bytes);
    CU(cudaMemcpy(host, ptr, need, cudaMemcpyDeviceToHost));
    return APDE_OK;
}

int apde_problem_set(apde_context *c, int field, const void *host, size_t bytes) {
    if (!c || !c->problem_active || !host) return fail(APDE_ERR_STATE, "problem_set: no active problem");
    CU(cudaSetDevice(c->device));
    const size_t P = (size_t)c->K.W * c->K.H;
    CU(cudaStreamSynchronize(c->stream));
    if (field == APDE_FIELD_VIEW_WEIGHT) {
        if (bytes != P * 32) return fail(APDE_ERR_ARG, "problem_set: view_weight needs %zu bytes", P * 32);
        uint8_t *tmp;
        CU(cudaMalloc(&tmp, P * 32));
        CU(cudaMemcpy(tmp, host, P * 32, cudaMemcpyHostToDevice));
        k_pack_vw<<<(unsigned)((P + 255) / 256), 256, 0, c->stream>>>(tmp, c->d_vw, (int)P);
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(tmp);
        return APDE_OK;
    }
    if (field == APDE_FIELD_IMAGE) return fail(APDE_ERR_ARG, "problem_set: field is read only");
    if (field == APDE_FIELD_WEAK_INFO) c->lists_dirty = true;
    c->sweep.valid = false;
    void *ptr; size_t need;
    int rc = field_ptr(c, field, &ptr, &need);
    if (rc) return rc;
    if (bytes != need) return fail(APDE_ERR_ARG, "problem_set: field %d needs %zu bytes, got %zu", field, need, bytes);
    CU(cudaMemcpy(ptr, host, need, cudaMemcpyHostToDevice));
    return APDE_OK;
}

int apde_problem_get_image(apde_context *c, int idx, float *host, size_t bytes) {
    if (!c || !c->problem_active || !host) return fail(APDE_ERR_STATE, "problem_get_image: no active problem");
    if (idx < 0 || idx > c->K.N) return fail(APDE_ERR_ARG, "problem_get_image: bad index");
    const size_t P = (size_t)c->K.W * c->K.H;
    if (bytes != P * 4) return fail(APDE_ERR_ARG, "problem_get_image: needs %zu bytes", P * 4);
    CU(cudaSetDevice(c->device));
    const int vid = (idx == 0) ? c->ref_view : c->views[c->ref_view].src[idx - 1];
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpy(host, c->d_level_lin + (size_t)vid * P, P * 4, cudaMemcpyDeviceToHost));
    return APDE_OK;
}

int apde_problem_get_cameras(apde_context *c, apde_camera *cams, apde_params *params) {
    if (!c || !c->problem_active) return fail(APDE_ERR_STATE, "problem_get_cameras: no active problem");
    if (cams) memcpy(cams, c->cams.data(), sizeof(apde_camera) * c->cams.size());
    if (params) *params = c->params;
    return APDE_OK;
}

static int problem_finish_impl(apde_context *c, bool jacobi) {
    if (!c || !c->problem_active) return fail(APDE_ERR_STATE, "problem_finish: no active problem");
    CU(cudaSetDevice(c->device));
    ViewStore &rv = c->views[c->ref_view];
    const size_t P = (size_t)c->K.W * c->K.H, Pfull = (size_t)c->W * c->H;
    float *dpool = c->d_depth_pool[jacobi ? (c->cur ^ 1) : c->cur];
    CU(launch_finish(c->d_planes, c->d_weak, dpool + (size_t)c->ref_view * Pfull, rv.d_normal, rv.d_weak, (int)P,
                     c->params.depth_min, c->params.depth_max, c->stream));
    c->launches++;
    if (c->params.geom_consistency || c->params.use_APD) {  // main.cpp:187-190
        CU(cudaMemcpyAsync(rv.d_conf, c->d_conf, P, cudaMemcpyDeviceToDevice, c->stream));
        rv.has_conf = true;
    }
    rv.mw = c->K.W; rv.mh = c->K.H;
    if (!jacobi) { rv.dw = rv.mw; rv.dh = rv.mh; }
    c->problem_active = false;
    return APDE_OK;
}

int apde_problem_finish(apde_context *c) { return problem_finish_impl(c, false); }

int apde_pass_run(apde_context *c, int ref_view, const apde_params *params, int scale_size, uint32_t seed) {
    int rc = apde_problem_setup(c, ref_view, params, scale_size, seed);
    if (rc) return rc;
    if ((rc = apde_problem_run(c))) return rc;
    return apde_problem_finish(c);
}

int apde_eval_costs(apde_context *c, int n, const int32_t *tuples, const float *planes, int mode, float *out) {
    if (!c || !c->problem_active) return fail(APDE_ERR_STATE, "eval_costs: no active problem");
    if (n <= 0 || !tuples || !planes || !out || mode < 0 || mode > 2) return fail(APDE_ERR_ARG, "eval_costs: bad argument");
    for (int i = 0; i < n; ++i) {
        const int x = tuples[3 * i], y = tuples[3 * i + 1], v = tuples[3 * i + 2];
        if (x < 0 || x >= c->K.W || y < 0 || y >= c->K.H || v < 1 || v > c->K.N)
            return fail(APDE_ERR_ARG, "eval_costs: tuple %d (%d, %d, %d) out of range", i, x, y, v);
    }
    CU(cudaSetDevice(c->device));
    int *d_t; float4 *d_p; float *d_o;
    CU(cudaMalloc(&d_t, (size_t)n * 12));
    CU(cudaMalloc(&d_p, (size_t)n * 16));
    CU(cudaMalloc(&d_o, (size_t)n * 4));
    CU(cudaMemcpyAsync(d_t, tuples, (size_t)n * 12, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_p, planes, (size_t)n * 16, cudaMemcpyHostToDevice, c->stream));
    CU(launch_eval_costs(c->K, n, d_t, d_p, mode, d_o, c->stream));
    c->launches++;
    CU(cudaMemcpyAsync(out, d_o, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(d_t); cudaFree(d_p); cudaFree(d_o);
    return APDE_OK;
}

__global__ void k_debug_tex(const __grid_constant__ PassK K, int layer, int n, const float2 *xy, float *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = K.tex_unorm > 0.0f ? fetch<true>(K, xy[i].x, xy[i].y, layer) : fetch<false>(K, xy[i].x, xy[i].y, layer);
}

int apde_debug_tex2d(apde_context *c, int idx, int n, const float *xy, float *out) {
    if (!c || !c->problem_active) return fail(APDE_ERR_STATE, "debug_tex2d: no active problem");
    if (idx < 0 || idx > c->K.N || n <= 0 || !xy || !out) return fail(APDE_ERR_ARG, "debug_tex2d: bad argument");
    CU(cudaSetDevice(c->device));
    const int layer = (idx == 0) ? c->ref_view : c->views[c->ref_view].src[idx - 1];
    float2 *d_xy; float *d_o;
    CU(cudaMalloc(&d_xy, (size_t)n * 8));
    CU(cudaMalloc(&d_o, (size_t)n * 4));
    CU(cudaMemcpyAsync(d_xy, xy, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    k_debug_tex<<<(n + 255) / 256, 256, 0, c->stream>>>(c->K, layer, n, d_xy, d_o);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, d_o, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(d_xy); cudaFree(d_o);
    return APDE_OK;
}

int apde_microbench(apde_context *c, double *fp32_tflops, double *tex_gsamples) {
    if (!c) return fail(APDE_ERR_ARG, "microbench: null context");
    CU(cudaSetDevice(c->device));
    if (fp32_tflops) CU(microbench_fp32(fp32_tflops, c->stream));
    if (tex_gsamples) {
        if (!c->level_tex) return fail(APDE_ERR_STATE, "microbench: no pyramid level built yet");
        CU(microbench_tex(c->level_tex, 0, c->lw, c->lh, tex_gsamples, c->stream));
    }
    return APDE_OK;
}

int apde_microbench_pattern(apde_context *c, int mode, float spread, double *tex_gsamples) {
    if (!c || !tex_gsamples || !c->level_tex) return fail(APDE_ERR_STATE, "microbench_pattern: no pyramid level built yet");
    CU(cudaSetDevice(c->device));
    CU(microbench_tex_pattern(c->level_tex, 0, c->lw, c->lh, mode, spread, tex_gsamples, c->stream));
    return APDE_OK;
}

int apde_get_counters(apde_context *c, uint64_t out[4], int reset) {
    if (!c || !out) return fail(APDE_ERR_ARG, "get_counters: null argument");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    unsigned long long h[kCounterWords];
    CU(cudaMemcpy(h, c->d_counters, sizeof(h), cudaMemcpyDeviceToHost));
    out[0] = out[1] = out[2] = 0;
    for (int s = 0; s <= kStages; ++s) { out[0] += h[4 * s]; out[1] += h[4 * s + 1]; out[2] += h[4 * s + 2]; }
    out[3] = c->launches;
    if (reset) { CU(cudaMemset(c->d_counters, 0, sizeof(h))); c->launches = 0; }
    return APDE_OK;
}

int apde_get_anchor_evals(apde_context *c, uint64_t *anchor_patches) {
    if (!c || !anchor_patches) return fail(APDE_ERR_ARG, "get_anchor_evals: null argument");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    unsigned long long h[kCounterWords];
    CU(cudaMemcpy(h, c->d_counters, sizeof(h), cudaMemcpyDeviceToHost));
    *anchor_patches = 0;
    for (int s = 0; s <= kStages; ++s) *anchor_patches += h[4 * s + 3];
    return APDE_OK;
}

int apde_depth_pool(apde_context *c, void **dev_ptr, size_t *bytes, size_t *bytes_per_view) {
    if (!c || !c->d_depth_pool[0]) return fail(APDE_ERR_STATE, "depth_pool: no scene");
    const size_t Pfull = (size_t)c->W * c->H;
    if (dev_ptr) *dev_ptr = c->d_depth_pool[c->cur];
    if (bytes) *bytes = (size_t)c->V * Pfull * sizeof(float);
    if (bytes_per_view) *bytes_per_view = Pfull * sizeof(float);
    return APDE_OK;
}

// ------------------------------------------------------------------------------------------------ schedule
static int compute_round_num(int W, int H) {  // main.cpp:129-146
    int max_size = W > H ? W : H;
    int r = 1;
    while (max_size > 800) { max_size /= 2; r++; }
    return r;
}

int apde_schedule_num_passes(apde_context *c, const apde_schedule *s) {
    if (!c || !s || c->V <= 0) return fail(APDE_ERR_ARG, "schedule_num_passes: bad argument");
    const int rounds = s->rounds > 0 ? s->rounds : compute_round_num(c->W, c->H);
    return rounds * (1 + s->geom_iterations);
}

int apde_schedule_pass_params(apde_context *c, const apde_schedule *s, int pass_index, apde_params *params, int *scale_size,
                              uint32_t *seed) {
    if (!c || !s || !params) return fail(APDE_ERR_ARG, "schedule_pass_params: null argument");
    const int rounds = s->rounds > 0 ? s->rounds : compute_round_num(c->W, c->H);
    const int per_round = 1 + s->geom_iterations;
    if (pass_index < 0 || pass_index >= rounds * per_round) return fail(APDE_ERR_ARG, "schedule_pass_params: pass %d out of range", pass_index);
    const int i = pass_index / per_round, j = pass_index % per_round - 1;  // j = -1: photometric pass
    apde_params p;
    apde_params_default(&p);
    p.geom_factor = s->geom_factor;
    p.use_impetus = s->use_impetus;
    p.use_sa = s->use_sa;  // main.cpp:324: consumed only for views that have a label map
    p.max_iterations = 3;
    if (i == 0) {
        p.use_APD = 0;
    } else {
        p.use_APD = 1;
        p.ransac_threshold = (float)(0.01 - i * 0.00125);                       // main.cpp:318
        p.rotate_time = std::min(static_cast<int>(std::pow(2, i)), 4);           // main.cpp:319
    }
    if (j < 0) {  // main.cpp:309-331
        p.state = (i == 0) ? APDE_FIRST_INIT : APDE_REFINE_INIT;
        p.geom_consistency = 0;
        p.weak_peak_radius = 6;
    } else {  // main.cpp:333-365
        p.state = APDE_REFINE_ITER;
        p.geom_consistency = 1;
        p.weak_peak_radius = std::max(4 - 2 * j, 2);
    }
    *params = p;
    if (scale_size) *scale_size = static_cast<int>(std::pow(2, rounds - 1 - i));
    if (seed) *seed = s->seed * 0x9E3779B1u + (uint32_t)pass_index * 0x85EBCA77u;
    return APDE_OK;
}

int apde_run_schedule_pass(apde_context *c, const apde_schedule *s, int pass_index, apde_timing *out) {
    if (!c || !s) return fail(APDE_ERR_ARG, "run_schedule_pass: null argument");
    if (!c->committed) return fail(APDE_ERR_STATE, "run_schedule_pass: scene not committed");
    CU(cudaSetDevice(c->device));
    const int rounds = s->rounds > 0 ? s->rounds : compute_round_num(c->W, c->H);
    const int per_round = 1 + s->geom_iterations;
    if (pass_index < 0 || pass_index >= rounds * per_round) return fail(APDE_ERR_ARG, "run_schedule_pass: pass %d out of range", pass_index);
    const int i = pass_index / per_round, j = pass_index % per_round - 1;  // j = -1: photometric pass
    (void)i; (void)j;
    int rc_tmp = 0;
    const size_t Pfull = (size_t)c->W * c->H;
    // a multi-GPU job (apde_comm_init): this rank's block of views, Jacobi order, depth maps exchanged view by view
    const bool job = apde_comm_attached(c);
    const bool jacobi = s->jacobi != 0 || job;
    if (jacobi && !c->d_depth_pool[1]) {
        CU(cudaMalloc(&c->d_depth_pool[1], (size_t)c->V * Pfull * sizeof(float)));
        CU(cudaMemsetAsync(c->d_depth_pool[1], 0, (size_t)c->V * Pfull * sizeof(float), c->stream));
    }
    apde_params p;
    int scale = 1;
    uint32_t seed = 0;
    {
        const int prc = apde_schedule_pass_params(c, s, pass_index, &p, &scale, &seed);
        if (prc) return prc;
    }
    int first = (s->num_views_local > 0) ? s->first_view : 0;
    int count = (s->num_views_local > 0) ? s->num_views_local : c->V;
    int longest = count;  // views in the largest block of the job: every rank takes part in that many exchange rounds
    if (job) {
        if (s->num_views_local > 0) return fail(APDE_ERR_ARG, "run_schedule_pass: a multi-GPU job deals the views out itself (num_views_local must be 0)");
        int rk = 0, wd = 1;
        if ((rc_tmp = apde_comm_info(c, &rk, &wd, &first, &count))) return rc_tmp;
        longest = (c->V + wd - 1) / wd;
    }
    if (first < 0 || first + count > c->V) return fail(APDE_ERR_ARG, "run_schedule_pass: bad shard [%d, %d)", first, first + count);

    // APDE_PROFILE_PASS=<pass>[:<occurrence>]: bracket that pass with cudaProfilerStart/Stop, for
    // `ncu --profile-from-start off` captures of one pass deep inside a run (profiles/README)
    static const int prof_pass = [] { const char *e = getenv("APDE_PROFILE_PASS"); return e ? atoi(e) : -1; }();
    static const int prof_occ = [] { const char *e = getenv("APDE_PROFILE_PASS"); const char *q = e ? strchr(e, ':') : nullptr; return q ? atoi(q + 1) : 1; }();
    static int prof_seen = 0;
    const bool prof_now = pass_index == prof_pass && ++prof_seen == prof_occ;
    if (prof_now) cudaProfilerStart();

    uint64_t c0[4];
    int rc = apde_get_counters(c, c0, 0);
    if (rc) return rc;
    const auto t0 = std::chrono::steady_clock::now();
    double pm_ms = 0.0;
    struct PassEvents {  // destroyed on every return path
        cudaEvent_t a = nullptr, b = nullptr;
        ~PassEvents() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    } pass_events;
    CU(cudaEventCreate(&pass_events.a));
    CU(cudaEventCreate(&pass_events.b));
    const cudaEvent_t evp0 = pass_events.a, evp1 = pass_events.b;
    CU(cudaEventRecord(evp0, c->stream));
    while ((int)c->pm_events.size() < 2 * std::max(count, 1)) {
        cudaEvent_t ev;
        CU(cudaEventCreate(&ev));
        c->pm_events.push_back(ev);
    }
    double exposed_ms = 0.0;
    uint64_t received = 0;
    const int pass_w = (int)std::round(c->W * (1.0f / (float)scale)), pass_h = (int)std::round(c->H * (1.0f / (float)scale));
    for (int k = 0; k < longest; ++k) {
        if (k < count) {
            const int v = first + k;
            if ((rc = apde_problem_setup(c, v, &p, scale, seed))) return rc;
            CU(cudaEventRecord(c->pm_events[2 * k], c->stream));
            if ((rc = apde_problem_run(c))) return rc;
            CU(cudaEventRecord(c->pm_events[2 * k + 1], c->stream));
            if ((rc = problem_finish_impl(c, jacobi))) return rc;  // no host sync: the host runs ahead of the stream
        }
        // the k-th views of all ranks travel while the (k + 1)-th are computed.  The row size is the pass's level size, derived
        // from the scale as ensure_level does -- a rank that holds no view in this job (more ranks than views) never set one up
        if (job && (rc = apde_comm_share_depth_row(c, k, c->cur ^ 1, pass_w, pass_h))) return rc;
    }
    if (job && (rc = apde_comm_join(c, &exposed_ms, &received))) return rc;  // the compute stream waits for the last rows
    CU(cudaEventRecord(evp1, c->stream));
    CU(cudaEventSynchronize(evp1));
    for (int k = 0; k < count; ++k) {
        float ms = 0.0f;
        CU(cudaEventElapsedTime(&ms, c->pm_events[2 * k], c->pm_events[2 * k + 1]));
        pm_ms += ms;
    }
    if (jacobi) {
        // views outside this rank's shard keep their previous depth until the host-side exchange overwrites them
        // (a job's exchange has already put the new rows there)
        if (count < c->V && !job) {
            for (int v = 0; v < c->V; ++v) {
                if (v >= first && v < first + count) continue;
                CU(cudaMemcpyAsync(c->d_depth_pool[c->cur ^ 1] + (size_t)v * Pfull, c->d_depth_pool[c->cur] + (size_t)v * Pfull,
                                   Pfull * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
                // the host-side exchange fills these at the pass resolution before the next pass reads them
                c->views[v].mw = c->lw; c->views[v].mh = c->lh;
            }
        }
        if (job) {
            for (int v = 0; v < c->V; ++v) {
                if (v >= first && v < first + count) continue;
                c->views[v].mw = pass_w; c->views[v].mh = pass_h;  // the received maps are at the pass resolution
            }
        }
        c->cur ^= 1;
        for (auto &v : c->views) { v.dw = v.mw; v.dh = v.mh; }
    }
    CU(cudaStreamSynchronize(c->stream));
    if (prof_now) cudaProfilerStop();
    float dev_ms = 0.0f;
    CU(cudaEventElapsedTime(&dev_ms, evp0, evp1));
    const auto t1 = std::chrono::steady_clock::now();
    if (out) {
        uint64_t c1[4];
        if ((rc = apde_get_counters(c, c1, 0))) return rc;
        out->patchmatch_ms += pm_ms;
        out->device_ms += dev_ms;
        out->total_ms += std::chrono::duration<double, std::milli>(t1 - t0).count();
        out->evals_ncc_old += c1[0] - c0[0];
        out->evals_ncc_new += c1[1] - c0[1];
        out->evals_geom += c1[2] - c0[2];
        out->kernel_launches += c1[3] - c0[3];
        out->passes += 1;
        out->exchange_ms += exposed_ms;
        out->exchange_bytes += received;
    }
    return APDE_OK;
}

int apde_run_schedule(apde_context *c, const apde_schedule *s, apde_timing *out) {
    const int n = apde_schedule_num_passes(c, s);
    if (n < 0) return n;
    if (out) memset(out, 0, sizeof(*out));
    for (int p = 0; p < n; ++p) {
        int rc = apde_run_schedule_pass(c, s, p, out);
        if (rc) return rc;
    }
    return APDE_OK;
}

// ------------------------------------------------------------------------------------------------ fusion
static int fusion_views(apde_context *c, std::vector<FusionView> &fv, int *mw, int *mh) {
    const size_t Pfull = (size_t)c->W * c->H;
    fv.resize(c->V);
    *mw = c->views[0].mw; *mh = c->views[0].mh;
    if (*mw == 0) return fail(APDE_ERR_STATE, "fusion: no depth maps yet");
    for (int v = 0; v < c->V; ++v) {
        const ViewStore &s = c->views[v];
        if (s.mw != *mw || s.mh != *mh) return fail(APDE_ERR_STATE, "fusion: view %d has a different map size", v);
        apde_camera cam = s.cam;
        scale_camera(cam, c->W, c->H, *mw, *mh);  // RescaleImageAndCamera, APD.cpp:844-864
        fv[v].cam = cam;
        fv[v].depth = c->d_depth_pool[c->cur] + (size_t)v * Pfull;
        fv[v].normal = s.d_normal;
        fv[v].weak = s.d_weak;
        fv[v].conf = s.d_conf;
        fv[v].bgr = (*mw == c->W && *mh == c->H) ? s.d_bgr : nullptr;  // colours only at full resolution
        fv[v].src = s.src;
    }
    return APDE_OK;
}

int apde_map_pool(apde_context *c, int which, void **device_ptr, size_t *bytes_total, size_t *bytes_per_view) {
    if (!c || c->V <= 0 || !device_ptr) return fail(APDE_ERR_STATE, "map_pool: no scene");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    const size_t P = (size_t)c->W * c->H;
    size_t per = 0;
    switch (which) {
        case APDE_POOL_DEPTH: *device_ptr = c->d_depth_pool[c->cur]; per = P * sizeof(float); break;
        case APDE_POOL_NORMAL: *device_ptr = c->d_normal_pool; per = P * 3 * sizeof(float); break;
        case APDE_POOL_WEAK: *device_ptr = c->d_weak_pool; per = P; break;
        case APDE_POOL_CONFIDENCE: *device_ptr = c->d_conf_pool; per = P; break;
        case APDE_POOL_SKIP:
            if (!c->d_skip) {
                CU(cudaMalloc(&c->d_skip, (size_t)c->V * P));
                CU(cudaMemset(c->d_skip, 0, (size_t)c->V * P));
            }
            *device_ptr = c->d_skip;
            per = c->views[0].mw ? (size_t)c->views[0].mw * c->views[0].mh : P;  // rows are strided by the MAP size
            break;
        default: return fail(APDE_ERR_ARG, "map_pool: unknown pool %d", which);
    }
    if (bytes_total) *bytes_total = per * c->V;
    if (bytes_per_view) *bytes_per_view = per;
    return APDE_OK;
}

int apde_views_mark_maps(apde_context *c, int width, int height) {
    if (!c || c->V <= 0) return fail(APDE_ERR_STATE, "views_mark_maps: no scene");
    if (width <= 0 || height <= 0 || width > c->W || height > c->H) return fail(APDE_ERR_ARG, "views_mark_maps: bad size");
    for (auto &v : c->views) { v.mw = v.dw = width; v.mh = v.dh = height; v.has_conf = true; }
    return APDE_OK;
}

int apde_weak_vis_filter_range(apde_context *c, int first_view, int num_views, uint8_t *skip_weaks) {
    if (!c || c->V <= 0) return fail(APDE_ERR_STATE, "weak_vis_filter: no scene");
    if (first_view < 0 || num_views < 0 || first_view + num_views > c->V) return fail(APDE_ERR_ARG, "weak_vis_filter: bad view range");
    CU(cudaSetDevice(c->device));
    std::vector<FusionView> fv;
    int mw, mh;
    int rc = fusion_views(c, fv, &mw, &mh);
    if (rc) return rc;
    const size_t P = (size_t)mw * mh;
    if (!c->d_skip) CU(cudaMalloc(&c->d_skip, (size_t)c->V * c->W * c->H));
    CU(cudaMemsetAsync(c->d_skip + (size_t)first_view * P, 0, (size_t)num_views * P, c->stream));
    CU(fusion_weak_vis_filter(fv, mw, mh, first_view, num_views, c->d_skip, c->stream));
    c->launches += num_views;
    CU(cudaStreamSynchronize(c->stream));
    if (skip_weaks) CU(cudaMemcpy(skip_weaks, c->d_skip + (size_t)first_view * P, (size_t)num_views * P, cudaMemcpyDeviceToHost));
    return APDE_OK;
}

int apde_weak_vis_filter(apde_context *c, uint8_t *skip_weaks) {
    if (!c || c->V <= 0) return fail(APDE_ERR_STATE, "weak_vis_filter: no scene");
    return apde_weak_vis_filter_range(c, 0, c->V, skip_weaks);
}

int apde_fuse(apde_context *c, int use_weak_filter, float *xyz, float *bgr, int64_t max_points, int64_t *num_points) {
    return apde_fuse_variant(c, APDE_FUSE_DEFAULT, use_weak_filter, xyz, bgr, max_points, num_points);
}

int apde_fuse_variant(apde_context *c, int variant, int use_weak_filter, float *xyz, float *bgr, int64_t max_points,
                      int64_t *num_points) {
    if (!c || c->V <= 0 || !num_points) return fail(APDE_ERR_STATE, "fuse: bad argument");
    if (variant < APDE_FUSE_DEFAULT || variant > APDE_FUSE_TAT_A) return fail(APDE_ERR_ARG, "fuse: unknown variant %d", variant);
    CU(cudaSetDevice(c->device));
    std::vector<FusionView> fv;
    int mw, mh;
    int rc = fusion_views(c, fv, &mw, &mh);
    if (rc) return rc;
    const size_t P = (size_t)mw * mh;
    if (!c->d_skip) CU(cudaMalloc(&c->d_skip, (size_t)c->V * c->W * c->H));
    if (use_weak_filter == APDE_WEAK_FILTER_KEEP) {
        // skip maps already in the pool (filled by apde_weak_vis_filter_range on this or on peer GPUs)
    } else if (use_weak_filter) {
        if ((rc = apde_weak_vis_filter(c, nullptr))) return rc;
    } else {
        CU(cudaMemsetAsync(c->d_skip, 0, (size_t)c->V * P, c->stream));
    }
    uint64_t launches = 0;
    // a count-only call keeps the cloud on the host side of the context, so that a caller who sizes its buffers from the count
    // (apde_fuse_take_points) does not run the fusion a second time
    c->fused.xyz.clear(); c->fused.bgr.clear(); c->fused.valid = false;
    FusedPoints *keep = (!xyz && !bgr) ? &c->fused : nullptr;
    cudaError_t e = variant == APDE_FUSE_DEFAULT
                        ? fusion_run(fv, mw, mh, c->d_skip, xyz, bgr, max_points, num_points, c->stream, &launches, keep)
                        : fusion_run_tat(fv, mw, mh, variant, c->d_skip, xyz, bgr, max_points, num_points, c->stream, &launches, keep);
    c->launches += launches;
    if (e != cudaSuccess) { c->fused.xyz.clear(); c->fused.bgr.clear(); return fail(APDE_ERR_CUDA, "fusion: %s", cudaGetErrorString(e)); }
    if (keep) keep->valid = true;
    return APDE_OK;
}

int apde_fuse_take_points(apde_context *c, float *xyz, float *bgr, int64_t max_points, int64_t *num_points) {
    if (!c || !num_points) return fail(APDE_ERR_ARG, "fuse_take_points: null argument");
    if (!c->fused.valid) return fail(APDE_ERR_STATE, "fuse_take_points: no cloud kept (call apde_fuse / apde_fuse_variant / apde_fuse_collective with xyz = bgr = NULL first)");
    const int64_t n = (int64_t)(c->fused.xyz.size() / 3), k = std::max<int64_t>(0, std::min(n, max_points));
    if (xyz && k) memcpy(xyz, c->fused.xyz.data(), (size_t)k * 3 * sizeof(float));
    if (bgr && k) memcpy(bgr, c->fused.bgr.data(), (size_t)k * 3 * sizeof(float));
    *num_points = n;
    std::vector<float>().swap(c->fused.xyz);
    std::vector<float>().swap(c->fused.bgr);
    c->fused.valid = false;
    return APDE_OK;
}

}  // extern "C"
