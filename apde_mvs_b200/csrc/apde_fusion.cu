// apde_fusion.cu -- GPU fusion: WeakVisFilter (APD.cpp:962-1049) and RunFusion (APD.cpp:1051-1227).  sm_100a.
//
// Compiled WITHOUT --use_fast_math and with -fmad=false: the reference fusion is host code (IEEE arithmetic, no FMA
// contraction), and its accept / reject tests must come out the same.
//
// RunFusion is greedy and order dependent (views in order, pixels row-major; an accepted point masks the source pixels it
// used, APD.cpp:1149,1176,1209).  The GPU version reproduces that order exactly without a serial sweep:
//   phase A (parallel): every (pixel, neighbour) reprojection test that does not depend on masks written during this view;
//   phase B (parallel fixed point): "pixel p uses neighbour pixel q unless an accepted pixel p' < p already used q".
//     claim[q] = min accepted user of q.  The decision of p depends only on decisions of smaller pixels, so the system is
//     triangular and has exactly one solution -- the serial result -- which Jacobi iteration reaches in (max dependency
//     chain length) sweeps, a handful in practice;
//   phase C: commit masks, ordered compaction of the accepted points (prefix sum keeps the reference's point order).
#include "apde_fusion.h"

#include <cub/cub.cuh>

#include <climits>

#include <map>
#include <mutex>

namespace apde {

// Fusion runs once per scene but allocates ~1 GB of scratch each time; cudaMalloc / cudaFree of that size cost more than the
// kernels.  Blocks are therefore recycled through a small exact-size pool (per device), emptied by fusion_release_cache().
static std::multimap<std::pair<int, size_t>, void *> g_pool;
static std::map<void *, std::pair<int, size_t>> g_live;
static std::mutex g_pool_mutex;  // the contexts of a multi-GPU job (a host thread each) filter their views concurrently
template <typename T>
static cudaError_t pool_malloc(T **p, size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    const auto key = std::make_pair(dev, bytes);
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    auto it = g_pool.find(key);
    if (it != g_pool.end()) {
        *p = static_cast<T *>(it->second);
        g_pool.erase(it);
        g_live[*p] = key;
        return cudaSuccess;
    }
    void *q = nullptr;
    const cudaError_t e = cudaMalloc(&q, bytes);
    if (e == cudaSuccess) { *p = static_cast<T *>(q); g_live[q] = key; }
    return e;
}
static void pool_free(void *p) {
    if (!p) return;
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    auto it = g_live.find(p);
    if (it == g_live.end()) { cudaFree(p); return; }
    g_pool.emplace(it->second, p);
    g_live.erase(it);
}
void fusion_release_cache() {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    for (auto &kv : g_pool) cudaFree(kv.second);
    g_pool.clear();
}

struct FCam {
    float K[9], R[9], t[3], c[3];
    const float *depth;
    const float *normal;
    const uint8_t *weak;
    const uint8_t *conf;
    const uint8_t *bgr;
};

// Get3DPointonWorld, APD.cpp:866-889 (camera centre recomputed in float)
__device__ __forceinline__ float3 point_on_world(int x, int y, float depth, const FCam &cam) {
    const float px = depth * (x - cam.K[2]) / cam.K[0];
    const float py = depth * (y - cam.K[5]) / cam.K[4];
    const float pz = depth;
    const float tx = cam.R[0] * px + cam.R[3] * py + cam.R[6] * pz;
    const float ty = cam.R[1] * px + cam.R[4] * py + cam.R[7] * pz;
    const float tz = cam.R[2] * px + cam.R[5] * py + cam.R[8] * pz;
    const float cx = -(cam.R[0] * cam.t[0] + cam.R[3] * cam.t[1] + cam.R[6] * cam.t[2]);
    const float cy = -(cam.R[1] * cam.t[0] + cam.R[4] * cam.t[1] + cam.R[7] * cam.t[2]);
    const float cz = -(cam.R[2] * cam.t[0] + cam.R[5] * cam.t[1] + cam.R[8] * cam.t[2]);
    return make_float3(tx + cx, ty + cy, tz + cz);
}
// ProjectCamera, APD.cpp:891-900
__device__ __forceinline__ void project_camera(float3 X, const FCam &cam, float2 &pt, float &depth) {
    const float tx = cam.R[0] * X.x + cam.R[1] * X.y + cam.R[2] * X.z + cam.t[0];
    const float ty = cam.R[3] * X.x + cam.R[4] * X.y + cam.R[5] * X.z + cam.t[1];
    const float tz = cam.R[6] * X.x + cam.R[7] * X.y + cam.R[8] * X.z + cam.t[2];
    depth = cam.K[6] * tx + cam.K[7] * ty + cam.K[8] * tz;
    pt.x = (cam.K[0] * tx + cam.K[1] * ty + cam.K[2] * tz) / depth;
    pt.y = (cam.K[3] * tx + cam.K[4] * ty + cam.K[5] * tz) / depth;
}
// GetAngle, APD.cpp:902-910 (cv::norm is double)
__device__ __forceinline__ float get_angle(float ax, float ay, float az, float bx, float by, float bz) {
    const float dot = ax * bx + ay * by + az * bz;
    const double na = sqrt((double)ax * ax + (double)ay * ay + (double)az * az);
    const double nb = sqrt((double)bx * bx + (double)by * by + (double)bz * bz);
    const float ang = acosf((float)(dot / (na * nb)));
    return (ang != ang) ? 0.0f : ang;
}
// confidences[..].at<float>(r, c) on a CV_8UC1 map (APD.cpp:1010-1011): four raw bytes at r*W + 4c, zero past the end
__device__ __forceinline__ float conf_as_float(const uint8_t *conf, int P, int W, int r, int c) {
    const int off = r * W + 4 * c;
    uint32_t bits = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) if (off + i < P) bits |= (uint32_t)conf[off + i] << (8 * i);
    return __uint_as_float(bits);
}

__global__ void __launch_bounds__(128) k_weak_vis(const FCam *__restrict__ cams, int V, int ref, int W, int H,
                                                  uint8_t *__restrict__ skip) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int P = W * H;
    if (idx >= P) return;
    const FCam &rc = cams[ref];
    if (rc.weak[idx] != APDE_WEAK) return;
    const int r = idx / W, c = idx % W;
    const float ref_depth = rc.depth[idx];
    const float3 X = point_on_world(c, r, ref_depth, rc);
    int so = 0, wo = 0;
    for (int s = 0; s < V; ++s) {
        if (s == ref) continue;
        const FCam &sc = cams[s];
        float ang = get_angle(rc.c[0] - X.x, rc.c[1] - X.y, rc.c[2] - X.z, sc.c[0] - X.x, sc.c[1] - X.y, sc.c[2] - X.z);
        ang = (float)((double)(ang * 180.0f) / 3.14159265358979323846);
        if (ang > 80.0f) continue;
        float2 pt;
        float pd;
        project_camera(X, sc, pt, pd);
        if (pd <= 0.0f) continue;
        const int sr = (int)(pt.y + 0.5f), scn = (int)(pt.x + 0.5f);
        if (scn >= 0 && scn < W && sr >= 0 && sr < H) {
            const float sd = sc.depth[sr * W + scn];
            const uint8_t wk = sc.weak[sr * W + scn];
            if (wk == APDE_STRONG) {
                if (pd < sd - 0.01f * sd) so++;
            } else if (wk == APDE_WEAK) {
                if (conf_as_float(sc.conf, P, W, sr, scn) < conf_as_float(rc.conf, P, W, r, c)) {
                    if (pd < sd - 0.01f * sd) wo++;
                }
            }
        }
    }
    if (so >= 2 || wo >= 4) skip[(size_t)ref * P + idx] = 1;
}

// phase A: mask-independent tests of pixel idx of view `ref` against its neighbours
__global__ void __launch_bounds__(128) k_fuse_candidates(const FCam *__restrict__ cams, const int *__restrict__ nbr, int N,
                                                         int ref, int W, int H, const uint8_t *__restrict__ masks,
                                                         const uint8_t *__restrict__ skip, int *__restrict__ cq,
                                                         float *__restrict__ ce, uint8_t *__restrict__ active) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int P = W * H;
    if (idx >= P) return;
    const FCam &rc = cams[ref];
    uint8_t act = 0;
    const float ref_depth = rc.depth[idx];
    if (masks[(size_t)ref * P + idx] != 1 && skip[(size_t)ref * P + idx] != 1 && !(ref_depth <= 0.0f)) {
        act = 1;
        const int r = idx / W, c = idx % W;
        const float3 X = point_on_world(c, r, ref_depth, rc);
        const float nx = rc.normal[3 * idx], ny = rc.normal[3 * idx + 1], nz = rc.normal[3 * idx + 2];
        for (int j = 0; j < N; ++j) {
            const int s = nbr[j];
            const FCam &sc = cams[s];
            int q = -1;
            float e = 0.0f;
            float2 pt;
            float pd;
            project_camera(X, sc, pt, pd);
            const int sr = (int)(pt.y + 0.5f), scn = (int)(pt.x + 0.5f);
            if (scn >= 0 && scn < W && sr >= 0 && sr < H) {
                const int sp = sr * W + scn;
                const float sd = sc.depth[sp];
                if (masks[(size_t)s * P + sp] != 1 && !(sd <= 0.0f)) {
                    const float3 tX = point_on_world(scn, sr, sd, sc);
                    float2 tp;
                    project_camera(tX, rc, tp, pd);
                    const float re = (float)sqrt(pow((double)(c - tp.x), 2.0) + pow((double)(r - tp.y), 2.0));
                    const float rdd = fabsf(pd - ref_depth) / ref_depth;
                    const float ang = get_angle(nx, ny, nz, sc.normal[3 * sp], sc.normal[3 * sp + 1], sc.normal[3 * sp + 2]);
                    if (re < 2.0f && rdd < 0.01f && ang < 0.174533f) {
                        q = sp;
                        e = (float)exp((double)-(re + 200 * rdd + ang * 10));
                    }
                }
            }
            cq[(size_t)j * P + idx] = q;
            ce[(size_t)j * P + idx] = e;
        }
    }
    active[idx] = act;
}

// phase B: one Jacobi sweep.  used[p] = bit mask of neighbours consumed by an accepted p (0 = rejected).
__global__ void __launch_bounds__(128) k_fuse_sweep(const int *__restrict__ slot, int N, int P, const uint8_t *__restrict__ active,
                                                    const uint8_t *__restrict__ weak, const int *__restrict__ cq,
                                                    const float *__restrict__ ce, const int *__restrict__ claim_prev,
                                                    int *__restrict__ claim_next, uint32_t *__restrict__ used,
                                                    int *__restrict__ changed) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P || !active[idx]) return;
    int num = 0;
    float dyn = 0.0f;
    uint32_t bits = 0;
    for (int j = 0; j < N; ++j) {
        const int q = cq[(size_t)j * P + idx];
        if (q < 0) continue;
        if (claim_prev[(size_t)slot[j] * P + q] < idx) continue;  // masked by an earlier accepted pixel of this view
        bits |= 1u << j;
        dyn += ce[(size_t)j * P + idx];
        num++;
    }
    const float factor = (weak[idx] == APDE_WEAK) ? 0.45f : 0.3f;
    if (!(num >= 1 && dyn > factor * num)) bits = 0;
    if (bits != used[idx]) { used[idx] = bits; *changed = 1; }
    for (int j = 0; j < N; ++j)
        if ((bits >> j) & 1u) atomicMin(&claim_next[(size_t)slot[j] * P + cq[(size_t)j * P + idx]], idx);
}

// phase C: commit masks, flag accepted pixels
__global__ void __launch_bounds__(128) k_fuse_commit(const int *__restrict__ nbr, int N, int P, const int *__restrict__ cq,
                                                     const uint32_t *__restrict__ used, uint8_t *__restrict__ masks,
                                                     int *__restrict__ flags) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const uint32_t bits = used[idx];
    flags[idx] = bits ? 1 : 0;
    for (int j = 0; j < N; ++j)
        if ((bits >> j) & 1u) masks[(size_t)nbr[j] * P + cq[(size_t)j * P + idx]] = 1;
}

__global__ void __launch_bounds__(128) k_fuse_emit(const FCam *__restrict__ cams, const int *__restrict__ nbr, int N, int ref,
                                                   int W, int H, const int *__restrict__ cq, const uint32_t *__restrict__ used,
                                                   const int *__restrict__ offs, float *__restrict__ xyz,
                                                   float *__restrict__ bgr) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int P = W * H;
    if (idx >= P) return;
    const uint32_t bits = used[idx];
    if (!bits) return;
    const FCam &rc = cams[ref];
    const int o = offs[idx];
    const float3 X = point_on_world(idx % W, idx / W, rc.depth[idx], rc);
    xyz[3 * o] = X.x; xyz[3 * o + 1] = X.y; xyz[3 * o + 2] = X.z;
    float col[3] = {0.0f, 0.0f, 0.0f};
    int n = 0;
    if (rc.bgr) {
        for (int k = 0; k < 3; ++k) col[k] = (float)rc.bgr[3 * idx + k];
        for (int j = 0; j < N; ++j) {
            if (!((bits >> j) & 1u)) continue;
            const uint8_t *sb = cams[nbr[j]].bgr;
            const int q = cq[(size_t)j * P + idx];
            if (sb) for (int k = 0; k < 3; ++k) col[k] += (float)sb[3 * q + k];
            n++;
        }
        for (int k = 0; k < 3; ++k) col[k] /= (n + 1);
    }
    bgr[3 * o] = col[0]; bgr[3 * o + 1] = col[1]; bgr[3 * o + 2] = col[2];
}

__global__ void k_fill_int(int *p, int v, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

static cudaError_t upload_cams(const std::vector<FusionView> &views, FCam **d_cams) {
    std::vector<FCam> hc(views.size());
    for (size_t v = 0; v < views.size(); ++v) {
        const apde_camera &c = views[v].cam;
        for (int i = 0; i < 9; ++i) { hc[v].K[i] = c.K[i]; hc[v].R[i] = c.R[i]; }
        for (int i = 0; i < 3; ++i) { hc[v].t[i] = c.t[i]; hc[v].c[i] = c.c[i]; }
        hc[v].depth = views[v].depth; hc[v].normal = views[v].normal; hc[v].weak = views[v].weak; hc[v].conf = views[v].conf;
        hc[v].bgr = views[v].bgr;  // the caller passes colours only when they have the maps' resolution
    }
    cudaError_t e = pool_malloc(d_cams, hc.size() * sizeof(FCam));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*d_cams, hc.data(), hc.size() * sizeof(FCam), cudaMemcpyHostToDevice);
}

cudaError_t fusion_weak_vis_filter(const std::vector<FusionView> &views, int w, int h, int first_view, int num_views, uint8_t *skip,
                                   cudaStream_t st) {
    FCam *d_cams = nullptr;
    cudaError_t e = upload_cams(views, &d_cams);
    if (e != cudaSuccess) return e;
    const int V = (int)views.size(), P = w * h;
    for (int ref = first_view; ref < first_view + num_views; ++ref) k_weak_vis<<<(P + 127) / 128, 128, 0, st>>>(d_cams, V, ref, w, h, skip);
    e = cudaStreamSynchronize(st);
    pool_free(d_cams);
    return e != cudaSuccess ? e : cudaGetLastError();
}

// the points of one reference view leave the device: into the caller's buffers (as far as they reach) and / or appended to `keep`
static cudaError_t deliver_points(const float *d_xyz, const float *d_bgr, int64_t n, int64_t total, float *xyz, float *bgr,
                                  int64_t max_points, FusedPoints *keep) {
    const int64_t room = std::max<int64_t>(0, std::min<int64_t>(n, max_points - total));
    cudaError_t e = cudaSuccess;
    if (room > 0 && xyz && (e = cudaMemcpy(xyz + 3 * total, d_xyz, (size_t)room * 3 * sizeof(float), cudaMemcpyDeviceToHost)) != cudaSuccess) return e;
    if (room > 0 && bgr && (e = cudaMemcpy(bgr + 3 * total, d_bgr, (size_t)room * 3 * sizeof(float), cudaMemcpyDeviceToHost)) != cudaSuccess) return e;
    if (keep && n > 0) {
        keep->xyz.resize((size_t)(total + n) * 3);
        keep->bgr.resize((size_t)(total + n) * 3);
        if ((e = cudaMemcpy(keep->xyz.data() + 3 * total, d_xyz, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost)) != cudaSuccess) return e;
        if ((e = cudaMemcpy(keep->bgr.data() + 3 * total, d_bgr, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost)) != cudaSuccess) return e;
    }
    return e;
}

#define FCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { err = e_; goto done; } } while (0)

cudaError_t fusion_run(const std::vector<FusionView> &views, int w, int h, const uint8_t *skip, float *xyz, float *bgr,
                       int64_t max_points, int64_t *num_points, cudaStream_t st, uint64_t *launches, FusedPoints *keep) {
    const int V = (int)views.size(), P = w * h;
    cudaError_t err = cudaSuccess;
    FCam *d_cams = nullptr;
    uint8_t *d_masks = nullptr, *d_active = nullptr;
    int *d_nbr = nullptr, *d_slot = nullptr, *d_cq = nullptr, *d_claim[2] = {nullptr, nullptr}, *d_flags = nullptr, *d_offs = nullptr,
        *d_changed = nullptr;
    float *d_ce = nullptr, *d_xyz = nullptr, *d_bgr = nullptr;
    uint32_t *d_used = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    int64_t total = 0;
    int maxN = 1;
    for (auto &v : views) maxN = std::max(maxN, (int)v.src.size());
    const int blocks = (P + 127) / 128;
    FCU(upload_cams(views, &d_cams));
    FCU(pool_malloc(&d_masks, (size_t)V * P));
    FCU(cudaMemsetAsync(d_masks, 0, (size_t)V * P, st));
    FCU(pool_malloc(&d_active, P));
    FCU(pool_malloc(&d_nbr, maxN * sizeof(int)));
    FCU(pool_malloc(&d_slot, maxN * sizeof(int)));
    FCU(pool_malloc(&d_cq, (size_t)maxN * P * sizeof(int)));
    FCU(pool_malloc(&d_ce, (size_t)maxN * P * sizeof(float)));
    FCU(pool_malloc(&d_claim[0], (size_t)maxN * P * sizeof(int)));
    FCU(pool_malloc(&d_claim[1], (size_t)maxN * P * sizeof(int)));
    FCU(pool_malloc(&d_used, (size_t)P * sizeof(uint32_t)));
    FCU(pool_malloc(&d_flags, (size_t)P * sizeof(int)));
    FCU(pool_malloc(&d_offs, (size_t)P * sizeof(int)));
    FCU(pool_malloc(&d_changed, sizeof(int)));
    FCU(pool_malloc(&d_xyz, (size_t)P * 3 * sizeof(float)));
    FCU(pool_malloc(&d_bgr, (size_t)P * 3 * sizeof(float)));
    FCU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_flags, d_offs, P, st));
    FCU(pool_malloc(&d_tmp, tmp_bytes));

    for (int ref = 0; ref < V; ++ref) {
        const int N = (int)views[ref].src.size();
        if (N == 0) continue;
        std::vector<int> slot(N);
        for (int j = 0; j < N; ++j) {  // duplicate neighbour ids share one claim plane, as they share masks[src] in the reference
            slot[j] = j;
            for (int k = 0; k < j; ++k) if (views[ref].src[k] == views[ref].src[j]) { slot[j] = k; break; }
        }
        FCU(cudaMemcpyAsync(d_nbr, views[ref].src.data(), N * sizeof(int), cudaMemcpyHostToDevice, st));
        FCU(cudaMemcpyAsync(d_slot, slot.data(), N * sizeof(int), cudaMemcpyHostToDevice, st));
        k_fuse_candidates<<<blocks, 128, 0, st>>>(d_cams, d_nbr, N, ref, w, h, d_masks, skip, d_cq, d_ce, d_active);
        FCU(cudaMemsetAsync(d_used, 0, (size_t)P * sizeof(uint32_t), st));
        const size_t nclaim = (size_t)N * P;
        k_fill_int<<<(unsigned)((nclaim + 255) / 256), 256, 0, st>>>(d_claim[0], INT_MAX, nclaim);
        if (launches) *launches += 2;
        int cur = 0;
        for (int sweep = 0; sweep < 100000; ++sweep) {
            k_fill_int<<<(unsigned)((nclaim + 255) / 256), 256, 0, st>>>(d_claim[cur ^ 1], INT_MAX, nclaim);
            FCU(cudaMemsetAsync(d_changed, 0, sizeof(int), st));
            k_fuse_sweep<<<blocks, 128, 0, st>>>(d_slot, N, P, d_active, views[ref].weak, d_cq, d_ce, d_claim[cur], d_claim[cur ^ 1],
                                                 d_used, d_changed);
            if (launches) *launches += 2;
            int changed = 0;
            FCU(cudaMemcpyAsync(&changed, d_changed, sizeof(int), cudaMemcpyDeviceToHost, st));
            FCU(cudaStreamSynchronize(st));
            cur ^= 1;
            if (!changed) break;
        }
        k_fuse_commit<<<blocks, 128, 0, st>>>(d_nbr, N, P, d_cq, d_used, d_masks, d_flags);
        FCU(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_flags, d_offs, P, st));
        k_fuse_emit<<<blocks, 128, 0, st>>>(d_cams, d_nbr, N, ref, w, h, d_cq, d_used, d_offs, d_xyz, d_bgr);
        if (launches) *launches += 3;
        int last_off = 0, last_flag = 0;
        FCU(cudaMemcpyAsync(&last_off, d_offs + P - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        FCU(cudaMemcpyAsync(&last_flag, d_flags + P - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        FCU(cudaStreamSynchronize(st));
        const int64_t n = (int64_t)last_off + last_flag;
        FCU(deliver_points(d_xyz, d_bgr, n, total, xyz, bgr, max_points, keep));
        total += n;
    }
    *num_points = total;
done:
    pool_free(d_cams); pool_free(d_masks); pool_free(d_active); pool_free(d_nbr); pool_free(d_slot); pool_free(d_cq); pool_free(d_ce);
    pool_free(d_claim[0]); pool_free(d_claim[1]); pool_free(d_used); pool_free(d_flags); pool_free(d_offs); pool_free(d_changed);
    pool_free(d_xyz); pool_free(d_bgr); pool_free(d_tmp);
    return err;
}

// ------------------------------------------------------------------------------------------------ Tanks-and-Temples variants
// RunFusion_TAT_I (variant 1, APD.cpp:1229-1431) and RunFusion_TAT_A (variant 2, APD.cpp:1433-1608).  masks[] of a view
// are written only while that view is the reference and read only through its neighbours, so one view's pixels are
// independent -- except for the reference's per-VIEW diff[] array: a neighbour that is out of bounds / masked / without
// depth at pixel p keeps the measurements of the last pixel q < p (row-major) that did update it.  That carry-over is a
// forward fill: last_j(p) = max{q <= p : neighbour j measured at q}, computed for every neighbour with one inclusive
// max-scan over the pixel axis; the decision kernel then reads its measurements at last_j(p).
struct TatMeasure { float dist, depth, angle; int q; };

__global__ void __launch_bounds__(128) k_tat_measure(const FCam *__restrict__ cams, const int *__restrict__ nbr, int N, int ref,
                                                     int W, int H, const uint8_t *__restrict__ masks,
                                                     const uint8_t *__restrict__ skip, TatMeasure *__restrict__ meas,
                                                     int *__restrict__ last, uint8_t *__restrict__ active) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int P = W * H;
    if (idx >= P) return;
    const FCam &rc = cams[ref];
    const float ref_depth = rc.depth[idx];
    const bool act = skip[(size_t)ref * P + idx] != 1 && !(ref_depth <= 0.0f);
    active[idx] = act ? 1 : 0;
    const int r = idx / W, c = idx % W;
    float3 X = make_float3(0.f, 0.f, 0.f);
    float nx = 0.f, ny = 0.f, nz = 0.f;
    if (act) {
        X = point_on_world(c, r, ref_depth, rc);
        nx = rc.normal[3 * idx]; ny = rc.normal[3 * idx + 1]; nz = rc.normal[3 * idx + 2];
    }
    for (int j = 0; j < N; ++j) {
        int l = -1;
        if (act) {
            const FCam &sc = cams[nbr[j]];
            float2 pt;
            float pd;
            project_camera(X, sc, pt, pd);
            const int sr = (int)(pt.y + 0.5f), scn = (int)(pt.x + 0.5f);
            if (scn >= 0 && scn < W && sr >= 0 && sr < H) {
                const int sp = sr * W + scn;
                const float sd = sc.depth[sp];
                if (masks[(size_t)nbr[j] * P + sp] != 1 && !(sd <= 0.0f)) {
                    const float3 tX = point_on_world(scn, sr, sd, sc);
                    float2 tp;
                    project_camera(tX, rc, tp, pd);
                    TatMeasure m;
                    m.dist = (float)sqrt(pow((double)(c - tp.x), 2.0) + pow((double)(r - tp.y), 2.0));
                    m.depth = fabsf(pd - ref_depth) / ref_depth;
                    m.angle = get_angle(nx, ny, nz, sc.normal[3 * sp], sc.normal[3 * sp + 1], sc.normal[3 * sp + 2]);
                    m.q = sp;
                    meas[(size_t)j * P + idx] = m;
                    l = idx;
                }
            }
        }
        last[(size_t)j * P + idx] = l;
    }
}

struct MaxInt {
    __device__ __forceinline__ int operator()(int a, int b) const { return a > b ? a : b; }
};

__global__ void __launch_bounds__(128) k_tat_decide(int N, int P, int variant, const uint8_t *__restrict__ active,
                                                    const TatMeasure *__restrict__ meas, const int *__restrict__ last,
                                                    uint32_t *__restrict__ used, int *__restrict__ flags,
                                                    uint8_t *__restrict__ mask_ref) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    uint32_t bits = 0;
    int ok = 0;
    if (active[idx]) {
        const float dist_base = 0.25f;
        const float depth_base = variant == 1 ? 1.0f / 3500.0f : 1.0f / 3000.0f;
        const float angle_base = 0.06981317007977318f, angle_grad = 0.05235987755982988f;
        for (int k = 2; k <= N && !ok; ++k) {
            int count = 0;
            bits = 0;
            for (int j = 0; j < N; ++j) {
                const int l = last[(size_t)j * P + idx];
                if (l < 0) continue;  // CostData() defaults: FLT_MAX never passes
                const TatMeasure m = meas[(size_t)j * P + l];
                bool pass = m.dist < k * dist_base && m.depth < k * depth_base;
                if (variant == 1) pass = pass && m.angle < (k * angle_grad + angle_base);
                if (pass) { count++; bits |= 1u << j; }
            }
            if (count >= k) ok = 1;
        }
    }
    used[idx] = ok ? bits : 0u;
    flags[idx] = ok;
    if (ok) mask_ref[idx] = 1;
}

__global__ void __launch_bounds__(128) k_tat_emit(const FCam *__restrict__ cams, const int *__restrict__ nbr, int N, int ref,
                                                  int W, int H, int variant, const TatMeasure *__restrict__ meas,
                                                  const int *__restrict__ last, const uint32_t *__restrict__ used,
                                                  const int *__restrict__ flags, const int *__restrict__ offs,
                                                  float *__restrict__ xyz, float *__restrict__ bgr) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int P = W * H;
    if (idx >= P || !flags[idx]) return;
    const FCam &rc = cams[ref];
    const int o = offs[idx];
    const float3 X = point_on_world(idx % W, idx / W, rc.depth[idx], rc);
    xyz[3 * o] = X.x; xyz[3 * o + 1] = X.y; xyz[3 * o + 2] = X.z;
    float col[3] = {0.0f, 0.0f, 0.0f};
    if (rc.bgr) {
        for (int k = 0; k < 3; ++k) col[k] = (float)rc.bgr[3 * idx + k];
        if (variant == 1) {
            const uint32_t bits = used[idx];
            int n = 0;
            for (int j = 0; j < N; ++j) {
                if (!((bits >> j) & 1u)) continue;
                const uint8_t *sb = cams[nbr[j]].bgr;
                const int q = meas[(size_t)j * P + last[(size_t)j * P + idx]].q;  // possibly another pixel's coordinates (quirk)
                if (sb) for (int k = 0; k < 3; ++k) col[k] += (float)sb[3 * q + k];
                n++;
            }
            for (int k = 0; k < 3; ++k) col[k] /= (n + 1.0f);
        }
    }
    bgr[3 * o] = col[0]; bgr[3 * o + 1] = col[1]; bgr[3 * o + 2] = col[2];
}

cudaError_t fusion_run_tat(const std::vector<FusionView> &views, int w, int h, int variant, const uint8_t *skip, float *xyz,
                           float *bgr, int64_t max_points, int64_t *num_points, cudaStream_t st, uint64_t *launches, FusedPoints *keep) {
    const int V = (int)views.size(), P = w * h;
    cudaError_t err = cudaSuccess;
    FCam *d_cams = nullptr;
    uint8_t *d_masks = nullptr, *d_active = nullptr;
    int *d_nbr = nullptr, *d_last = nullptr, *d_flags = nullptr, *d_offs = nullptr;
    TatMeasure *d_meas = nullptr;
    float *d_xyz = nullptr, *d_bgr = nullptr;
    uint32_t *d_used = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_a = 0, tmp_b = 0;
    int64_t total = 0;
    int maxN = 1;
    for (auto &v : views) maxN = std::max(maxN, (int)v.src.size());
    const int blocks = (P + 127) / 128;
    FCU(upload_cams(views, &d_cams));
    FCU(pool_malloc(&d_masks, (size_t)V * P));
    FCU(cudaMemsetAsync(d_masks, 0, (size_t)V * P, st));
    FCU(pool_malloc(&d_active, P));
    FCU(pool_malloc(&d_nbr, maxN * sizeof(int)));
    FCU(pool_malloc(&d_last, (size_t)maxN * P * sizeof(int)));
    FCU(pool_malloc(&d_meas, (size_t)maxN * P * sizeof(TatMeasure)));
    FCU(pool_malloc(&d_used, (size_t)P * sizeof(uint32_t)));
    FCU(pool_malloc(&d_flags, (size_t)P * sizeof(int)));
    FCU(pool_malloc(&d_offs, (size_t)P * sizeof(int)));
    FCU(pool_malloc(&d_xyz, (size_t)P * 3 * sizeof(float)));
    FCU(pool_malloc(&d_bgr, (size_t)P * 3 * sizeof(float)));
    FCU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_a, d_flags, d_offs, P, st));
    FCU(cub::DeviceScan::InclusiveScan(nullptr, tmp_b, d_last, d_last, MaxInt(), P, st));
    FCU(pool_malloc(&d_tmp, std::max(tmp_a, tmp_b)));
    tmp_a = tmp_b = std::max(tmp_a, tmp_b);

    for (int ref = 0; ref < V; ++ref) {
        const int N = (int)views[ref].src.size();
        if (N < 2) continue;  // the k loop starts at 2 (APD.cpp:1382)
        FCU(cudaMemcpyAsync(d_nbr, views[ref].src.data(), N * sizeof(int), cudaMemcpyHostToDevice, st));
        k_tat_measure<<<blocks, 128, 0, st>>>(d_cams, d_nbr, N, ref, w, h, d_masks, skip, d_meas, d_last, d_active);
        for (int j = 0; j < N; ++j)
            FCU(cub::DeviceScan::InclusiveScan(d_tmp, tmp_b, d_last + (size_t)j * P, d_last + (size_t)j * P, MaxInt(), P, st));
        k_tat_decide<<<blocks, 128, 0, st>>>(N, P, variant, d_active, d_meas, d_last, d_used, d_flags, d_masks + (size_t)ref * P);
        FCU(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_a, d_flags, d_offs, P, st));
        k_tat_emit<<<blocks, 128, 0, st>>>(d_cams, d_nbr, N, ref, w, h, variant, d_meas, d_last, d_used, d_flags, d_offs, d_xyz,
                                           d_bgr);
        if (launches) *launches += 4 + N;
        int last_off = 0, last_flag = 0;
        FCU(cudaMemcpyAsync(&last_off, d_offs + P - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        FCU(cudaMemcpyAsync(&last_flag, d_flags + P - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        FCU(cudaStreamSynchronize(st));
        const int64_t n = (int64_t)last_off + last_flag;
        FCU(deliver_points(d_xyz, d_bgr, n, total, xyz, bgr, max_points, keep));
        total += n;
    }
    *num_points = total;
done:
    pool_free(d_cams); pool_free(d_masks); pool_free(d_active); pool_free(d_nbr); pool_free(d_last); pool_free(d_meas); pool_free(d_used);
    pool_free(d_flags); pool_free(d_offs); pool_free(d_xyz); pool_free(d_bgr); pool_free(d_tmp);
    return err;
}

}  // namespace apde
