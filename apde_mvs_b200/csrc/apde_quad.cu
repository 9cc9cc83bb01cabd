// apde_quad.cu -- quad-cooperative cost evaluation (sm_100a): the B200 mapping of the hot kernels.
//
// Measured on B200 (tools/diag_texpattern.py, profiles/r01_*): with one thread per evaluation the texture unit delivers
// its 4 filtered samples/clk/SM only while the 32 lanes of a warp sample coherent patches; as soon as neighbouring pixels
// hold different plane hypotheses (random initialisation, refinement candidates, occlusion edges) the L1TEX data-stage
// needs several wavefronts per request and throughput falls to 32-48 % of peak.  Here FOUR lanes (a quad == one texture
// request group) evaluate ONE (pixel, hypothesis, view): each lane gathers 9 of the 36 samples, arranged so that every
// instruction of the quad touches a 2x2 block of neighbouring sample positions.  That keeps 98 % of the texture rate for
// any scatter of the hypotheses.
//
// The 36 samples are then accumulated in the REFERENCE ORDER (i outer, j inner, APD.cu:629-643) by a width-4 shuffle
// chain executed redundantly by the four lanes: fp32 addition is not associative, and on weak texture
// var = E[x^2] - E[x]^2 is rounding-noise dominated, so the order is part of the reference's formulation.
#include <cfloat>

#include "apde_common.cuh"
#include "apde_kernels.h"

namespace apde {

// ------------------------------------------------------------------------------------------------ quad primitives
struct QuadCtx {
    int q;           // lane within the quad
    unsigned qmask;  // the quad's four lanes
};

// reference patch of the quad's pixel, all 36 texels on every lane (needed by the redundant sum_rs chain)
struct QRef {
    float r[kPatch];
    float mean, var;
};

__device__ __forceinline__ void q_load_ref(const PassK &K, int px, int py, QRef &rp) {
    float s = 0.0f, ss = 0.0f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const float v = K.tex_unorm > 0.0f ? fetch<true>(K, px + (2 * i - 5) + 0.5f, py + (2 * j - 5) + 0.5f, K.ref_layer) : fetch<false>(K, px + (2 * i - 5) + 0.5f, py + (2 * j - 5) + 0.5f, K.ref_layer);
            rp.r[i * 6 + j] = v;
            s += v;
            ss = fmaf(v, v, ss);
        }
    }
    const float inv = 1.0f / 36.0f;
    rp.mean = inv * s;
    rp.var = fmaf(inv, ss, -__fmul_rn(rp.mean, rp.mean));
}

// one NCC-Old evaluation by the four lanes of a quad; every lane returns the same cost.  APD.cu:596-663.
__device__ __forceinline__ float q_ncc_old(const PassK &K, const ViewK &vk, int px, int py, float3 m, const QRef &rp,
                                           const QuadCtx &qc) {
    const Homog Hm = make_homography(vk, m);
    const float *h = Hm.h;
    const float fxp = (float)px, fyp = (float)py;
    {
        const float Z = h[6] * fxp + h[7] * fyp + h[8];
        const float iz = rcp_approx(Z);
        const float ptx = (h[0] * fxp + h[1] * fyp + h[2]) * iz;
        const float pty = (h[3] * fxp + h[4] * fyp + h[5]) * iz;
        if (ptx >= (float)K.W || ptx < 0.0f || pty >= (float)K.H || pty < 0.0f) return 2.0f;  // quad-uniform
    }
    const float g0 = fmaf(0.5f, h[6], h[0]), g1 = fmaf(0.5f, h[7], h[1]), g2 = fmaf(0.5f, h[8], h[2]);
    const float g3 = fmaf(0.5f, h[6], h[3]), g4 = fmaf(0.5f, h[7], h[4]), g5 = fmaf(0.5f, h[8], h[5]);
    const int layer = vk.layer;
    const int qi = qc.q & 1, qj = qc.q >> 1;
    float s[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int i = 2 * (t / 3) + qi, j = 2 * (t % 3) + qj;
        const float xi = (float)(px + 2 * i - 5), yj = (float)(py + 2 * j - 5);
        const float bx = fmaf(g0, xi, g2), by = fmaf(g3, xi, g5), bz = fmaf(h[6], xi, h[8]);
        const float X = fmaf(g1, yj, bx), Y = fmaf(g4, yj, by), Z = fmaf(h[7], yj, bz);
        const float iz = rcp_approx(Z);
        s[t] = K.tex_unorm > 0.0f ? fetch<true>(K, X * iz, Y * iz, layer) : fetch<false>(K, X * iz, Y * iz, layer);
    }
    float sum_s = 0.0f, sum_ss = 0.0f, sum_rs = 0.0f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const float v = __shfl_sync(qc.qmask, s[(i >> 1) * 3 + (j >> 1)], (i & 1) | ((j & 1) << 1), 4);
            sum_s += v;
            sum_ss = fmaf(v, v, sum_ss);
            sum_rs = fmaf(rp.r[i * 6 + j], v, sum_rs);
        }
    }
    const float inv = 1.0f / 36.0f;
    const float mean_s = inv * sum_s, e_rs = inv * sum_rs;
    const float var_s = fmaf(inv, sum_ss, -__fmul_rn(mean_s, mean_s));
    if (rp.var < 1e-5f || var_s < 1e-5f) return 2.0f;
    const float covar = fmaf(-rp.mean, mean_s, e_rs);
    return fmaxf(0.0f, fminf(2.0f, fmaf(-covar, rsqrtf(rp.var * var_s), 1.0f)));
}

// pixel <-> tile-lane mappings shared with the thread-per-pixel kernels (apde_common.cuh)
__device__ __forceinline__ void half_tile_pixel(int tile, int tiles_x, int color, int bit, int &px, int &py) {
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    py = ty * 8 + (bit >> 2);
    px = tx * 8 + 2 * (bit & 3) + ((py + color) & 1);
}
__device__ __forceinline__ void full_tile_pixel(int tile, int tiles_x, int bit, int &px, int &py) {
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    py = ty * 4 + (bit >> 3);
    px = tx * 8 + (bit & 7);
}

__device__ __forceinline__ void q_count(const PassK &K, const QuadCtx &qc, unsigned n_old, unsigned n_new, unsigned n_geom) {
    if (qc.q == 0) {
        if (n_old) atomicAdd(&K.counters[0], (unsigned long long)n_old);
        if (n_new) atomicAdd(&K.counters[1], (unsigned long long)n_new);
        if (n_geom) atomicAdd(&K.counters[2], (unsigned long long)n_geom);
    }
}

// ------------------------------------------------------------------------------------------------ K6 strong propagation
// Black/RedPixelUpdateStrong -> CheckerboardPropagationStrong -> PlaneHypothesisRefinementStrong (APD.cu:1654-1692,
// 1098-1440, 950-1006), one quad per pixel, eight pixels per warp round.  Same phases, arithmetic and tie rules as
// k_prop_strong (apde_kernels.cu); scalar steps are executed redundantly by the four lanes.
__global__ void __launch_bounds__(128) k_qprop_strong(const __grid_constant__ PassK K, int iter, int color, int tiles_x,
                                                      int ylimit) {
    extern __shared__ float smem[];
    const ViewK *s_vk = stage_views(K, smem);
    const int W = K.W, N = K.N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    QuadCtx qc;
    qc.q = lane & 3;
    qc.qmask = 0xFu << (lane & ~3);
    const int e = lane >> 2;
    float *sc = smem + views_smem_floats(N) + (size_t)(warp * 8 + e) * 9 * N;  // quad-private: cost[8][N], prob[N]
    float *sp = sc + 8 * N;
    const int tile = blockIdx.x * (blockDim.x >> 5) + warp;

    int lpx, lpy;
    half_tile_pixel(tile, tiles_x, color, lane, lpx, lpy);
    const bool act = lpx < W && lpy < ylimit && K.weak[lpy * W + lpx] != APDE_WEAK;
    const unsigned amask = __ballot_sync(0xffffffffu, act);
    const int nact = __popc(amask);
    unsigned n_old = 0, n_geom = 0;
    const bool use_geom = K.geom && K.impetus;
    const float dmin = K.depth_min, dmax = K.depth_max;

    for (int base = 0; base < nact; base += 8) {
        __syncwarp();
        if (base + e >= nact) continue;
        const int bit = __fns(amask, 0, base + e + 1);
        int px, py;
        half_tile_pixel(tile, tiles_x, color, bit, px, py);
        const int center = py * W + px;

        QRef rp;
        q_load_ref(K, px, py, rp);
        int pos[8];
        const unsigned flags = checkerboard_candidates(K.costs, W, K.H, px, py, pos);
        // ---- phase 1: 8 candidates x N views
#pragma unroll 1
        for (int h = 0; h < 8; ++h) {
            if ((flags >> h) & 1u) {
                const float3 m = plane_row(K, K.planes[pos[h]]);
#pragma unroll 1
                for (int v = 0; v < N; ++v) {
                    const float c = q_ncc_old(K, s_vk[v], px, py, m, rp, qc);
                    if (qc.q == 0) sc[h * N + v] = c;
                }
                n_old += N;
            } else if (qc.q == 0) {
                for (int v = 0; v < N; ++v) sc[h * N + v] = (h == 0 && v == 0) ? 2.0f : 0.0f;  // quirk 2
            }
        }
        __syncwarp(qc.qmask);
        // ---- phase 2: joint view selection (APD.cu:1339-1386); lanes split the per-view statistics
        uint32_t nbr_sel[4];
        const unsigned nbr_valid = ((flags >> 0) & 1u) | (((flags >> 2) & 1u) << 1) | (((flags >> 4) & 1u) << 2) |
                                   (((flags >> 6) & 1u) << 3);
        nbr_sel[0] = (nbr_valid & 1u) ? K.sel[center - W] : 0u;
        nbr_sel[1] = (nbr_valid & 2u) ? K.sel[center + W] : 0u;
        nbr_sel[2] = (nbr_valid & 4u) ? K.sel[center - 1] : 0u;
        nbr_sel[3] = (nbr_valid & 8u) ? K.sel[center + 1] : 0u;
        {
            const float cost_threshold = (float)(0.8 * (double)__expf((float)(iter * iter) / (-90.0f)));
            const float fallback = __expf(cost_threshold * cost_threshold / (-0.32f));
            for (int v = qc.q; v < N; v += 4) {
                float cnt = 0.0f, tmpw = 0.0f;
                int cnt_false = 0;
#pragma unroll
                for (int h = 0; h < 8; ++h) {
                    const float c = sc[h * N + v];
                    if (c < cost_threshold) { tmpw += __expf(c * c / (-0.18f)); cnt += 1.0f; }
                    if (c > 1.2f) cnt_false++;
                }
                float p = 0.0f;
                if (cnt > 2.0f && cnt_false < 3) p = tmpw / cnt;
                else if (cnt_false < 3) p = fallback;
                float prior = 0.0f;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if ((nbr_valid >> i) & 1u) prior += ((nbr_sel[i] >> v) & 1u) ? 0.9f : 0.1f;
                sp[v] = p * prior;
            }
        }
        __syncwarp(qc.qmask);
        Rng rng(K.seed, K.stream, (uint32_t)center, SITE_STRONG + iter);
        uint4 w = make_uint4(0, 0, 0, 0);
        {
            float psum = 0.0f;
            for (int v = 0; v < N; ++v) psum += sp[v];
            const float inv = 1.0f / psum;
            for (int s = 0; s < 15; ++s) {
                const float r = rng.uniform() - FLT_EPSILON;
                float cum = 0.0f;
                for (int v = 0; v < N; ++v) {
                    cum += sp[v] * inv;  // TransformPDFToCDF, APD.cu:174-188
                    if (cum > r) { vw_inc(w, v); break; }
                }
            }
        }
        uint32_t wmask = 0;
        float wnorm = 0.0f;
        for (int v = 0; v < N; ++v) {
            const uint32_t wv = vw_get(w, v);
            if (wv > 0) { wmask |= 1u << v; wnorm += (float)wv; }
        }
        if (qc.q == 0) K.vw[center] = w;

        float fc_min = 0.0f;
        int min_idx = 0;
#pragma unroll 1
        for (int h = 0; h < 8; ++h) {
            float acc = 0.0f;
            for (uint32_t mk = wmask; mk; mk &= mk - 1) {
                const int v = __ffs(mk) - 1;
                acc += (float)vw_get(w, v) * sc[h * N + v];
            }
            const float fc = acc / wnorm;
            if (h == 0 || fc <= fc_min) { fc_min = fc; min_idx = h; }  // ties -> last (quirk 3); NaN keeps index 0
        }
        // ---- phase 3: current hypothesis + 5 refinement candidates on the selected views
        const float4 plane_c = K.planes[center];
        float4 plane_now = plane_c;
        float cost_now;
        {
            const float3 m = plane_row(K, plane_c);
            float acc = 0.0f;
            for (uint32_t mk = wmask; mk; mk &= mk - 1) {
                const int v = __ffs(mk) - 1;
                float c = q_ncc_old(K, s_vk[v], px, py, m, rp, qc);
                n_old++;
                if (use_geom) { c = c + K.geom_factor * geom_cost(K, s_vk[v], v, px, py, plane_c); n_geom++; }
                acc += (float)vw_get(w, v) * c;
            }
            cost_now = acc / wnorm;
        }
        const float cost_written = cost_now;
        float depth_now = depth_from_plane(K, plane_c, px, py);
        if ((flags >> min_idx) & 1u) {
            const float4 cand = K.planes[pos[min_idx]];
            const float db = depth_from_plane(K, cand, px, py);
            if (db >= dmin && db <= dmax && fc_min < cost_now) {
                depth_now = db; plane_now = cand; cost_now = fc_min;
                if (qc.q == 0) K.sel[center] = wmask;
            }
        }
        const float depth_rand = rng.uniform() * (dmax - dmin) + dmin;
        const float4 n_rand = random_normal(K, px, py, rng, depth_now);
        float depth_pert;
        {
            const float lo = (1.0f - 0.02f) * depth_now, hi = (1.0f + 0.02f) * depth_now;
            depth_pert = rng.uniform() * (hi - lo) + lo;
        }
        const float4 n_pert = perturbed_normal(K, px, py, plane_now, rng, (float)(0.02 * 3.14159265358979323846));
        const float4 base_n = plane_now;
        const float base_d = depth_now;
#pragma unroll 1
        for (int i = 0; i < 5; ++i) {
            float4 tp = (i == 0 || i == 4) ? base_n : (i == 3 ? n_pert : n_rand);
            const float d = (i == 0 || i == 2) ? depth_rand : (i == 4 ? depth_pert : base_d);
            tp.w = distance_to_origin(K, px, py, d, tp);
            const float3 m = plane_row(K, tp);
            float acc = 0.0f;
            for (uint32_t mk = wmask; mk; mk &= mk - 1) {
                const int v = __ffs(mk) - 1;
                float c = q_ncc_old(K, s_vk[v], px, py, m, rp, qc);
                n_old++;
                if (use_geom) { c = c + K.geom_factor * geom_cost(K, s_vk[v], v, px, py, tp); n_geom++; }
                acc += (float)vw_get(w, v) * c;
            }
            const float tc = acc / wnorm;
            const float db = depth_from_plane(K, tp, px, py);
            if (db >= dmin && db <= dmax && tc < cost_now) { depth_now = db; plane_now = tp; cost_now = tc; }
        }
        if (qc.q == 0) {
            if (K.state == APDE_REFINE_INIT) {
                if ((double)cost_now < (double)cost_written - 0.1) { K.costs[center] = cost_now; K.planes[center] = plane_now; }
                else K.costs[center] = cost_written;
            } else {
                K.costs[center] = cost_now;
                K.planes[center] = plane_now;
            }
        }
    }
    q_count(K, qc, n_old, 0, n_geom);
}

// ------------------------------------------------------------------------------------------------ K11 DepthToWeak
// APD.cu:2103-2250, one quad per pixel; the 61-entry cost curve lives in quad-private shared memory.
__global__ void __launch_bounds__(128) k_qdepth_to_weak(const __grid_constant__ PassK K, int tiles_x, float *curve) {
    extern __shared__ float smem[];
    const ViewK *s_vk = stage_views(K, smem);
    const int W = K.W, H = K.H, N = K.N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    QuadCtx qc;
    qc.q = lane & 3;
    qc.qmask = 0xFu << (lane & ~3);
    const int e = lane >> 2;
    float *pc = smem + views_smem_floats(N) + (size_t)(warp * 8 + e) * 64;
    const int tile = blockIdx.x * (blockDim.x >> 5) + warp;
    const int min_margin = 6, radius = 30, n = 61;

    // per-lane prologue: the cheap early outs of the reference (border, zero depth, no selected view)
    int lpx, lpy;
    full_tile_pixel(tile, tiles_x, lane, lpx, lpy);
    bool act = false;
    if (lpx < W && lpy < H) {
        const int c = lpy * W + lpx;
        if (lpx < min_margin || lpy < min_margin || lpx >= W - min_margin || lpy >= H - min_margin) K.weak[c] = APDE_UNKNOWN;
        else if (K.planes[c].w == 0.0f) K.weak[c] = APDE_UNKNOWN;  // rotation does not touch w (APD.cu:2127-2129)
        else if (K.sel[c] == 0u) K.weak[c] = APDE_UNKNOWN;
        else act = true;
    }
    const unsigned amask = __ballot_sync(0xffffffffu, act);
    const int nact = __popc(amask);
    unsigned n_old = 0, n_geom = 0;

    for (int base = 0; base < nact; base += 8) {
        __syncwarp();
        if (base + e >= nact) continue;
        const int bit = __fns(amask, 0, base + e + 1);
        int px, py;
        full_tile_pixel(tile, tiles_x, bit, px, py);
        const int center = py * W + px;
        const float4 opl = normal_to_refcam(K, K.planes[center]);
        const float origin_depth = opl.w;
        const uint32_t sel = K.sel[center];
        const uint4 w = K.vw[center];
        float base_line = 0.0f, weight_normal = 0.0f;
        int valid_src = 0;
        for (uint32_t mk = sel; mk; mk &= mk - 1) {
            const int v = __ffs(mk) - 1;
            weight_normal += (float)vw_get(w, v);
            base_line += s_vk[v].baseline;
            valid_src++;
        }
        base_line /= valid_src;
        QRef rp;
        q_load_ref(K, px, py, rp);
        const float fb = K.fx * base_line;
        const float disp = fb / origin_depth;
#pragma unroll 1
        for (int pd = -radius; pd <= radius; ++pd) {
            const float p_depth = fb / (disp + pd);
            float val = 2.0f;
            if (!(p_depth < K.depth_min || p_depth > K.depth_max)) {
                float4 tp = opl;
                tp.w = distance_to_origin(K, px, py, p_depth, tp);
                const float3 m = plane_row(K, tp);
                float p_cost = 0.0f;
                for (uint32_t mk = sel; mk; mk &= mk - 1) {
                    const int v = __ffs(mk) - 1;
                    float tc = q_ncc_old(K, s_vk[v], px, py, m, rp, qc);
                    n_old++;
                    if (K.geom) { tc += K.geom_factor * geom_cost(K, s_vk[v], v, px, py, tp); n_geom++; }
                    p_cost += tc * (float)vw_get(w, v);
                }
                p_cost /= weight_normal;
                val = (2.0f > p_cost) ? p_cost : 2.0f;  // OpenCV MIN(2.0f, p_cost): NaN -> 2
            }
            if (qc.q == 0) pc[pd + radius] = val;
        }
        __syncwarp(qc.qmask);
        if (curve && qc.q == 0) for (int i = 0; i < n; ++i) curve[(size_t)center * n + i] = pc[i];
        // peak analysis (APD.cu:2200-2249), redundantly on the four lanes
        unsigned long long peaks = 0ull;
        int peak_count = 0, min_peak = 0;
        float min_cost = 2.0f;
        for (int i = 2; i < n - 2; ++i) {
            const float a = pc[i - 1], b = pc[i], c = pc[i + 1];
            if (a > b && c > b) {
                peaks |= 1ull << i;
                peak_count++;
                if (b < min_cost) { min_peak = i; min_cost = b; }
            }
        }
        uint8_t state;
        if (abs(min_peak - radius) > K.weak_peak_radius || pc[min_peak] > 0.5f) state = APDE_WEAK;
        else if (peak_count == 1) state = (pc[min_peak] <= 0.15f) ? APDE_STRONG : APDE_WEAK;
        else {
            float var = 0.0f;
            for (int i = 2; i < n - 2; ++i)
                if (((peaks >> i) & 1ull) && i != min_peak) { const float d = pc[i] - min_cost; var += d * d; }
            var = sqrtf(var);
            var /= (peak_count - 1);
            state = (var > 0.2f) ? APDE_STRONG : APDE_WEAK;
        }
        if (qc.q == 0) K.weak[center] = state;
        __syncwarp(qc.qmask);
    }
    q_count(K, qc, n_old, 0, n_geom);
}

// ------------------------------------------------------------------------------------------------ K13 LocalRefine
// APD.cu:2346-2432, one quad per pixel
__global__ void __launch_bounds__(128) k_qlocal_refine(const __grid_constant__ PassK K, int tiles_x) {
    extern __shared__ float smem[];
    const ViewK *s_vk = stage_views(K, smem);
    const int W = K.W, H = K.H;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    QuadCtx qc;
    qc.q = lane & 3;
    qc.qmask = 0xFu << (lane & ~3);
    const int e = lane >> 2;
    const int tile = blockIdx.x * (blockDim.x >> 5) + warp;
    int lpx, lpy;
    full_tile_pixel(tile, tiles_x, lane, lpx, lpy);
    bool act = false;
    if (lpx < W && lpy < H) {
        const int c = lpy * W + lpx;
        act = K.planes[c].w != 0.0f && K.sel[c] != 0u;
    }
    const unsigned amask = __ballot_sync(0xffffffffu, act);
    const int nact = __popc(amask);
    unsigned n_old = 0, n_geom = 0;
    for (int base = 0; base < nact; base += 8) {
        __syncwarp();
        if (base + e >= nact) continue;
        const int bit = __fns(amask, 0, base + e + 1);
        int px, py;
        full_tile_pixel(tile, tiles_x, bit, px, py);
        const int center = py * W + px;
        const float4 opl = normal_to_refcam(K, K.planes[center]);
        const float origin_depth = opl.w;
        const uint32_t sel = K.sel[center];
        const uint4 w = K.vw[center];
        QRef rp;
        q_load_ref(K, px, py, rp);
        float cost_now = 0.0f, base_line = 0.0f, weight_normal = 0.0f;
        int valid_src = 0;
        {
            float4 tp = opl;
            tp.w = distance_to_origin(K, px, py, origin_depth, tp);
            const float3 m = plane_row(K, tp);
            for (uint32_t mk = sel; mk; mk &= mk - 1) {
                const int v = __ffs(mk) - 1;
                float tc = q_ncc_old(K, s_vk[v], px, py, m, rp, qc);
                n_old++;
                if (K.geom) { tc += K.geom_factor * geom_cost(K, s_vk[v], v, px, py, tp); n_geom++; }
                const float wv = (float)vw_get(w, v);
                cost_now += tc * wv;
                weight_normal += wv;
                base_line += s_vk[v].baseline;
                valid_src++;
            }
        }
        if (weight_normal == 0.0f) continue;
        cost_now /= weight_normal;
        base_line /= valid_src;
        const float fb = K.fx * base_line;
        const float disp = fb / origin_depth;
        float min_cost = 2.0f, best_depth = origin_depth;
#pragma unroll 1
        for (int pd = -5; pd <= 5; ++pd) {
            const float p_depth = fb / (disp + pd);
            if (p_depth < K.depth_min || p_depth > K.depth_max) continue;
            float4 tp = opl;
            tp.w = distance_to_origin(K, px, py, p_depth, tp);
            const float3 m = plane_row(K, tp);
            float tc = 0.0f;
            for (uint32_t mk = sel; mk; mk &= mk - 1) {
                const int v = __ffs(mk) - 1;
                const float wv = (float)vw_get(w, v);
                tc += q_ncc_old(K, s_vk[v], px, py, m, rp, qc) * wv;
                n_old++;
                if (K.geom) { tc += K.geom_factor * geom_cost(K, s_vk[v], v, px, py, tp) * wv; n_geom++; }
            }
            tc /= weight_normal;
            if (tc < min_cost) { min_cost = tc; best_depth = p_depth; }
        }
        if (qc.q == 0 && (double)(cost_now - min_cost) > 0.1) K.planes[center].w = best_depth;
    }
    q_count(K, qc, n_old, 0, n_geom);
}

// ------------------------------------------------------------------------------------------------ launchers
static size_t qprop_smem(int N) { return sizeof(float) * ((size_t)views_smem_floats(N) + (size_t)32 * 9 * N); }
static size_t qsweep_smem(int N) { return sizeof(float) * ((size_t)views_smem_floats(N) + (size_t)32 * 64); }

cudaError_t launch_stage_quad(const PassK &K, int stage, int iter, int color, cudaStream_t st, float *curve, bool *handled) {
    const int W = K.W, H = K.H, N = K.N;
    const int ylimit = min(H, half_rows_limit(H));
    const int tiles8x = (W + 7) / 8;
    *handled = true;
    switch (stage) {
        case APDE_STAGE_PROP_STRONG: {
            const size_t smem = qprop_smem(N);
            static size_t configured = 0;
            if (smem > configured) {
                cudaError_t e = cudaFuncSetAttribute(k_qprop_strong, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return e;
                configured = smem;
            }
            const int tiles = tiles8x * ((ylimit + 7) / 8);
            k_qprop_strong<<<(tiles + 3) / 4, 128, smem, st>>>(K, iter, color, tiles8x, ylimit);
            break;
        }
        case APDE_STAGE_DEPTH_TO_WEAK: {
            const int tiles = tiles8x * ((H + 3) / 4);
            k_qdepth_to_weak<<<(tiles + 3) / 4, 128, qsweep_smem(N), st>>>(K, tiles8x, curve);
            break;
        }
        case APDE_STAGE_LOCAL_REFINE: {
            const int tiles = tiles8x * ((H + 3) / 4);
            k_qlocal_refine<<<(tiles + 3) / 4, 128, sizeof(float) * views_smem_floats(N), st>>>(K, tiles8x);
            break;
        }
        default:
            *handled = false;
            return cudaSuccess;
    }
    return cudaGetLastError();
}

}  // namespace apde
