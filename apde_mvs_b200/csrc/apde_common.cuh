// apde_common.cuh -- device helpers shared by the pass kernels: pixel mappings, view-constant staging and the
// multi-hypothesis joint view selection.
#pragma once
#include <cfloat>

#include "apde_device.cuh"

namespace apde {

// -------------------------------------------------------------------------------------------- pixel mappings
// checkerboard kernels: one warp = the 32 pixels of one colour inside a tile of 64 pixels, (1 << tile_shift) wide: 8x8 by
// default (compact texture footprint), 16x4 / 32x2 / 64x1 as experiment knobs (APDE_TILE_SHIFT)
__device__ __forceinline__ bool half_pixel(const PassK &K, int color, int tiles_x, int ylimit, int &px, int &py) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (blockDim.x >> 5) + warp;
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    const int sh = K.tile_shift;
    py = ty * (64 >> sh) + (lane >> (sh - 1));
    px = (tx << sh) + 2 * (lane & ((1 << (sh - 1)) - 1)) + ((py + color) & 1);
    return px < K.W && py < ylimit;
}
// grid of the implicit checkerboard mapping
__host__ __device__ inline int half_tiles_x(const PassK &K) { return (K.W + (1 << K.tile_shift) - 1) >> K.tile_shift; }
__host__ __device__ inline int half_tiles(const PassK &K, int ylimit) {
    const int th = 64 >> K.tile_shift;
    return half_tiles_x(K) * ((ylimit + th - 1) / th);
}
// checkerboard kernels in rounds that have WEAK pixels: threads walk a compacted pixel list (no idle lanes on the
// other class); without a list the implicit tile mapping above is used
__device__ __forceinline__ bool half_pixel_or_list(const PassK &K, int color, int tiles_x, int ylimit, int &px, int &py) {
    if (K.list == nullptr) return half_pixel(K, color, tiles_x, ylimit, px, py);
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= *K.list_count) return false;
    const int center = K.list[idx];
    px = center % K.W;
    py = center / K.W;
    return true;
}
// full kernels: one warp = an 8x4 tile
__device__ __forceinline__ bool full_pixel(const PassK &K, int tiles_x, int &px, int &py) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (blockDim.x >> 5) + warp;
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    py = ty * 4 + (lane >> 3);
    px = tx * 8 + (lane & 7);
    return px < K.W && py < K.H;
}

// copy the per-view constants into shared memory (divergent per-lane view indices must not hit the constant bank)
__device__ __forceinline__ ViewK *stage_views(const PassK &K, float *smem) {
    const int words = K.N * (int)(sizeof(ViewK) / 4);
    const uint32_t *src = reinterpret_cast<const uint32_t *>(K.v);
    uint32_t *dst = reinterpret_cast<uint32_t *>(smem);
    for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
    return reinterpret_cast<ViewK *>(smem);
}
__host__ __device__ inline int views_smem_floats(int N) { return ((N * (int)sizeof(ViewK) + 15) / 16) * 4; }

// Multi-hypothesis joint view selection, APD.cu:1339-1386.  cost(h, v) = sc[(h * N + v) * stride]; prob scratch
// sp[v * stride].  Returns the packed weights; *selmask = views with weight > 0, *wnorm = sum of weights.
template <int NBR>
__device__ __forceinline__ uint4 select_views(const PassK &K, const float *sc, float *sp, int stride,
                                              const uint32_t (&nbr_sel)[NBR], unsigned nbr_valid, int iter, Rng &rng,
                                              uint32_t *selmask, float *wnorm) {
    const int N = K.N;
    const float cost_threshold = (float)(0.8 * (double)__expf((float)(iter * iter) / (-90.0f)));
    const float fallback = __expf(cost_threshold * cost_threshold / (-0.32f));
    float psum = 0.0f;
    for (int v = 0; v < N; ++v) {
        float cnt = 0.0f, tmpw = 0.0f;
        int cnt_false = 0;
#pragma unroll
        for (int h = 0; h < 8; ++h) {
            const float c = sc[(h * N + v) * stride];
            if (c < cost_threshold) { tmpw += __expf(c * c / (-0.18f)); cnt += 1.0f; }
            if (c > 1.2f) cnt_false++;
        }
        float p = 0.0f;
        if (cnt > 2.0f && cnt_false < 3) p = tmpw / cnt;
        else if (cnt_false < 3) p = fallback;
        float prior = 0.0f;
#pragma unroll
        for (int i = 0; i < NBR; ++i)
            if ((nbr_valid >> i) & 1u) prior += ((nbr_sel[i] >> v) & 1u) ? 0.9f : 0.1f;
        p *= prior;
        sp[v * stride] = p;
        psum += p;
    }
    // TransformPDFToCDF, APD.cu:174-188
    const float inv = 1.0f / psum;
    float cum = 0.0f;
    for (int v = 0; v < N; ++v) { cum += sp[v * stride] * inv; sp[v * stride] = cum; }
    uint4 w = make_uint4(0, 0, 0, 0);
    for (int s = 0; s < 15; ++s) {
        const float r = rng.uniform() - FLT_EPSILON;
        for (int v = 0; v < N; ++v) {
            if (sp[v * stride] > r) { vw_inc(w, v); break; }
        }
    }
    uint32_t mask = 0;
    float wn = 0.0f;
    for (int v = 0; v < N; ++v) {
        const uint32_t wv = vw_get(w, v);
        if (wv > 0) { mask |= 1u << v; wn += (float)wv; }
    }
    *selmask = mask;
    *wnorm = wn;
    return w;
}


// -------------------------------------------------------------------------------------------- checkerboard candidates
// APD.cu:1127-1314.  pos[k]: 0 up_near 1 up_far 2 down_near 3 down_far 4 left_near 5 left_far 6 right_near 7 right_far
__device__ __forceinline__ unsigned checkerboard_candidates(const float *costs, int width, int height, int px,
                                                            int py, int pos[8]) {
    const int center = py * width + px;
    unsigned flags = 0;
    float cmin;
    int cpt;
    int up_far = center - 3 * width, down_far = center + 3 * width, left_far = center - 3, right_far = center + 3;
    int up_near = center - width, down_near = center + width, left_near = center - 1, right_near = center + 1;
#define APDE_TRY(cond, idx)                                      \
    if (cond) {                                                  \
        const int pt_ = (idx);                                   \
        const float c_ = costs[pt_];                     \
        if (c_ < cmin) { cmin = c_; cpt = pt_; }                 \
    }
    if (py > 2) {
        flags |= 2u; cmin = costs[up_far]; cpt = up_far;
#pragma unroll
        for (int i = 1; i < 11; ++i) APDE_TRY(py > 2 + 2 * i, up_far - 2 * i * width)
        up_far = cpt;
    }
    if (py < height - 3) {
        flags |= 8u; cmin = costs[down_far]; cpt = down_far;
#pragma unroll
        for (int i = 1; i < 11; ++i) APDE_TRY(py < height - 3 - 2 * i, down_far + 2 * i * width)
        down_far = cpt;
    }
    if (px > 2) {
        flags |= 32u; cmin = costs[left_far]; cpt = left_far;
#pragma unroll
        for (int i = 1; i < 11; ++i) APDE_TRY(px > 2 + 2 * i, left_far - 2 * i)
        left_far = cpt;
    }
    if (px < width - 3) {
        flags |= 128u; cmin = costs[right_far]; cpt = right_far;
#pragma unroll
        for (int i = 1; i < 11; ++i) APDE_TRY(px < width - 3 - 2 * i, right_far + 2 * i)
        right_far = cpt;
    }
    if (py > 0) {
        flags |= 1u; cmin = costs[up_near]; cpt = up_near;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            APDE_TRY(py > 1 + i && px > i, up_near - (1 + i) * width - (i + 1))
            APDE_TRY(py > 1 + i && px < width - 1 - i, up_near - (1 + i) * width + (i + 1))
        }
        up_near = cpt;
    }
    if (py < height - 1) {
        flags |= 4u; cmin = costs[down_near]; cpt = down_near;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            APDE_TRY(py < height - 2 - i && px > i, down_near + (1 + i) * width - (i + 1))
            APDE_TRY(py < height - 2 - i && px < width - 1 - i, down_near + (1 + i) * width + (i + 1))
        }
        down_near = cpt;
    }
    if (px > 0) {
        flags |= 16u; cmin = costs[left_near]; cpt = left_near;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            APDE_TRY(px > 1 + i && py > i, left_near - (1 + i) - (i + 1) * width)
            APDE_TRY(px > 1 + i && py < height - 1 - i, left_near - (1 + i) + (i + 1) * width)
        }
        left_near = cpt;
    }
    if (px < width - 1) {
        flags |= 64u; cmin = costs[right_near]; cpt = right_near;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            APDE_TRY(px < width - 2 - i && py > i, right_near + (1 + i) - (i + 1) * width)
            APDE_TRY(px < width - 2 - i && py < height - 1 - i, right_near + (1 + i) + (i + 1) * width)
        }
        right_near = cpt;
    }
#undef APDE_TRY
    pos[0] = up_near; pos[1] = up_far; pos[2] = down_near; pos[3] = down_far;
    pos[4] = left_near; pos[5] = left_far; pos[6] = right_near; pos[7] = right_far;
    return flags;
}


static inline int half_rows_limit(int H) { return 32 * (((H / 2) + 15) / 16); }  // quirk 7, APD.cu:2676-2678

}  // namespace apde
