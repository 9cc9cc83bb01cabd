// apde_device.cuh -- device-side types, counter RNG and per-hypothesis cost functions (sm_100a).
//
// Design: one thread per pixel; the 36 reference-patch texels and their sums are loaded ONCE per
// pixel per kernel and stay in registers (the reference re-fetches them through the texture unit for every
// evaluation, APD.cu:632); the camera-pair part of the homography is precomputed on the host per (ref, src) pair
// (the reference recomputes it per evaluation, APD.cu:336-362), so that H = A - b m^T costs 9 FMAs; source texels
// are gathered through ONE layered texture object (linear filter, clamp) holding every view of the pyramid level.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/apde.h"

namespace apde {

constexpr int kMaxSrc = APDE_MAX_IMAGES - 1;
constexpr int kPatch = 36;  // (2*5/2+1)^2 samples: strong_radius 5, strong_increment 2 (main.h:87-88)

// per (reference, source) pair constants, precomputed in double on the host
struct ViewK {
    float A[9];   // K_s R_rel K_r^-1                (homography:  H = A - b m^T,  m = n^T K_r^-1 / w)
    float b[3];   // K_s t_rel,  t_rel = R_s (C_r - C_s)
    float Ai[9];  // K_r R_rel^T K_s^-1              (backward reprojection)
    float bi[3];  // K_r R_r (C_s - C_r)
    float baseline;
    int layer;    // layer of this source view in the level's texture
    int pad_[2];  // 112 bytes: 16-byte aligned rows in shared memory
};

struct PassK {
    int W, H, N;  // N = number of source views
    int state, geom, impetus, use_apd, top_k, weak_peak_radius, rotate_time, max_iterations;
    float depth_min, depth_max, geom_factor, ransac_threshold;
    float fx, fy, cx, cy;  // reference intrinsics at this level
    float R[9];            // reference rotation
    int ref_layer;
    uint32_t seed, stream;
    cudaTextureObject_t tex;  // layered images of the level (linear filter, clamp, unnormalised coordinates)
    float tex_unorm, tex_inv; // 0: float32 texels.  > 0: integer texels read as normalised float; sample = rint(t * tex_unorm) * tex_inv
    float4 *planes;
    float *costs;
    uint32_t *sel;
    uint4 *vw;  // packed view weights: 4 bits per source view (0..15)
    uint8_t *weak;
    uint8_t *conf;
    float4 *fit;
    uint8_t *reliable;
    short2 *nearest;
    uint16_t *ns_tiles;    // [ceil(H/8)][ceil(W/8)] nearest-strong pruning summary: 0 = no STRONG pixel in the 8x8 tile, else 1 + max confidence
    short2 *anchors;       // [P][9]
    const float *depth;    // [N+1][P] working-resolution depth maps, 0 = ref
    const uint8_t *sa;     // [P] segment labels of the reference view at this level (sa_masks/*.bin, APD.cpp:641-649); nullptr = all zero
    unsigned long long *counters;
    // compacted pixel lists of the checkerboard kernels (rounds with WEAK pixels): [colour][strong | weak], see k_build_lists
    const int *list;  // nullptr: implicit 8x8-tile mapping
    const int *list_count;
    int list_pair_stride;  // > 0: blockIdx.y = colour selects list + y * stride, count[2 * y] (both weak lists in one launch)
    ViewK v[kMaxSrc];
};

// ------------------------------------------------------------------------------------------------ RNG
// Philox4x32-10, key = (seed, stream), counter = (pixel, site, block, 0).  Same spec as oracle/apd_oracle.cpp.
enum Site : uint32_t { SITE_ANCHOR = 1, SITE_INIT = 2, SITE_STRONG = 16, SITE_FIT = 48, SITE_WEAK = 80 };

__device__ __forceinline__ void philox4x32_10(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0,
                                              uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

struct Rng {
    uint32_t k0, k1, pixel, site, n;
    uint32_t b0, b1, b2, b3;
    __device__ __forceinline__ Rng(uint32_t seed, uint32_t stream, uint32_t pixel_, uint32_t site_)
        : k0(seed), k1(stream), pixel(pixel_), site(site_), n(0), b0(0), b1(0), b2(0), b3(0) {}
    __device__ __forceinline__ uint32_t next() {
        const uint32_t lane = n & 3u;
        if (lane == 0u) {
            b0 = pixel; b1 = site; b2 = n >> 2; b3 = 0u;
            philox4x32_10(b0, b1, b2, b3, k0, k1);
        }
        ++n;
        return lane == 0u ? b0 : (lane == 1u ? b1 : (lane == 2u ? b2 : b3));
    }
    // (0,1], curand_uniform's mapping: x * 2^-32 + 2^-33
    __device__ __forceinline__ float uniform() {
        return __fmaf_rn(__uint2float_rn(next()), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    }
};

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }
// single MUFU.RCP (what "x / z" compiles to under the reference's --use_fast_math, CMakeLists.txt:26)
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// reference-camera plane helpers (APD.cu:190-240), zero-skew pinhole
// (explicit intrinsics, see plane_row: the same plane must give the same depth in every kernel it is inlined into)
__device__ __forceinline__ float depth_from_plane(const PassK &K, float4 pl, int px, int py) {
    const float a = __fmul_rn(__fsub_rn((float)px, K.cx), pl.x);
    const float b = __fmul_rn(__fdividef(K.fx, K.fy), __fsub_rn((float)py, K.cy));
    const float den = __fmaf_rn(K.fx, pl.z, __fmaf_rn(b, pl.y, a));
    return __fdividef(__fmul_rn(-pl.w, K.fx), den);
}
__device__ __forceinline__ float distance_to_origin(const PassK &K, int px, int py, float depth, float4 n) {
    const float X0 = depth * (px - K.cx) / K.fx, X1 = depth * (py - K.cy) / K.fy;
    return -(n.x * X0 + n.y * X1 + n.z * depth);
}
__device__ __forceinline__ float3 view_direction(const PassK &K, int px, int py, float depth) {
    const float X0 = depth * (px - K.cx) / K.fx, X1 = depth * (py - K.cy) / K.fy, X2 = depth;
    const float nrm = sqrtf(X0 * X0 + X1 * X1 + X2 * X2);
    return make_float3(X0 / nrm, X1 / nrm, X2 / nrm);
}
__device__ __forceinline__ void normalize3(float4 &v) {
    const float inv = rsqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
    v.x *= inv; v.y *= inv; v.z *= inv;
}
__device__ __forceinline__ float4 normal_to_world(const PassK &K, float4 p) {  // TransformNormal APD.cu:405
    return make_float4(K.R[0] * p.x + K.R[3] * p.y + K.R[6] * p.z, K.R[1] * p.x + K.R[4] * p.y + K.R[7] * p.z,
                       K.R[2] * p.x + K.R[5] * p.y + K.R[8] * p.z, p.w);
}
__device__ __forceinline__ float4 normal_to_refcam(const PassK &K, float4 p) {  // TransformNormal2RefCam APD.cu:415
    return make_float4(K.R[0] * p.x + K.R[1] * p.y + K.R[2] * p.z, K.R[3] * p.x + K.R[4] * p.y + K.R[5] * p.z,
                       K.R[6] * p.x + K.R[7] * p.y + K.R[8] * p.z, p.w);
}

__device__ __forceinline__ float4 random_normal(const PassK &K, int px, int py, Rng &rng, float depth) {  // APD.cu:242
    float q1 = 1.0f, q2 = 1.0f, s = 2.0f;
    while (s >= 1.0f) {
        q1 = 2.0f * rng.uniform() - 1.0f;
        q2 = 2.0f * rng.uniform() - 1.0f;
        s = q1 * q1 + q2 * q2;
    }
    const float sq = sqrtf(1.0f - s);
    float4 n = make_float4(2.0f * q1 * sq, 2.0f * q2 * sq, 1.0f - 2.0f * s, 0.0f);
    const float3 vd = view_direction(K, px, py, depth);
    if (n.x * vd.x + n.y * vd.y + n.z * vd.z > 0.0f) { n.x = -n.x; n.y = -n.y; n.z = -n.z; }
    normalize3(n);
    return n;
}

__device__ __forceinline__ float4 perturbed_normal(const PassK &K, int px, int py, float4 normal, Rng &rng,
                                                   float perturbation) {  // APD.cu:270
    const float3 vd = view_direction(K, px, py, 1.0f);
    const float a1 = (rng.uniform() - 0.5f) * perturbation;
    const float a2 = (rng.uniform() - 0.5f) * perturbation;
    const float a3 = (rng.uniform() - 0.5f) * perturbation;
    float s1, c1, s2, c2, s3, c3;
    sincosf(a1, &s1, &c1); sincosf(a2, &s2, &c2); sincosf(a3, &s3, &c3);
    const float r0 = c2 * c3, r1 = c3 * s1 * s2 - c1 * s3, r2 = s1 * s3 + c1 * c3 * s2;
    const float r3 = c2 * s3, r4 = c1 * c3 + s1 * s2 * s3, r5 = c1 * s2 * s3 - c3 * s1;
    const float r6 = -s2, r7 = c2 * s1, r8 = c1 * c2;
    float4 np = make_float4(r0 * normal.x + r1 * normal.y + r2 * normal.z, r3 * normal.x + r4 * normal.y + r5 * normal.z,
                            r6 * normal.x + r7 * normal.y + r8 * normal.z, 0.0f);
    if (np.x * vd.x + np.y * vd.y + np.z * vd.z >= 0.0f) np = normal;
    normalize3(np);
    return np;
}

// One filtered sample of the level.  With integer-valued texels the exact filtered value is a multiple of 1/256 (the
// texture unit's four weights are integers that sum to 256, profiles/r01_texture_filter_model.md), so a level stored as
// 8-bit UNORM texels -- a quarter of the bytes behind every gather, and the L1TEX data pipe is the binding unit -- gives
// back exactly the float32 texture's sample after one multiply and one round: rint(t * 255 * 256) / 256.  Coarser pyramid
// levels of even sizes are 2x2 means (multiples of 1/4): stored x4 as 16-bit UNORM, sample = rint(t * 65535 * 256) / 1024.
// U (compile time) = the level holds UNORM texels.  Decode = 2 FMA-pipe instructions: y = fma(t, 255*256, 1.5*2^23) is
// 1.5*2^23 + n with n = rint(t*255*256) exactly, and fma(y, 2^-8, -1.5*2^15) = n / 256 exactly.
template <bool U>
__device__ __forceinline__ float fetch(const PassK &K, float x, float y, int layer) {
    const float t = tex2DLayered<float>(K.tex, x, y, layer);
    if (!U) return t;
    const float y2 = fmaf(t, K.tex_unorm, 12582912.0f);
    return fmaf(y2, K.tex_inv, -12582912.0f * K.tex_inv);
}

// ------------------------------------------------------------------------------------------------ costs
// reference patch of one pixel: texels at (x+i, y+j), i outer, j inner, i,j in {-5,-3,...,5} (APD.cu:629-632).
// A texel-centre fetch of the clamped linear texture returns the texel exactly.
struct RefPatch {
    float r[kPatch];
    float mean, var;  // E[r], E[r^2]-E[r]^2 with the reference's accumulation order
};

template <bool U>
__device__ __forceinline__ void load_ref_patch(const PassK &K, int px, int py, RefPatch &rp) {
    float s = 0.0f, ss = 0.0f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const float v = fetch<U>(K, px + (2 * i - 5) + 0.5f, py + (2 * j - 5) + 0.5f, K.ref_layer);
            rp.r[i * 6 + j] = v;
            s += v;
            ss = fmaf(v, v, ss);
        }
    }
    // epilogue arithmetic exactly as the reference build's SASS: mean = inv * sum; var = FFMA(inv, sum_xx, -(mean * mean))
    const float inv = 1.0f / 36.0f;
    rp.mean = inv * s;
    rp.var = fmaf(inv, ss, -__fmul_rn(rp.mean, rp.mean));
}

// m = n^T K_r^-1 / w : the hypothesis-dependent row vector of H = A - b m^T
// Explicit intrinsics: every kernel that evaluates a hypothesis must produce the SAME row vector from the same plane, whatever
// the compiler would contract around an inlined copy (a 1-ulp difference moves sample coordinates across the texture unit's
// 1/256 weight buckets and shows up as 1e-4-level cost differences between kernel variants).
__device__ __forceinline__ float3 plane_row(const PassK &K, float4 pl) {
    const float iw = rcp_approx(pl.w);
    const float mx = __fdividef(pl.x, K.fx), my = __fdividef(pl.y, K.fy);
    const float mz = __fmaf_rn(-my, K.cy, __fmaf_rn(-mx, K.cx, pl.z));
    return make_float3(__fmul_rn(mx, iw), __fmul_rn(my, iw), __fmul_rn(mz, iw));
}

struct Homog { float h[9]; };

__device__ __forceinline__ Homog make_homography(const ViewK &vk, float3 m) {
    Homog Hm;
    Hm.h[0] = fmaf(-vk.b[0], m.x, vk.A[0]); Hm.h[1] = fmaf(-vk.b[0], m.y, vk.A[1]); Hm.h[2] = fmaf(-vk.b[0], m.z, vk.A[2]);
    Hm.h[3] = fmaf(-vk.b[1], m.x, vk.A[3]); Hm.h[4] = fmaf(-vk.b[1], m.y, vk.A[4]); Hm.h[5] = fmaf(-vk.b[1], m.z, vk.A[5]);
    Hm.h[6] = fmaf(-vk.b[2], m.x, vk.A[6]); Hm.h[7] = fmaf(-vk.b[2], m.y, vk.A[7]); Hm.h[8] = fmaf(-vk.b[2], m.z, vk.A[8]);
    return Hm;
}

// NCC of the warped 6x6 patch against the register-resident reference patch.  APD.cu:622-662.
template <bool U>
__device__ __forceinline__ float patch_ncc36(const PassK &K, const Homog &Hm, int layer, int px, int py,
                                             const RefPatch &rp) {
    const float *h = Hm.h;
    // fold the +0.5 texel-centre offset into the numerators: (X + 0.5 Z) / Z
    const float g0 = fmaf(0.5f, h[6], h[0]), g1 = fmaf(0.5f, h[7], h[1]), g2 = fmaf(0.5f, h[8], h[2]);
    const float g3 = fmaf(0.5f, h[6], h[3]), g4 = fmaf(0.5f, h[7], h[4]), g5 = fmaf(0.5f, h[8], h[5]);
    float sum_s = 0.0f, sum_ss = 0.0f, sum_rs = 0.0f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const float xi = (float)(px + 2 * i - 5);
        const float bx = fmaf(g0, xi, g2), by = fmaf(g3, xi, g5), bz = fmaf(h[6], xi, h[8]);
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const float yj = (float)(py + 2 * j - 5);
            const float X = fmaf(g1, yj, bx), Y = fmaf(g4, yj, by), Z = fmaf(h[7], yj, bz);
            const float iz = rcp_approx(Z);
            const float s = fetch<U>(K, X * iz, Y * iz, layer);
            sum_s += s;
            sum_ss = fmaf(s, s, sum_ss);
            sum_rs = fmaf(rp.r[i * 6 + j], s, sum_rs);
        }
    }
    const float inv = 1.0f / 36.0f;
    const float mean_s = inv * sum_s, e_rs = inv * sum_rs;
    const float var_s = fmaf(inv, sum_ss, -__fmul_rn(mean_s, mean_s));
    if (rp.var < 1e-5f || var_s < 1e-5f) return 2.0f;
    const float covar = fmaf(-rp.mean, mean_s, e_rs);
    return fmaxf(0.0f, fminf(2.0f, fmaf(-covar, rsqrtf(rp.var * var_s), 1.0f)));
}

// ComputeBilateralNCCOld, APD.cu:596-663 (branch A).  Only the patch centre is bounds-checked (quirk 8).
template <bool U>
__device__ __forceinline__ float ncc_old(const PassK &K, const ViewK &vk, int px, int py, float3 m, const RefPatch &rp) {
    const Homog Hm = make_homography(vk, m);
    const float *h = Hm.h;
    const float fxp = (float)px, fyp = (float)py;
    const float Z = __fadd_rn(__fmaf_rn(h[7], fyp, __fmul_rn(h[6], fxp)), h[8]);
    const float iz = rcp_approx(Z);
    const float ptx = __fmul_rn(__fadd_rn(__fmaf_rn(h[1], fyp, __fmul_rn(h[0], fxp)), h[2]), iz);
    const float pty = __fmul_rn(__fadd_rn(__fmaf_rn(h[4], fyp, __fmul_rn(h[3], fxp)), h[5]), iz);
    if (ptx >= (float)K.W || ptx < 0.0f || pty >= (float)K.H || pty < 0.0f) return 2.0f;
    return patch_ncc36<U>(K, Hm, vk.layer, px, py, rp);
}

// Reference side of the (up to) 8 anchor patches of a WEAK pixel: 3x3 taps (weak_radius 5, weak_increment 5) and their
// statistics.  They depend on the pixel only, so they are gathered once per pixel per kernel (the reference re-fetches
// them for every evaluation, APD.cu:531) -- per evaluation only the 9 SOURCE samples per anchor remain.
struct AnchorRef {
    float r[8 * 9];
    float mean[8], var[8];
    short2 a[8];  // x == -1: no anchor in this slot
};

template <bool U>
__device__ __forceinline__ void load_anchor_ref(const PassK &K, const short2 *anc, AnchorRef &ar) {
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
        const short2 a = anc[k + 1];
        ar.a[k] = a;
        if (a.x == -1 || a.y == -1) { ar.a[k].x = -1; continue; }
        float sr = 0.0f, srr = 0.0f;
        int t = 0;
#pragma unroll
        for (int i = -5; i <= 5; i += 5) {
#pragma unroll
            for (int j = -5; j <= 5; j += 5) {
                const float r = fetch<U>(K, (float)(a.x + i) + 0.5f, (float)(a.y + j) + 0.5f, K.ref_layer);
                ar.r[k * 9 + t++] = r;
                sr += r;
                srr = fmaf(r, r, srr);
            }
        }
        const float inv = 1.0f / 9.0f;
        ar.mean[k] = inv * sr;
        ar.var[k] = fmaf(inv, srr, -__fmul_rn(ar.mean[k], ar.mean[k]));
    }
}

// NCC of one 3x3 anchor patch: only the source samples are gathered here
template <bool U>
__device__ __forceinline__ float patch_ncc9(const PassK &K, const Homog &Hm, int layer, int ax, int ay, const float *r9,
                                            float mean_r, float var_r) {
    const float *h = Hm.h;
    float ss = 0.0f, sss = 0.0f, srs = 0.0f;
    int t = 0;
#pragma unroll
    for (int i = -5; i <= 5; i += 5) {
#pragma unroll
        for (int j = -5; j <= 5; j += 5) {
            const float fxp = (float)(ax + i), fyp = (float)(ay + j);
            const float Z = h[6] * fxp + h[7] * fyp + h[8];
            const float iz = rcp_approx(Z);
            const float X = (h[0] * fxp + h[1] * fyp + h[2]) * iz, Y = (h[3] * fxp + h[4] * fyp + h[5]) * iz;
            const float s = fetch<U>(K, X + 0.5f, Y + 0.5f, layer);
            ss += s; sss = fmaf(s, s, sss); srs = fmaf(r9[t++], s, srs);
        }
    }
    const float inv = 1.0f / 9.0f;
    const float mean_s = inv * ss, e_rs = inv * srs;
    const float var_s = fmaf(inv, sss, -__fmul_rn(mean_s, mean_s));
    if (var_r < 1e-5f || var_s < 1e-5f) return 2.0f;
    const float covar = fmaf(-mean_r, mean_s, e_rs);
    return fmaxf(0.0f, fminf(2.0f, fmaf(-covar, rsqrtf(var_r * var_s), 1.0f)));
}

// ComputeBilateralNCCNew, APD.cu:448-593 (sa_mask == 0): centre patch + focal-weighted anchor patches.
template <bool U>
__device__ __forceinline__ float ncc_new(const PassK &K, const ViewK &vk, int view_bit, int px, int py, float3 m,
                                         const RefPatch &rp, const AnchorRef &ar) {
    const Homog Hm = make_homography(vk, m);
    const float *h = Hm.h;
    const float fW = (float)K.W, fH = (float)K.H;
    {
        const float fxp = (float)px, fyp = (float)py;
        const float iz = rcp_approx(h[6] * fxp + h[7] * fyp + h[8]);
        const float ptx = (h[0] * fxp + h[1] * fyp + h[2]) * iz, pty = (h[3] * fxp + h[4] * fyp + h[5]) * iz;
        if (ptx >= fW || ptx < 0.0f || pty >= fH || pty < 0.0f) return 2.0f;
    }
    // anchor 0 is the pixel itself (APD.cu:1887): its bounds test repeats the centre test above
    const float center_cost = patch_ncc36<U>(K, Hm, vk.layer, px, py, rp);
    float sc[8];
    int ns = 0;
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
        const short2 a = ar.a[k];
        if (a.x == -1) continue;
        const float fxp = (float)a.x, fyp = (float)a.y;
        const float iz = rcp_approx(h[6] * fxp + h[7] * fyp + h[8]);
        const float ax = (h[0] * fxp + h[1] * fyp + h[2]) * iz, ay = (h[3] * fxp + h[4] * fyp + h[5]) * iz;
        if (ax < 0.0f || ay < 0.0f || ax >= fW || ay >= fH) {
            if ((K.sel[a.x + a.y * K.W] >> view_bit) & 1u) sc[ns++] = 2.0f;
            continue;
        }
        sc[ns++] = patch_ncc9<U>(K, Hm, vk.layer, a.x, a.y, &ar.r[k * 9], ar.mean[k], ar.var[k]);
    }
    if (ns == 0) return center_cost;
    float mx = -1e10f;
    for (int i = 0; i < ns; ++i) mx = fmaxf(mx, sc[i]);
    float sum = 0.0f, acc = 0.0f;
    float wts[8];
    for (int i = 0; i < ns; ++i) { wts[i] = __expf(sc[i] - mx); sum += wts[i]; }
    for (int i = 0; i < ns; ++i) acc += (wts[i] / sum) * sc[i];
    acc = fminf(acc, 2.0f);
    return 0.25f * center_cost + 0.75f * acc;
}

// ------------------------------------------------------------------------------------------------ segment labels (SAM masks)
// With a label map (sa_masks/<id>.bin, written by tools/run_SAM.py, consumed when use_sa && use_APD, APD.cpp:641-649) the
// reference restricts the patches to the centre pixel's segment:
//   NCC-Old (APD.cu:619-719): branch B is taken when the label AT THE PROJECTED POINT'S INDEX in the reference label map
//     ("center = pt.y * src_camera.width + pt.x", a float expression truncated to int -- quirk, reproduced) is non-zero.  It
//     walks the same 36 odd offsets as branch A, quadrant by quadrant in a fixed order, skips taps outside the image and
//     leaves a quadrant at the first tap of another label.
//   NCC-New (APD.cu:463-465, 493-497, 526-530): anchors and taps of another label than the (non-zero) centre label are dropped.
// Which taps take part depends on the pixel only, so the participation masks and the reference-side sums are built once per
// pixel per kernel (SaInfo / AnchorRefSa) and an evaluation walks the mask.  Kernels carry these in <.., SA = true>
// instantiations launched only for problems with a label map; the SA = false kernels are unchanged.
struct SaInfo {
    unsigned long long bmask;  // NCC-Old branch B: bit t = tap t of the quadrant walk takes part
    unsigned long long nmask;  // NCC-New centre patch: bit i*6+j (branch A order)
    float b_mean, b_var, b_inv;
    float n_mean, n_var, n_inv;
    int label;
};
struct SaNone {};

__device__ __forceinline__ constexpr int sa_b_xoff(int t) {  // APD.cu:665-666
    const int q = t / 9, j = t % 9;
    const int ox = (j == 0 || j == 2 || j == 3) ? 1 : (j == 1 || j == 4 || j == 7) ? 3 : 5;
    return (q == 0 || q == 2) ? ox : -ox;
}
__device__ __forceinline__ constexpr int sa_b_yoff(int t) {
    const int q = t / 9, j = t % 9;
    const int oy = (j == 0 || j == 1 || j == 5) ? 1 : (j == 2 || j == 4 || j == 6) ? 3 : 5;
    return (q == 0 || q == 3) ? oy : -oy;
}
__device__ __forceinline__ float sa_inv_count(int cnt, int full, float inv_full) {
    return cnt == full ? inv_full : rcp_approx((float)cnt);  // 1.0f / bilateral_weight_sum under --use_fast_math
}

__device__ __forceinline__ void load_sa_info(const PassK &K, int px, int py, const RefPatch &rp, SaInfo &si) {
    const uint8_t *sa = K.sa;
    const int W = K.W, H = K.H;
    const int lab = sa[py * W + px];
    si.label = lab;
    {
        unsigned long long bm = 0ull;
        float s = 0.0f, ss = 0.0f;
        int cnt = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            bool open = true;
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                const int t = q * 9 + j;
                const int xo = sa_b_xoff(t), yo = sa_b_yoff(t);
                const int x = px + xo, y = py + yo;
                if (open && x >= 0 && x < W && y >= 0 && y < H) {
                    if (sa[y * W + x] != lab) open = false;
                    else {
                        const float v = rp.r[((xo + 5) / 2) * 6 + (yo + 5) / 2];
                        bm |= 1ull << t;
                        s += v;
                        ss = fmaf(v, v, ss);
                        cnt++;
                    }
                }
            }
        }
        si.bmask = bm;
        si.b_inv = sa_inv_count(cnt, 36, 1.0f / 36.0f);
        si.b_mean = si.b_inv * s;
        si.b_var = fmaf(si.b_inv, ss, -__fmul_rn(si.b_mean, si.b_mean));
    }
    {
        unsigned long long nm = 0ull;
        float s = 0.0f, ss = 0.0f;
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                // the reference indexes the label map without a bounds test here (APD.cu:527); WEAK pixels are >= 6 px inside
                const int x = clampi(px + 2 * i - 5, 0, W - 1), y = clampi(py + 2 * j - 5, 0, H - 1);
                if (lab == 0 || sa[y * W + x] == lab) {
                    const float v = rp.r[i * 6 + j];
                    nm |= 1ull << (i * 6 + j);
                    s += v;
                    ss = fmaf(v, v, ss);
                    cnt++;
                }
            }
        }
        si.nmask = nm;
        si.n_inv = sa_inv_count(cnt, 36, 1.0f / 36.0f);
        si.n_mean = si.n_inv * s;
        si.n_var = fmaf(si.n_inv, ss, -__fmul_rn(si.n_mean, si.n_mean));
    }
}

// masked twin of patch_ncc36: ORDER 0 = branch A order (i outer, j inner), ORDER 1 = the quadrant walk.  Sample coordinates
// are formed exactly as in patch_ncc36, so a full mask in order 0 gives the same samples.
template <bool U, int ORDER>
__device__ __forceinline__ float patch_ncc36_masked(const PassK &K, const Homog &Hm, int layer, int px, int py, const RefPatch &rp,
                                                    unsigned long long mask, float mean_r, float var_r, float inv) {
    const float *h = Hm.h;
    const float g0 = fmaf(0.5f, h[6], h[0]), g1 = fmaf(0.5f, h[7], h[1]), g2 = fmaf(0.5f, h[8], h[2]);
    const float g3 = fmaf(0.5f, h[6], h[3]), g4 = fmaf(0.5f, h[7], h[4]), g5 = fmaf(0.5f, h[8], h[5]);
    float sum_s = 0.0f, sum_ss = 0.0f, sum_rs = 0.0f;
#pragma unroll
    for (int t = 0; t < 36; ++t) {
        const int xo = ORDER == 0 ? 2 * (t / 6) - 5 : sa_b_xoff(t), yo = ORDER == 0 ? 2 * (t % 6) - 5 : sa_b_yoff(t);
        if ((mask >> t) & 1ull) {
            const float xi = (float)(px + xo), yj = (float)(py + yo);
            const float bx = fmaf(g0, xi, g2), by = fmaf(g3, xi, g5), bz = fmaf(h[6], xi, h[8]);
            const float X = fmaf(g1, yj, bx), Y = fmaf(g4, yj, by), Z = fmaf(h[7], yj, bz);
            const float iz = rcp_approx(Z);
            const float s = fetch<U>(K, X * iz, Y * iz, layer);
            sum_s += s;
            sum_ss = fmaf(s, s, sum_ss);
            sum_rs = fmaf(rp.r[((xo + 5) / 2) * 6 + (yo + 5) / 2], s, sum_rs);
        }
    }
    const float mean_s = inv * sum_s, e_rs = inv * sum_rs;
    const float var_s = fmaf(inv, sum_ss, -__fmul_rn(mean_s, mean_s));
    if (var_r < 1e-5f || var_s < 1e-5f) return 2.0f;
    const float covar = fmaf(-mean_r, mean_s, e_rs);
    return fmaxf(0.0f, fminf(2.0f, fmaf(-covar, rsqrtf(var_r * var_s), 1.0f)));
}

// ComputeBilateralNCCOld with a label map, APD.cu:596-721
template <bool U>
__device__ __forceinline__ float ncc_old_sa(const PassK &K, const ViewK &vk, int px, int py, float3 m, const RefPatch &rp,
                                            const SaInfo &si) {
    const Homog Hm = make_homography(vk, m);
    const float *h = Hm.h;
    const float fxp = (float)px, fyp = (float)py;
    const float Z = __fadd_rn(__fmaf_rn(h[7], fyp, __fmul_rn(h[6], fxp)), h[8]);
    const float iz = rcp_approx(Z);
    const float ptx = __fmul_rn(__fadd_rn(__fmaf_rn(h[1], fyp, __fmul_rn(h[0], fxp)), h[2]), iz);
    const float pty = __fmul_rn(__fadd_rn(__fmaf_rn(h[4], fyp, __fmul_rn(h[3], fxp)), h[5]), iz);
    if (ptx >= (float)K.W || ptx < 0.0f || pty >= (float)K.H || pty < 0.0f) return 2.0f;
    // "const int center = pt.y * src_camera.width + pt.x": one FFMA and a truncation; the float can round up to W*H (clamped)
    const int c = min(__float2int_rz(__fmaf_rn(pty, (float)K.W, ptx)), K.W * K.H - 1);
    if (K.sa[c] == 0) return patch_ncc36<U>(K, Hm, vk.layer, px, py, rp);
    if (si.bmask == 0ull) return 2.0f;  // 0 * (1/0) = NaN everywhere: max(0, min(2, NaN)) = 2
    return patch_ncc36_masked<U, 1>(K, Hm, vk.layer, px, py, rp, si.bmask, si.b_mean, si.b_var, si.b_inv);
}

struct AnchorRefSa : AnchorRef {
    float inv[8];
    unsigned short tmask[8];
};

// label tests on an anchor cache whose taps (r), positions (a) are filled: anchors of another segment are dropped, taps of
// another segment leave the sums (APD.cu:493-497, 526-530); tap values come from ar.r, so the column pipeline can apply this to
// the cache it re-loads from global memory
__device__ __forceinline__ void anchor_ref_apply_labels(const PassK &K, AnchorRefSa &ar, const SaInfo &si) {
    const uint8_t *sa = K.sa;
    const int W = K.W, H = K.H, lab = si.label;
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
        const short2 a = ar.a[k];
        if (a.x == -1) continue;
        if (lab != 0 && sa[a.x + a.y * W] != lab) { ar.a[k].x = -1; continue; }
        float sr = 0.0f, srr = 0.0f;
        int t = 0, cnt = 0;
        unsigned tm = 0u;
#pragma unroll
        for (int i = -5; i <= 5; i += 5) {
#pragma unroll
            for (int j = -5; j <= 5; j += 5) {
                // the reference indexes the label map without a bounds test (APD.cu:527); clamped here
                const int x = clampi(a.x + i, 0, W - 1), y = clampi(a.y + j, 0, H - 1);
                if (lab == 0 || sa[y * W + x] == lab) {
                    const float r = ar.r[k * 9 + t];
                    tm |= 1u << t;
                    sr += r;
                    srr = fmaf(r, r, srr);
                    cnt++;
                }
                t++;
            }
        }
        ar.tmask[k] = (unsigned short)tm;
        ar.inv[k] = sa_inv_count(cnt, 9, 1.0f / 9.0f);
        ar.mean[k] = ar.inv[k] * sr;
        ar.var[k] = fmaf(ar.inv[k], srr, -__fmul_rn(ar.mean[k], ar.mean[k]));
    }
}

template <bool U>
__device__ __forceinline__ void load_anchor_ref_sa(const PassK &K, const short2 *anc, AnchorRefSa &ar, const SaInfo &si) {
    load_anchor_ref<U>(K, anc, ar);
    anchor_ref_apply_labels(K, ar, si);
}

template <bool U>
__device__ __forceinline__ float patch_ncc9_masked(const PassK &K, const Homog &Hm, int layer, int ax, int ay, const float *r9,
                                                   unsigned tmask, float mean_r, float var_r, float inv) {
    const float *h = Hm.h;
    float ss = 0.0f, sss = 0.0f, srs = 0.0f;
    int t = 0;
#pragma unroll
    for (int i = -5; i <= 5; i += 5) {
#pragma unroll
        for (int j = -5; j <= 5; j += 5) {
            if ((tmask >> t) & 1u) {
                const float fxp = (float)(ax + i), fyp = (float)(ay + j);
                const float Z = h[6] * fxp + h[7] * fyp + h[8];
                const float iz = rcp_approx(Z);
                const float X = (h[0] * fxp + h[1] * fyp + h[2]) * iz, Y = (h[3] * fxp + h[4] * fyp + h[5]) * iz;
                const float s = fetch<U>(K, X + 0.5f, Y + 0.5f, layer);
                ss += s; sss = fmaf(s, s, sss); srs = fmaf(r9[t], s, srs);
            }
            t++;
        }
    }
    const float mean_s = inv * ss, e_rs = inv * srs;
    const float var_s = fmaf(inv, sss, -__fmul_rn(mean_s, mean_s));
    if (var_r < 1e-5f || var_s < 1e-5f) return 2.0f;
    const float covar = fmaf(-mean_r, mean_s, e_rs);
    return fmaxf(0.0f, fminf(2.0f, fmaf(-covar, rsqrtf(var_r * var_s), 1.0f)));
}

// ComputeBilateralNCCNew with a label map, APD.cu:448-593
template <bool U>
__device__ __forceinline__ float ncc_new_sa(const PassK &K, const ViewK &vk, int view_bit, int px, int py, float3 m,
                                            const RefPatch &rp, const AnchorRefSa &ar, const SaInfo &si) {
    const Homog Hm = make_homography(vk, m);
    const float *h = Hm.h;
    const float fW = (float)K.W, fH = (float)K.H;
    {
        const float fxp = (float)px, fyp = (float)py;
        const float iz = rcp_approx(h[6] * fxp + h[7] * fyp + h[8]);
        const float ptx = (h[0] * fxp + h[1] * fyp + h[2]) * iz, pty = (h[3] * fxp + h[4] * fyp + h[5]) * iz;
        if (ptx >= fW || ptx < 0.0f || pty >= fH || pty < 0.0f) return 2.0f;
    }
    // anchor 0 = the pixel itself: no tap of its segment -> "continue" with center_cost still 0 (APD.cu:543-545)
    const float center_cost = si.nmask == 0ull ? 0.0f
                              : patch_ncc36_masked<U, 0>(K, Hm, vk.layer, px, py, rp, si.nmask, si.n_mean, si.n_var, si.n_inv);
    float sc[8];
    int ns = 0;
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
        const short2 a = ar.a[k];
        if (a.x == -1) continue;
        const float fxp = (float)a.x, fyp = (float)a.y;
        const float iz = rcp_approx(h[6] * fxp + h[7] * fyp + h[8]);
        const float ax = (h[0] * fxp + h[1] * fyp + h[2]) * iz, ay = (h[3] * fxp + h[4] * fyp + h[5]) * iz;
        if (ax < 0.0f || ay < 0.0f || ax >= fW || ay >= fH) {
            if ((K.sel[a.x + a.y * K.W] >> view_bit) & 1u) sc[ns++] = 2.0f;
            continue;
        }
        sc[ns++] = patch_ncc9_masked<U>(K, Hm, vk.layer, a.x, a.y, &ar.r[k * 9], ar.tmask[k], ar.mean[k], ar.var[k], ar.inv[k]);
    }
    if (ns == 0) return center_cost;
    float mx = -1e10f;
    for (int i = 0; i < ns; ++i) mx = fmaxf(mx, sc[i]);
    float sum = 0.0f, acc = 0.0f;
    float wts[8];
    for (int i = 0; i < ns; ++i) { wts[i] = __expf(sc[i] - mx); sum += wts[i]; }
    for (int i = 0; i < ns; ++i) acc += (wts[i] / sum) * sc[i];
    acc = fminf(acc, 2.0f);
    return 0.25f * center_cost + 0.75f * acc;
}

// compile-time switch used by the kernel bodies that exist in both flavours
template <bool SA> struct SaTypes { typedef SaNone Info; typedef AnchorRef Anchors; };
template <> struct SaTypes<true> { typedef SaInfo Info; typedef AnchorRefSa Anchors; };
template <bool SA>
__device__ __forceinline__ void load_sa(const PassK &K, int px, int py, const RefPatch &rp, typename SaTypes<SA>::Info &si) {
    if constexpr (SA) load_sa_info(K, px, py, rp, si);
}
template <bool U, bool SA>
__device__ __forceinline__ float ncc_old_x(const PassK &K, const ViewK &vk, int px, int py, float3 m, const RefPatch &rp,
                                           const typename SaTypes<SA>::Info &si) {
    if constexpr (SA) return ncc_old_sa<U>(K, vk, px, py, m, rp, si);
    else return ncc_old<U>(K, vk, px, py, m, rp);
}
template <bool U, bool SA>
__device__ __forceinline__ void load_anchor_ref_x(const PassK &K, const short2 *anc, typename SaTypes<SA>::Anchors &ar,
                                                  const typename SaTypes<SA>::Info &si) {
    if constexpr (SA) load_anchor_ref_sa<U>(K, anc, ar, si);
    else load_anchor_ref<U>(K, anc, ar);
}
template <bool U, bool SA>
__device__ __forceinline__ float ncc_new_x(const PassK &K, const ViewK &vk, int view_bit, int px, int py, float3 m, const RefPatch &rp,
                                           const typename SaTypes<SA>::Anchors &ar, const typename SaTypes<SA>::Info &si) {
    if constexpr (SA) return ncc_new_sa<U>(K, vk, view_bit, px, py, m, rp, ar, si);
    else return ncc_new<U>(K, vk, view_bit, px, py, m, rp, ar);
}

// ComputeGeomConsistencyCost, APD.cu:865-902, with the camera pair pre-composed:
//   forward:  x_s ~ depth * A p~ + b ;  backward:  x_r ~ d_s * Ai s~ + bi
__device__ __forceinline__ float geom_cost(const PassK &K, const ViewK &vk, int v /* 0-based source */, int px, int py,
                                           float4 plane) {
    const float depth = depth_from_plane(K, plane, px, py);
    const float fxp = (float)px, fyp = (float)py;
    // explicit intrinsics throughout: the truncation below turns a 1-ulp difference into another depth texel
    const float qx = __fadd_rn(__fmaf_rn(vk.A[1], fyp, __fmul_rn(vk.A[0], fxp)), vk.A[2]);
    const float qy = __fadd_rn(__fmaf_rn(vk.A[4], fyp, __fmul_rn(vk.A[3], fxp)), vk.A[5]);
    const float qz = __fadd_rn(__fmaf_rn(vk.A[7], fyp, __fmul_rn(vk.A[6], fxp)), vk.A[8]);
    const float Z = __fmaf_rn(depth, qz, vk.b[2]);
    const float sx = __fdividef(__fmaf_rn(depth, qx, vk.b[0]), Z), sy = __fdividef(__fmaf_rn(depth, qy, vk.b[1]), Z);
    // "(int)src_pt.x + 0.5f" through a clamped texture == clamped truncation (cvt.rzi saturates, NaN -> 0)
    const int ix = clampi(__float2int_rz(sx), 0, K.W - 1), iy = clampi(__float2int_rz(sy), 0, K.H - 1);
    const float sd = K.depth[(size_t)(v + 1) * K.W * K.H + (size_t)iy * K.W + ix];
    if (sd == 0.0f) return 3.0f;
    const float rx = __fadd_rn(__fmaf_rn(vk.Ai[1], sy, __fmul_rn(vk.Ai[0], sx)), vk.Ai[2]);
    const float ry = __fadd_rn(__fmaf_rn(vk.Ai[4], sy, __fmul_rn(vk.Ai[3], sx)), vk.Ai[5]);
    const float rz = __fadd_rn(__fmaf_rn(vk.Ai[7], sy, __fmul_rn(vk.Ai[6], sx)), vk.Ai[8]);
    const float Zr = __fmaf_rn(sd, rz, vk.bi[2]);
    const float bx = __fdividef(__fmaf_rn(sd, rx, vk.bi[0]), Zr), by = __fdividef(__fmaf_rn(sd, ry, vk.bi[1]), Zr);
    const float dc = __fsub_rn(fxp, bx), dr = __fsub_rn(fyp, by);
    return fminf(3.0f, sqrtf(__fmaf_rn(dr, dr, __fmul_rn(dc, dc))));
}

// packed view weights: 4 bits per view in a uint4
__device__ __forceinline__ uint32_t vw_get(const uint4 &w, int v) {
    const uint32_t word = (v < 8) ? w.x : (v < 16) ? w.y : (v < 24) ? w.z : w.w;
    return (word >> ((v & 7) * 4)) & 15u;
}
__device__ __forceinline__ void vw_inc(uint4 &w, int v) {
    const uint32_t inc = 1u << ((v & 7) * 4);
    if (v < 8) w.x += inc; else if (v < 16) w.y += inc; else if (v < 24) w.z += inc; else w.w += inc;
}

__device__ __forceinline__ void count_evals(const PassK &K, unsigned n_old, unsigned n_new, unsigned n_geom) {
    const unsigned m = __activemask();
    const unsigned a = __reduce_add_sync(m, n_old), b = __reduce_add_sync(m, n_new), c = __reduce_add_sync(m, n_geom);
    if ((int)(threadIdx.x & 31u) == __ffs(m) - 1) {
        if (a) atomicAdd(&K.counters[0], (unsigned long long)a);
        if (b) atomicAdd(&K.counters[1], (unsigned long long)b);
        if (c) atomicAdd(&K.counters[2], (unsigned long long)c);
    }
}

}  // namespace apde
