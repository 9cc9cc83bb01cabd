// apde_device.cuh -- device-side types, counter RNG and per-hypothesis cost functions (sm_100a).
//
// Design: one thread per pixel; the 36 reference-patch texels and their sums are loaded ONCE per
// pixel per kernel and stay in registers (the reference re-fetches them through the texture unit for every
// evaluation, APD.cu:632); the camera-pair part of the homography (R_rel, t_rel) is computed once per (ref, src) pair
// (the reference recomputes it per evaluation, APD.cu:336-362); source texels are gathered through ONE layered texture
// object (linear filter, clamp) holding every view of the pyramid level.
//
// ARITHMETIC CONTRACT.  Every floating-point operation that feeds a sample coordinate, a cost or a depth follows the
// reference BUILD (nvcc -O3 --use_fast_math, CMakeLists.txt:26) operation for operation: the same association, the same
// FMUL / FFMA / FADD split the compiler chose there (read from the SASS of the reference build, recorded in
// profiles/r02_reference_sass_arithmetic.md), MUFU.RCP / MUFU.SQRT where the reference build has them.  All of it is
// written with explicit intrinsics so that no inlining context can re-associate or re-contract it.  The reason: sample
// coordinates land on the texture unit's 1/256 weight grid, a 1-ulp difference in the homography flips buckets and moves a
// cost by ~1e-4; with the reference's own operation order the costs agree bit for bit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/apde.h"

namespace apde {

constexpr int kMaxSrc = APDE_MAX_IMAGES - 1;
constexpr int kPatch = 36;  // (2*5/2+1)^2 samples: strong_radius 5, strong_increment 2 (main.h:87-88)

// per (reference, source) pair constants
struct ViewK {
    float Rrel[9];  // R_s R_r^T            fp32, in the reference's operation order (APD.cu:348-356; host twin: pair_terms())
    float trel[3];  // R_s (C_r - C_s)      (APD.cu:338-343, 357-362)
    float K[9], R[9], t[3], c[3];  // the source camera at this level: the geometric cost and the confidence count go
                                   // through world coordinates exactly as the reference does (APD.cu:831-863)
    int layer;      // layer of this source view in the level's texture
    int pad_[2];    // 160 bytes: 16-byte aligned rows in shared memory
};

struct PassK {
    int W, H, N;  // N = number of source views
    int state, geom, impetus, use_apd, top_k, weak_peak_radius, rotate_time, max_iterations;
    float depth_min, depth_max, geom_factor, ransac_threshold;
    float fx, fy, cx, cy;  // reference intrinsics at this level (= Kr[0], Kr[4], Kr[2], Kr[5])
    float R[9];            // reference rotation
    float Kr[9], t[3], c[3];  // the rest of the reference camera (world route of the geometric cost)
    int ref_layer;
    uint32_t seed, stream;
    cudaTextureObject_t tex;  // layered images of the level (linear filter, clamp, unnormalised coordinates)
    int tile_shift;           // implicit checkerboard mapping: a warp = the 32 same-colour pixels of a (1 << tile_shift) x (64 >> tile_shift) tile
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

// single MUFU.SQRT (what sqrt() / sqrtf() compile to under --use_fast_math on sm_100)
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// a0*b0 + a1*b1 + a2*b2 as the reference build contracts every such sum: FMUL on the MIDDLE product, FFMA of the first,
// FFMA of the last (APD.cu:338-362, 841-843, 856-862, 222)
__device__ __forceinline__ float dot3_ref(float a0, float b0, float a1, float b1, float a2, float b2) {
    return __fmaf_rn(a2, b2, __fmaf_rn(a0, b0, __fmul_rn(a1, b1)));
}

// reference-camera plane helpers (APD.cu:190-240), zero-skew pinhole; operation order of the reference build
// ComputeDepthfromPlaneHypothesis, APD.cu:237-240:
//   -w*K0 / ((x-K2)*nx + (K0/K4)*(y-K5)*ny + K0*nz)  ==  (w * -K0) * RCP(FFMA(K0, nz, FFMA(x-K2, nx, ((K0*RCP(K4))*(y-K5))*ny)))
__device__ __forceinline__ float depth_from_plane(const PassK &K, float4 pl, int px, int py) {
    const float ty = __fmul_rn(__fmul_rn(__fmul_rn(K.fx, rcp_approx(K.fy)), __fsub_rn((float)py, K.cy)), pl.y);
    const float den = __fmaf_rn(K.fx, pl.z, __fmaf_rn(__fsub_rn((float)px, K.cx), pl.x, ty));
    return __fmul_rn(__fmul_rn(pl.w, -K.fx), rcp_approx(den));
}
// GetDistance2Origin + Get3DPoint, APD.cu:190-195, 218-223:  X0 = (depth*(x-K2))*RCP(K0), X1 = (depth*(y-K5))*RCP(K4)
__device__ __forceinline__ float distance_to_origin(const PassK &K, int px, int py, float depth, float4 n) {
    const float X0 = __fmul_rn(__fmul_rn(depth, __fsub_rn((float)px, K.cx)), rcp_approx(K.fx));
    const float X1 = __fmul_rn(__fmul_rn(depth, __fsub_rn((float)py, K.cy)), rcp_approx(K.fy));
    return -dot3_ref(X0, n.x, X1, n.y, depth, n.z);
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
    return make_float4(dot3_ref(K.R[0], p.x, K.R[3], p.y, K.R[6], p.z), dot3_ref(K.R[1], p.x, K.R[4], p.y, K.R[7], p.z),
                       dot3_ref(K.R[2], p.x, K.R[5], p.y, K.R[8], p.z), p.w);
}
__device__ __forceinline__ float4 normal_to_refcam(const PassK &K, float4 p) {  // TransformNormal2RefCam APD.cu:415
    return make_float4(dot3_ref(K.R[0], p.x, K.R[1], p.y, K.R[2], p.z), dot3_ref(K.R[3], p.x, K.R[4], p.y, K.R[5], p.z),
                       dot3_ref(K.R[6], p.x, K.R[7], p.y, K.R[8], p.z), p.w);
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
    rp.mean = __fmul_rn(inv, s);
    rp.var = __fmaf_rn(inv, ss, -__fmul_rn(rp.mean, rp.mean));
}

// The hypothesis-dependent part of ComputeHomography: the camera-frame normal and MUFU.RCP(w) ("x / w" under
// --use_fast_math).  Shared by every source view of the hypothesis.
struct PlaneM { float nx, ny, nz, rw; };
__device__ __forceinline__ PlaneM plane_row(const PassK &, float4 pl) {
    PlaneM m;
    m.nx = pl.x; m.ny = pl.y; m.nz = pl.z; m.rw = rcp_approx(pl.w);
    return m;
}

struct Homog { float h[9]; };

// ComputeHomography, APD.cu:364-393, operation for operation as the reference build executes it:
//   H_i   = FFMA(-(n_c * t_rel[r]), RCP(w), R_rel[i])                                     (:364-372)
//   tmp0  = H0 * RCP(K0)   tmp1 = H1 * RCP(K4)                                            (:375-376)
//   tmp2  = H2 + FFMA(H0 * -K2, RCP(K0), -((H1 * K5) * RCP(K4)))                          (:377)
//   H'0   = FFMA(Ks0, tmp0, Ks2 * tmp6) ...  H'3 = FFMA(Ks4, tmp3, Ks5 * tmp6) ...  H'6 = Ks8 * tmp6   (:385-393)
// (K = reference intrinsics, Ks = source intrinsics; R_rel, t_rel come per camera pair from pair_terms())
__device__ __forceinline__ Homog make_homography(const PassK &K, const ViewK &vk, const PlaneM &m) {
    const float rK0 = rcp_approx(K.fx), rK4 = rcp_approx(K.fy);
    float t[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float tr = vk.trel[r];
        const float H0 = __fmaf_rn(-__fmul_rn(m.nx, tr), m.rw, vk.Rrel[3 * r + 0]);
        const float H1 = __fmaf_rn(-__fmul_rn(m.ny, tr), m.rw, vk.Rrel[3 * r + 1]);
        const float H2 = __fmaf_rn(-__fmul_rn(m.nz, tr), m.rw, vk.Rrel[3 * r + 2]);
        t[3 * r + 0] = __fmul_rn(H0, rK0);
        t[3 * r + 1] = __fmul_rn(H1, rK4);
        t[3 * r + 2] = __fadd_rn(H2, __fmaf_rn(__fmul_rn(H0, -K.cx), rK0, -__fmul_rn(__fmul_rn(H1, K.cy), rK4)));
    }
    const float ks0 = vk.K[0], ks2 = vk.K[2], ks4 = vk.K[4], ks5 = vk.K[5], ks8 = vk.K[8];
    Homog Hm;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        Hm.h[c] = __fmaf_rn(ks0, t[c], __fmul_rn(ks2, t[6 + c]));
        Hm.h[3 + c] = __fmaf_rn(ks4, t[3 + c], __fmul_rn(ks5, t[6 + c]));
        Hm.h[6 + c] = __fmul_rn(ks8, t[6 + c]);
    }
    return Hm;
}

// ComputeCorrespondingPoint of a single point (APD.cu:396-403) as the reference build contracts it OUTSIDE the patch loops:
//   X = H2 + FFMA(H0, x, H1 * y)  (same for Y, Z);  pt = (X * RCP(Z), Y * RCP(Z))
__device__ __forceinline__ void project_point(const float *h, float x, float y, float &ptx, float &pty) {
    const float X = __fadd_rn(h[2], __fmaf_rn(h[0], x, __fmul_rn(h[1], y)));
    const float Y = __fadd_rn(h[5], __fmaf_rn(h[3], x, __fmul_rn(h[4], y)));
    const float Z = __fadd_rn(h[8], __fmaf_rn(h[6], x, __fmul_rn(h[7], y)));
    const float iz = rcp_approx(Z);
    ptx = __fmul_rn(X, iz);
    pty = __fmul_rn(Y, iz);
}
// ... and the texture coordinate of that point, "pt + 0.5f" contracted with the division: FFMA(X, RCP(Z), 0.5)
// (the quadrant walk of NCC-Old branch B, APD.cu:687-689)
__device__ __forceinline__ void sample_coord_point(const float *h, float x, float y, float &u, float &v) {
    const float X = __fadd_rn(h[2], __fmaf_rn(h[0], x, __fmul_rn(h[1], y)));
    const float Y = __fadd_rn(h[5], __fmaf_rn(h[3], x, __fmul_rn(h[4], y)));
    const float Z = __fadd_rn(h[8], __fmaf_rn(h[6], x, __fmul_rn(h[7], y)));
    const float iz = rcp_approx(Z);
    u = __fmaf_rn(X, iz, 0.5f);
    v = __fmaf_rn(Y, iz, 0.5f);
}
// Inside the patch loops (i outer over x, j inner over y; APD.cu:629-634, 523-533) the reference build hoists the x
// products out of the inner loop, which changes the contraction:  X = H2 + FFMA(H1, y, H0 * x);  u = FFMA(X, RCP(Z), 0.5)
struct RowX { float x0, x3, x6; };
__device__ __forceinline__ RowX sample_row(const float *h, float x) {
    RowX r;
    r.x0 = __fmul_rn(h[0], x); r.x3 = __fmul_rn(h[3], x); r.x6 = __fmul_rn(h[6], x);
    return r;
}
__device__ __forceinline__ void sample_coord_loop(const float *h, const RowX &r, float y, float &u, float &v) {
    const float X = __fadd_rn(h[2], __fmaf_rn(h[1], y, r.x0));
    const float Y = __fadd_rn(h[5], __fmaf_rn(h[4], y, r.x3));
    const float Z = __fadd_rn(h[8], __fmaf_rn(h[7], y, r.x6));
    const float iz = rcp_approx(Z);
    u = __fmaf_rn(X, iz, 0.5f);
    v = __fmaf_rn(Y, iz, 0.5f);
}

// NCC from the five sums, APD.cu:644-662 as built: means = inv * sums; var = FFMA(inv, sum_xx, -(mean * mean));
// covar = FFMA(-mean_r, mean_s, inv * sum_rs); cost = max(0, min(2, FFMA(-covar, RCP(SQRT(var_r * var_s)), 1)))
__device__ __forceinline__ float ncc_epilogue(float inv, float mean_r, float var_r, float sum_s, float sum_ss, float sum_rs) {
    const float mean_s = __fmul_rn(inv, sum_s), e_rs = __fmul_rn(inv, sum_rs);
    const float var_s = __fmaf_rn(inv, sum_ss, -__fmul_rn(mean_s, mean_s));
    if (var_r < 1e-5f || var_s < 1e-5f) return 2.0f;
    const float covar = __fmaf_rn(-mean_r, mean_s, e_rs);
    const float c = __fmaf_rn(-covar, rcp_approx(sqrt_approx(__fmul_rn(var_r, var_s))), 1.0f);
    return fmaxf(0.0f, fminf(c, 2.0f));
}

// NCC of the warped 6x6 patch against the register-resident reference patch.  APD.cu:622-662.
template <bool U>
__device__ __forceinline__ float patch_ncc36(const PassK &K, const Homog &Hm, int layer, int px, int py,
                                             const RefPatch &rp) {
    const float *h = Hm.h;
    // "if (var_ref < kMinVar || var_src < kMinVar) cost = cost_max" (APD.cu:655): a texture-less reference patch decides the
    // cost on its own, whatever the 36 source samples are
    if (rp.var < 1e-5f) return 2.0f;
    float sum_s = 0.0f, sum_ss = 0.0f, sum_rs = 0.0f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const RowX row = sample_row(h, (float)(px + 2 * i - 5));
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            float u, v;
            sample_coord_loop(h, row, (float)(py + 2 * j - 5), u, v);
            const float s = fetch<U>(K, u, v, layer);
            sum_s = __fadd_rn(sum_s, s);
            sum_ss = __fmaf_rn(s, s, sum_ss);
            sum_rs = __fmaf_rn(rp.r[i * 6 + j], s, sum_rs);
        }
    }
    return ncc_epilogue(1.0f / 36.0f, rp.mean, rp.var, sum_s, sum_ss, sum_rs);
}

// ComputeBilateralNCCOld, APD.cu:596-663 (branch A).  Only the patch centre is bounds-checked (quirk 8).
template <bool U>
__device__ __forceinline__ float ncc_old(const PassK &K, const ViewK &vk, int px, int py, const PlaneM &m, const RefPatch &rp) {
    const Homog Hm = make_homography(K, vk, m);
    float ptx, pty;
    project_point(Hm.h, (float)px, (float)py, ptx, pty);
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
                sr = __fadd_rn(sr, r);
                srr = __fmaf_rn(r, r, srr);
            }
        }
        const float inv = 1.0f / 9.0f;
        ar.mean[k] = __fmul_rn(inv, sr);
        ar.var[k] = __fmaf_rn(inv, srr, -__fmul_rn(ar.mean[k], ar.mean[k]));
    }
}

// NCC of one 3x3 anchor patch: only the source samples are gathered here
template <bool U>
__device__ __forceinline__ float patch_ncc9(const PassK &K, const Homog &Hm, int layer, int ax, int ay, const float *r9,
                                            float mean_r, float var_r) {
    const float *h = Hm.h;
    if (var_r < 1e-5f) return 2.0f;  // decided by the reference side alone (APD.cu:556)
    float ss = 0.0f, sss = 0.0f, srs = 0.0f;
    int t = 0;
#pragma unroll
    for (int i = -5; i <= 5; i += 5) {
        const RowX row = sample_row(h, (float)(ax + i));
#pragma unroll
        for (int j = -5; j <= 5; j += 5) {
            float u, v;
            sample_coord_loop(h, row, (float)(ay + j), u, v);
            const float s = fetch<U>(K, u, v, layer);
            ss = __fadd_rn(ss, s); sss = __fmaf_rn(s, s, sss); srs = __fmaf_rn(r9[t++], s, srs);
        }
    }
    return ncc_epilogue(1.0f / 9.0f, mean_r, var_r, ss, sss, srs);
}

// Softmax-weighted mean of the anchor costs + the focal mix with the centre cost, APD.cu:431-446, 572-587 as built:
//   e_i = EX2((c_i - max) * log2e); w_i = e_i * RCP(sum e); strong = FFMA chain of w_i * c_i; MIN(strong, 2);
//   "0.25 * center + 0.75 * strong" is evaluated in double and rounded once: DFMA(center, 0.25, strong * 0.75) ->
//   in float that is FFMA(0.75, strong, 0.25 * center) (0.25 * center is exact, the FFMA rounds the exact sum once)
__device__ __forceinline__ float focal_mix(float center_cost, const float *sc, int ns) {
    if (ns == 0) return center_cost;
    float mx = -1e10f;
    for (int i = 0; i < ns; ++i) if (sc[i] > mx) mx = sc[i];
    float sum = 0.0f, acc = 0.0f;
    float wts[8];
    for (int i = 0; i < ns; ++i) { wts[i] = __expf(__fsub_rn(sc[i], mx)); sum = __fadd_rn(sum, wts[i]); }
    const float rs = rcp_approx(sum);
    for (int i = 0; i < ns; ++i) acc = __fmaf_rn(__fmul_rn(wts[i], rs), sc[i], acc);
    acc = acc > 2.0f ? 2.0f : acc;
    return __fmaf_rn(0.75f, acc, __fmul_rn(0.25f, center_cost));
}

// ComputeBilateralNCCNew, APD.cu:448-593 (sa_mask == 0): centre patch + focal-weighted anchor patches.
template <bool U>
__device__ __forceinline__ float ncc_new(const PassK &K, const ViewK &vk, int view_bit, int px, int py, const PlaneM &m,
                                         const RefPatch &rp, const AnchorRef &ar) {
    const Homog Hm = make_homography(K, vk, m);
    const float *h = Hm.h;
    const float fW = (float)K.W, fH = (float)K.H;
    {
        float ptx, pty;
        project_point(h, (float)px, (float)py, ptx, pty);
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
        float ax, ay;
        project_point(h, (float)a.x, (float)a.y, ax, ay);
        if (ax < 0.0f || ay < 0.0f || ax >= fW || ay >= fH) {
            if ((K.sel[a.x + a.y * K.W] >> view_bit) & 1u) sc[ns++] = 2.0f;
            continue;
        }
        sc[ns++] = patch_ncc9<U>(K, Hm, vk.layer, a.x, a.y, &ar.r[k * 9], ar.mean[k], ar.var[k]);
    }
    return focal_mix(center_cost, sc, ns);
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
                        s = __fadd_rn(s, v);
                        ss = __fmaf_rn(v, v, ss);
                        cnt++;
                    }
                }
            }
        }
        si.bmask = bm;
        si.b_inv = sa_inv_count(cnt, 36, 1.0f / 36.0f);
        si.b_mean = __fmul_rn(si.b_inv, s);
        si.b_var = __fmaf_rn(si.b_inv, ss, -__fmul_rn(si.b_mean, si.b_mean));
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
                    s = __fadd_rn(s, v);
                    ss = __fmaf_rn(v, v, ss);
                    cnt++;
                }
            }
        }
        si.nmask = nm;
        si.n_inv = sa_inv_count(cnt, 36, 1.0f / 36.0f);
        si.n_mean = __fmul_rn(si.n_inv, s);
        si.n_var = __fmaf_rn(si.n_inv, ss, -__fmul_rn(si.n_mean, si.n_mean));
    }
}

// masked twin of patch_ncc36: ORDER 0 = branch A order (i outer, j inner: the loops of APD.cu:523-533 with their hoisted
// x products), ORDER 1 = the quadrant walk of APD.cu:665-698 (single points).  A full mask in order 0 gives patch_ncc36.
template <bool U, int ORDER>
__device__ __forceinline__ float patch_ncc36_masked(const PassK &K, const Homog &Hm, int layer, int px, int py, const RefPatch &rp,
                                                    unsigned long long mask, float mean_r, float var_r, float inv) {
    const float *h = Hm.h;
    if (var_r < 1e-5f) return 2.0f;  // decided by the reference side alone
    float sum_s = 0.0f, sum_ss = 0.0f, sum_rs = 0.0f;
#pragma unroll
    for (int t = 0; t < 36; ++t) {
        const int xo = ORDER == 0 ? 2 * (t / 6) - 5 : sa_b_xoff(t), yo = ORDER == 0 ? 2 * (t % 6) - 5 : sa_b_yoff(t);
        if ((mask >> t) & 1ull) {
            float u, v;
            if (ORDER == 0) sample_coord_loop(h, sample_row(h, (float)(px + xo)), (float)(py + yo), u, v);
            else sample_coord_point(h, (float)(px + xo), (float)(py + yo), u, v);
            const float s = fetch<U>(K, u, v, layer);
            sum_s = __fadd_rn(sum_s, s);
            sum_ss = __fmaf_rn(s, s, sum_ss);
            sum_rs = __fmaf_rn(rp.r[((xo + 5) / 2) * 6 + (yo + 5) / 2], s, sum_rs);
        }
    }
    return ncc_epilogue(inv, mean_r, var_r, sum_s, sum_ss, sum_rs);
}

// ComputeBilateralNCCOld with a label map, APD.cu:596-721
template <bool U>
__device__ __forceinline__ float ncc_old_sa(const PassK &K, const ViewK &vk, int px, int py, const PlaneM &m, const RefPatch &rp,
                                            const SaInfo &si) {
    const Homog Hm = make_homography(K, vk, m);
    float ptx, pty;
    project_point(Hm.h, (float)px, (float)py, ptx, pty);
    if (ptx >= (float)K.W || ptx < 0.0f || pty >= (float)K.H || pty < 0.0f) return 2.0f;
    // "const int center = pt.y * src_camera.width + pt.x": one FFMA of the FLOAT coordinates and a truncation.  For points in
    // the last source row the index runs past the map (up to W - 1 elements): the reference reads whatever follows its
    // cudaMalloc'ed buffer there (undefined).  Defined here as label 0 -> branch A (what the reference golden vectors show).
    const int c = __float2int_rz(__fmaf_rn(pty, (float)K.W, ptx));
    if (c >= K.W * K.H || K.sa[c] == 0) return patch_ncc36<U>(K, Hm, vk.layer, px, py, rp);
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
                    sr = __fadd_rn(sr, r);
                    srr = __fmaf_rn(r, r, srr);
                    cnt++;
                }
                t++;
            }
        }
        ar.tmask[k] = (unsigned short)tm;
        ar.inv[k] = sa_inv_count(cnt, 9, 1.0f / 9.0f);
        ar.mean[k] = __fmul_rn(ar.inv[k], sr);
        ar.var[k] = __fmaf_rn(ar.inv[k], srr, -__fmul_rn(ar.mean[k], ar.mean[k]));
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
    if (var_r < 1e-5f) return 2.0f;  // decided by the reference side alone
    float ss = 0.0f, sss = 0.0f, srs = 0.0f;
    int t = 0;
#pragma unroll
    for (int i = -5; i <= 5; i += 5) {
        const RowX row = sample_row(h, (float)(ax + i));
#pragma unroll
        for (int j = -5; j <= 5; j += 5) {
            if ((tmask >> t) & 1u) {
                float u, v;
                sample_coord_loop(h, row, (float)(ay + j), u, v);
                const float s = fetch<U>(K, u, v, layer);
                ss = __fadd_rn(ss, s); sss = __fmaf_rn(s, s, sss); srs = __fmaf_rn(r9[t], s, srs);
            }
            t++;
        }
    }
    return ncc_epilogue(inv, mean_r, var_r, ss, sss, srs);
}

// ComputeBilateralNCCNew with a label map, APD.cu:448-593
template <bool U>
__device__ __forceinline__ float ncc_new_sa(const PassK &K, const ViewK &vk, int view_bit, int px, int py, const PlaneM &m,
                                            const RefPatch &rp, const AnchorRefSa &ar, const SaInfo &si) {
    const Homog Hm = make_homography(K, vk, m);
    const float *h = Hm.h;
    const float fW = (float)K.W, fH = (float)K.H;
    {
        float ptx, pty;
        project_point(h, (float)px, (float)py, ptx, pty);
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
        float ax, ay;
        project_point(h, (float)a.x, (float)a.y, ax, ay);
        if (ax < 0.0f || ay < 0.0f || ax >= fW || ay >= fH) {
            if ((K.sel[a.x + a.y * K.W] >> view_bit) & 1u) sc[ns++] = 2.0f;
            continue;
        }
        sc[ns++] = patch_ncc9_masked<U>(K, Hm, vk.layer, a.x, a.y, &ar.r[k * 9], ar.tmask[k], ar.mean[k], ar.var[k], ar.inv[k]);
    }
    return focal_mix(center_cost, sc, ns);
}

// compile-time switch used by the kernel bodies that exist in both flavours
template <bool SA> struct SaTypes { typedef SaNone Info; typedef AnchorRef Anchors; };
template <> struct SaTypes<true> { typedef SaInfo Info; typedef AnchorRefSa Anchors; };
template <bool SA>
__device__ __forceinline__ void load_sa(const PassK &K, int px, int py, const RefPatch &rp, typename SaTypes<SA>::Info &si) {
    if constexpr (SA) load_sa_info(K, px, py, rp, si);
}
template <bool U, bool SA>
__device__ __forceinline__ float ncc_old_x(const PassK &K, const ViewK &vk, int px, int py, const PlaneM &m, const RefPatch &rp,
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
__device__ __forceinline__ float ncc_new_x(const PassK &K, const ViewK &vk, int view_bit, int px, int py, const PlaneM &m, const RefPatch &rp,
                                           const typename SaTypes<SA>::Anchors &ar, const typename SaTypes<SA>::Info &si) {
    if constexpr (SA) return ncc_new_sa<U>(K, vk, view_bit, px, py, m, rp, ar, si);
    else return ncc_new<U>(K, vk, view_bit, px, py, m, rp, ar);
}

// ------------------------------------------------------------------------------------------------ world route (APD.cu:831-863)
// Get3DPointonWorld_cu as built:  X = ((x - K2) * depth) * RCP(K0), Y = ((y - K5) * depth) * RCP(K4), Z = depth;
// world = R^T (X, Y, Z) + c with every row a dot3_ref
__device__ __forceinline__ float3 point_to_world(const float *Kc, const float *R, const float *c, float x, float y, float depth) {
    const float X = __fmul_rn(__fmul_rn(__fsub_rn(x, Kc[2]), depth), rcp_approx(Kc[0]));
    const float Y = __fmul_rn(__fmul_rn(__fsub_rn(y, Kc[5]), depth), rcp_approx(Kc[4]));
    return make_float3(__fadd_rn(dot3_ref(R[0], X, R[3], Y, R[6], depth), c[0]),
                       __fadd_rn(dot3_ref(R[1], X, R[4], Y, R[7], depth), c[1]),
                       __fadd_rn(dot3_ref(R[2], X, R[5], Y, R[8], depth), c[2]));
}
// ProjectonCamera_cu as built:  tmp = R P + t (dot3_ref + FADD);  depth = dot3_ref(K6.., tmp);  point = dot3_ref(K0.., tmp) * RCP(depth).
// Returns the two numerators and RCP(depth) so that callers can contract as the reference build does.
struct Proj { float nx, ny, d, rd; };
__device__ __forceinline__ Proj project_to_camera(const float *Kc, const float *R, const float *t, float3 P) {
    const float tx = __fadd_rn(dot3_ref(R[0], P.x, R[1], P.y, R[2], P.z), t[0]);
    const float ty = __fadd_rn(dot3_ref(R[3], P.x, R[4], P.y, R[5], P.z), t[1]);
    const float tz = __fadd_rn(dot3_ref(R[6], P.x, R[7], P.y, R[8], P.z), t[2]);
    Proj q;
    q.d = dot3_ref(Kc[6], tx, Kc[7], ty, Kc[8], tz);
    q.rd = rcp_approx(q.d);
    q.nx = dot3_ref(Kc[0], tx, Kc[1], ty, Kc[2], tz);
    q.ny = dot3_ref(Kc[3], tx, Kc[4], ty, Kc[5], tz);
    return q;
}
// depth-map texel under a projected point: "tex2D(depth_image, (int)x + 0.5f, (int)y + 0.5f)" through a clamped linear
// texture == the texel at the clamped truncation (cvt.rzi saturates, NaN -> 0)
__device__ __forceinline__ float source_depth(const PassK &K, int v, float sx, float sy) {
    const int ix = clampi(__float2int_rz(sx), 0, K.W - 1), iy = clampi(__float2int_rz(sy), 0, K.H - 1);
    return K.depth[(size_t)(v + 1) * K.W * K.H + (size_t)iy * K.W + ix];
}

// ComputeGeomConsistencyCost, APD.cu:865-902: forward through world coordinates into the source view, the source depth
// texel, back into the reference view; diff = FFMA(-numerator, RCP(depth), p); cost = min(3, SQRT(FFMA(dc, dc, dr * dr)))
__device__ __forceinline__ float geom_cost(const PassK &K, const ViewK &vk, int v /* 0-based source */, int px, int py,
                                           float4 plane) {
    const float depth = depth_from_plane(K, plane, px, py);
    const float fxp = (float)px, fyp = (float)py;
    const float3 Pw = point_to_world(K.Kr, K.R, K.c, fxp, fyp, depth);
    const Proj s = project_to_camera(vk.K, vk.R, vk.t, Pw);
    const float sx = __fmul_rn(s.nx, s.rd), sy = __fmul_rn(s.ny, s.rd);
    const float sd = source_depth(K, v, sx, sy);
    if (sd == 0.0f) return 3.0f;
    const float3 Ps = point_to_world(vk.K, vk.R, vk.c, sx, sy, sd);
    const Proj b = project_to_camera(K.Kr, K.R, K.t, Ps);
    const float dc = __fmaf_rn(-b.nx, b.rd, fxp), dr = __fmaf_rn(-b.ny, b.rd, fyp);
    return fminf(sqrt_approx(__fmaf_rn(dc, dc, __fmul_rn(dr, dr))), 3.0f);
}

// baseline of a source view as DepthToWeak / LocalRefine compute it (APD.cu:2142-2147): float differences of the camera
// centres, FFMA(d2, d2, FFMA(d0, d0, d1 * d1)), MUFU.SQRT
__device__ __forceinline__ float view_baseline(const PassK &K, const ViewK &vk) {
    const float d0 = __fsub_rn(K.c[0], vk.c[0]), d1 = __fsub_rn(K.c[1], vk.c[1]), d2 = __fsub_rn(K.c[2], vk.c[2]);
    return sqrt_approx(dot3_ref(d0, d0, d1, d1, d2, d2));
}

// The disparity sweep of DepthToWeak / LocalRefine (APD.cu:2155-2165, 2399-2407) as built:
//   base_line = sum * RCP(float(valid_src));  fb = base_line * K0;  disp = fb * RCP(origin_depth);
//   p_depth(pd) = fb * RCP(disp + float(pd))
struct SweepDepths {
    float fb, disp;
    __device__ __forceinline__ float depth(int pd) const { return __fmul_rn(fb, rcp_approx(__fadd_rn(disp, (float)pd))); }
};
__device__ __forceinline__ SweepDepths sweep_depths(const PassK &K, float base_sum, int valid_src, float origin_depth) {
    SweepDepths s;
    s.fb = __fmul_rn(__fmul_rn(rcp_approx((float)valid_src), base_sum), K.fx);
    s.disp = __fmul_rn(s.fb, rcp_approx(origin_depth));
    return s;
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

// counters[0..2] = NCC-Old, NCC-New (centre patches), geometric evaluations; counters[3] = 3x3 anchor patches SAMPLED by the
// anchor-sorted weak pipeline (9 samples each; the real sample count of the deformable cost: 36 * [1] + 9 * [3])
__device__ __forceinline__ void count_evals(const PassK &K, unsigned n_old, unsigned n_new, unsigned n_geom, unsigned n_anchor = 0) {
    const unsigned m = __activemask();
    const unsigned a = __reduce_add_sync(m, n_old), b = __reduce_add_sync(m, n_new), c = __reduce_add_sync(m, n_geom);
    const unsigned d = __reduce_add_sync(m, n_anchor);
    if ((int)(threadIdx.x & 31u) == __ffs(m) - 1) {
        if (a) atomicAdd(&K.counters[0], (unsigned long long)a);
        if (b) atomicAdd(&K.counters[1], (unsigned long long)b);
        if (c) atomicAdd(&K.counters[2], (unsigned long long)c);
        if (d) atomicAdd(&K.counters[3], (unsigned long long)d);
    }
}

}  // namespace apde
