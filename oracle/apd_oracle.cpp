/*
 * apd_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see apd_oracle.h).
 *
 * Plain fp32 restatement of the reference algorithm, one function per reference function, each citing the
 * reference file:line it follows.  Compiled with -ffp-contract=off so every operation rounds like the
 * source says.  Nothing in the product (apde_mvs_b200/) links, imports or executes this file.
 *
 * ARITHMETIC.  The reference is built with nvcc -O3 --use_fast_math (CMakeLists.txt:26): the compiler contracts
 * multiply-adds into FFMA and turns every division / sqrt into MUFU.RCP / MUFU.SQRT.  Which product of "a*b + c*d + e*f"
 * is rounded on its own decides sample coordinates at the 1-ulp level, and those land on the texture unit's 1/256 weight
 * grid -- so the restatement follows the reference BUILD operation for operation (pattern read from the SASS of
 * oracle/_ref/libapd_ref.so, profiles/r02_reference_sass_arithmetic.md): std::fmaf where the build has an FFMA, a separately
 * rounded product where it has an FMUL.  The only approximation left: MUFU.RCP / MUFU.SQRT / MUFU.EX2 are hardware tables
 * (<= 1 ulp off the correctly rounded value); the oracle uses the correctly rounded 1/x, sqrt, exp2 instead (mufu_rcp,
 * mufu_sqrt, mufu_ex2 below).  That residual is what the golden-vector tests bound.
 */
#include "apd_oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct F2 { float x, y; };
struct F3 { float x, y, z; };
struct F4 { float x, y, z, w; };
struct S2 { int16_t x, y; };
struct I2 { int x, y; };

constexpr float kPi = 3.14159265358979323846f;
constexpr double kPiD = 3.14159265358979323846;

/* OpenCV's MIN/MAX macros as used by the reference (NaN behaviour matters, APD.cu:2185) */
#define ORC_MIN(a, b) ((a) > (b) ? (b) : (a))
#define ORC_MAX(a, b) ((a) < (b) ? (b) : (a))

/* ------------------------------------------------------------------------------------------------
 * RNG: Philox4x32-10, key = (seed, stream), counter = (pixel, site, block, 0).
 * Replaces curand XORWOW (APD.cu:904-917); uniform() keeps curand_uniform's (0,1] mapping.
 * ---------------------------------------------------------------------------------------------- */
inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c[0];
        const uint64_t p1 = (uint64_t)M1 * c[2];
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        const uint32_t n0 = hi1 ^ c[1] ^ k0;
        const uint32_t n1 = lo1;
        const uint32_t n2 = hi0 ^ c[3] ^ k1;
        const uint32_t n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += W0; k1 += W1;
    }
}

enum Site : uint32_t {
    SITE_ANCHOR = 1,
    SITE_INIT = 2,
    SITE_STRONG = 16,  /* + iter */
    SITE_FIT = 48,     /* + iter */
    SITE_WEAK = 80,    /* + iter */
};

struct Rng {
    uint32_t k0, k1, pixel, site, n;
    uint32_t buf[4];
    Rng(uint32_t seed, uint32_t stream, uint32_t pixel_, uint32_t site_)
        : k0(seed), k1(stream), pixel(pixel_), site(site_), n(0) {}
    uint32_t next() {
        if ((n & 3u) == 0u) {
            buf[0] = pixel; buf[1] = site; buf[2] = n >> 2; buf[3] = 0u;
            philox4x32_10(buf, k0, k1);
        }
        return buf[(n++) & 3u];
    }
    /* (0,1], like curand_uniform: x * 2^-32 + 2^-33 */
    float uniform() { return std::fmaf((float)next(), 2.3283064365386963e-10f, 1.1641532182693481e-10f); }
};

/* ------------------------------------------------------------------------------------------------
 * texture semantics (APD.cpp:699-706): float32, linear filter, unnormalised, "wrap" == clamp.
 * tex_mode 1 emulates the texture unit's 1.8 fixed-point interpolation weights.
 * ---------------------------------------------------------------------------------------------- */
inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

inline float tex_linear(const float *T, int W, int H, float x, float y, int mode) {
    float xB = x - 0.5f, yB = y - 0.5f;
    if (!(xB == xB)) xB = 0.0f;
    if (!(yB == yB)) yB = 0.0f;
    xB = std::min(std::max(xB, -1.0f), (float)W);
    yB = std::min(std::max(yB, -1.0f), (float)H);
    const float fi = std::floor(xB), fj = std::floor(yB);
    const float a = xB - fi, b = yB - fj;
    const int i0 = clampi((int)fi, 0, W - 1), i1 = clampi((int)fi + 1, 0, W - 1);
    const int j0 = clampi((int)fj, 0, H - 1), j1 = clampi((int)fj + 1, 0, H - 1);
    const double t00 = T[j0 * (size_t)W + i0], t10 = T[j0 * (size_t)W + i1];
    const double t01 = T[j1 * (size_t)W + i0], t11 = T[j1 * (size_t)W + i1];
    if (mode == 1) {
        /* B200 texture unit, measured (tools/diag_tex.py, profiles/r01_texture_filter_model.md): 1.8 fixed-point
         * fractions A, B (round half up) and FOUR integer weights that sum to 256 with consistent marginals:
         * W11 = round(A*B/256), W10 = A - W11, W01 = B - W11, W00 = 256 - A - B + W11.  Bit-exact on 1e5 probes. */
        const double A = std::floor((double)a * 256.0 + 0.5), B = std::floor((double)b * 256.0 + 0.5);
        const double W11 = std::floor(A * B / 256.0 + 0.5), W10 = A - W11, W01 = B - W11, W00 = 256.0 - A - B + W11;
        return (float)((W00 * t00 + W10 * t10 + W01 * t01 + W11 * t11) / 256.0);
    }
    const double da = a, db = b;
    return (float)((1.0 - da) * (1.0 - db) * t00 + da * (1.0 - db) * t10 + (1.0 - da) * db * t01 + da * db * t11);
}

/* texel-centre fetch (x+0.5, y+0.5): exact texel, clamped */
inline float tex_point(const float *T, int W, int H, int x, int y) {
    return T[clampi(y, 0, H - 1) * (size_t)W + clampi(x, 0, W - 1)];
}

/* float -> int with CUDA cvt.rzi semantics (saturating, NaN -> 0), used by "(int)src_pt.x" APD.cu:885 */
inline int cuda_f2i(float v) {
    if (!(v == v)) return 0;
    if (v >= 2147483648.0f) return INT32_MAX;
    if (v <= -2147483648.0f) return INT32_MIN;
    return (int)v;
}

/* ------------------------------------------------------------------------------------------------
 * small helpers (APD.cu:3-38, 60-86, 157-240)
 * ---------------------------------------------------------------------------------------------- */
inline void sort_small(float *d, int n) { /* APD.cu:3-12 */
    int j;
    for (int i = 1; i < n; i++) {
        float tmp = d[i];
        for (j = i; j >= 1 && tmp < d[j - 1]; j--) d[j] = d[j - 1];
        d[j] = tmp;
    }
}

inline int find_min_cost_index(const float *costs, int n) { /* APD.cu:60-71, ties -> last */
    float min_cost = costs[0];
    int idx_min = 0;
    for (int idx = 1; idx < n; ++idx) {
        if (costs[idx] <= min_cost) { min_cost = costs[idx]; idx_min = idx; }
    }
    return idx_min;
}

inline int is_set(uint32_t v, int n) { return (v >> n) & 1u; }

/* hardware special-function units of the reference build, approximated by the correctly rounded value (<= 1 ulp) */
inline float mufu_rcp(float x) { return 1.0f / x; }
inline float mufu_sqrt(float x) { return std::sqrt(x); }
inline float mufu_ex2(float x) { return std::exp2(x); }
/* a0*b0 + a1*b1 + a2*b2 as the reference build contracts every such sum: FMUL of the MIDDLE product, FFMA of the first,
 * FFMA of the last (APD.cu:338-362, 408-422, 841-843, 856-862, 222) */
inline float dot3_ref(float a0, float b0, float a1, float b1, float a2, float b2) {
    const float mid = a1 * b1;
    return std::fmaf(a2, b2, std::fmaf(a0, b0, mid));
}

inline void normalize3(F4 *v) { /* APD.cu:157-164 (rsqrtf) */
    const float n2 = v->x * v->x + v->y * v->y + v->z * v->z;
    const float inv = 1.0f / std::sqrt(n2);
    v->x *= inv; v->y *= inv; v->z *= inv;
}

inline void normalize2(F2 *v) { /* APD.cu:166-172 */
    const float n2 = v->x * v->x + v->y * v->y;
    const float inv = 1.0f / std::sqrt(n2);
    v->x *= inv; v->y *= inv;
}

inline void pdf_to_cdf(float *probs, int n) { /* APD.cu:174-188 */
    float s = 0.0f;
    for (int i = 0; i < n; ++i) s += probs[i];
    const float inv = 1.0f / s;
    float cum = 0.0f;
    for (int i = 0; i < n; ++i) { const float p = probs[i] * inv; cum += p; probs[i] = cum; }
}

inline void get_3d_point(const orc_camera &cam, float px, float py, float depth, float *X) { /* APD.cu:190-202 */
    X[0] = (depth * (px - cam.K[2])) * mufu_rcp(cam.K[0]); /* as built: FMUL, FMUL by MUFU.RCP(K0) */
    X[1] = (depth * (py - cam.K[5])) * mufu_rcp(cam.K[4]);
    X[2] = depth;
}

inline F4 view_direction(const orc_camera &cam, I2 p, float depth) { /* APD.cu:204-216 */
    float X[3];
    get_3d_point(cam, (float)p.x, (float)p.y, depth, X);
    const float norm = std::sqrt(X[0] * X[0] + X[1] * X[1] + X[2] * X[2]);
    return F4{X[0] / norm, X[1] / norm, X[2] / norm, 0.0f};
}

inline float distance_to_origin(const orc_camera &cam, I2 p, float depth, F4 n) { /* APD.cu:218-223 */
    float X[3];
    get_3d_point(cam, (float)p.x, (float)p.y, depth, X);
    return -dot3_ref(X[0], n.x, X[1], n.y, X[2], n.z);
}

inline float depth_from_plane(const orc_camera &cam, F4 pl, I2 p) { /* APD.cu:237-240 */
    /* as built: (w * -K0) * RCP(FFMA(K0, nz, FFMA(x - K2, nx, ((K0 * RCP(K4)) * (y - K5)) * ny))) */
    const float ty = ((cam.K[0] * mufu_rcp(cam.K[4])) * ((float)p.y - cam.K[5])) * pl.y;
    const float den = std::fmaf(cam.K[0], pl.z, std::fmaf((float)p.x - cam.K[2], pl.x, ty));
    return (pl.w * -cam.K[0]) * mufu_rcp(den);
}

inline F4 random_normal(const orc_camera &cam, I2 p, Rng &rng, float depth) { /* APD.cu:242-268 */
    float q1 = 1.0f, q2 = 1.0f, s = 2.0f;
    while (s >= 1.0f) {
        q1 = 2.0f * rng.uniform() - 1.0f;
        q2 = 2.0f * rng.uniform() - 1.0f;
        s = q1 * q1 + q2 * q2;
    }
    const float sq = std::sqrt(1.0f - s);
    F4 n{2.0f * q1 * sq, 2.0f * q2 * sq, 1.0f - 2.0f * s, 0.0f};
    const F4 vd = view_direction(cam, p, depth);
    const float dp = n.x * vd.x + n.y * vd.y + n.z * vd.z;
    if (dp > 0.0f) { n.x = -n.x; n.y = -n.y; n.z = -n.z; }
    normalize3(&n);
    return n;
}

inline F4 perturbed_normal(const orc_camera &cam, I2 p, F4 normal, Rng &rng, float perturbation) { /* APD.cu:270-305 */
    const F4 vd = view_direction(cam, p, 1.0f);
    const float a1 = (rng.uniform() - 0.5f) * perturbation;
    const float a2 = (rng.uniform() - 0.5f) * perturbation;
    const float a3 = (rng.uniform() - 0.5f) * perturbation;
    const float s1 = std::sin(a1), s2 = std::sin(a2), s3 = std::sin(a3);
    const float c1 = std::cos(a1), c2 = std::cos(a2), c3 = std::cos(a3);
    float R[9];
    R[0] = c2 * c3;
    R[1] = c3 * s1 * s2 - c1 * s3;
    R[2] = s1 * s3 + c1 * c3 * s2;
    R[3] = c2 * s3;
    R[4] = c1 * c3 + s1 * s2 * s3;
    R[5] = c1 * s2 * s3 - c3 * s1;
    R[6] = -s2;
    R[7] = c2 * s1;
    R[8] = c1 * c2;
    F4 np;
    np.x = R[0] * normal.x + R[1] * normal.y + R[2] * normal.z;
    np.y = R[3] * normal.x + R[4] * normal.y + R[5] * normal.z;
    np.z = R[6] * normal.x + R[7] * normal.y + R[8] * normal.z;
    np.w = 0.0f; /* Mat33DotVec3 leaves w untouched (uninitialised in the reference); callers overwrite it */
    if (np.x * vd.x + np.y * vd.y + np.z * vd.z >= 0.0f) np = normal;
    normalize3(&np);
    return np;
}

inline F4 random_plane(const orc_camera &cam, I2 p, Rng &rng, float dmin, float dmax) { /* APD.cu:307-313 */
    const float depth = rng.uniform() * (dmax - dmin) + dmin;
    F4 pl = random_normal(cam, p, rng, depth);
    pl.w = distance_to_origin(cam, p, depth, pl);
    return pl;
}

inline F4 normal_to_world(const orc_camera &cam, F4 pl) { /* TransformNormal APD.cu:405-413 */
    F4 r;
    r.x = dot3_ref(cam.R[0], pl.x, cam.R[3], pl.y, cam.R[6], pl.z);
    r.y = dot3_ref(cam.R[1], pl.x, cam.R[4], pl.y, cam.R[7], pl.z);
    r.z = dot3_ref(cam.R[2], pl.x, cam.R[5], pl.y, cam.R[8], pl.z);
    r.w = pl.w;
    return r;
}

inline F4 normal_to_refcam(const orc_camera &cam, F4 pl) { /* TransformNormal2RefCam APD.cu:415-423 */
    F4 r;
    r.x = dot3_ref(cam.R[0], pl.x, cam.R[1], pl.y, cam.R[2], pl.z);
    r.y = dot3_ref(cam.R[3], pl.x, cam.R[4], pl.y, cam.R[5], pl.z);
    r.z = dot3_ref(cam.R[6], pl.x, cam.R[7], pl.y, cam.R[8], pl.z);
    r.w = pl.w;
    return r;
}

/* APD.cu:334-394, as built (see the header): every three-term sum is a dot3_ref; C_rel = ref_C - src_C;
 *   H_i  = FFMA(-(t_rel[r] * n_c), RCP(w), R_rel[i]);  tmp0 = H0 * RCP(K0);  tmp1 = H1 * RCP(K4);
 *   tmp2 = H2 + FFMA(H0 * -K2, RCP(K0), -((H1 * K5) * RCP(K4)));  H'0 = FFMA(Ks0, tmp0, Ks2 * tmp6) ...  H'6 = Ks8 * tmp6 */
void homography(const orc_camera &rc, const orc_camera &sc, F4 pl, float *H) {
    float Sr[3], Ss[3]; /* -ref_C, -src_C */
    for (int j = 0; j < 3; ++j) {
        Sr[j] = dot3_ref(rc.R[0 + j], rc.t[0], rc.R[3 + j], rc.t[1], rc.R[6 + j], rc.t[2]);
        Ss[j] = dot3_ref(sc.R[0 + j], sc.t[0], sc.R[3 + j], sc.t[1], sc.R[6 + j], sc.t[2]);
    }
    float Rr[9], Cr[3], tr[3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            Rr[3 * i + j] = dot3_ref(sc.R[3 * i], rc.R[3 * j], sc.R[3 * i + 1], rc.R[3 * j + 1], sc.R[3 * i + 2], rc.R[3 * j + 2]);
    for (int j = 0; j < 3; ++j) Cr[j] = Ss[j] - Sr[j]; /* (-Sr) - (-Ss) */
    for (int i = 0; i < 3; ++i) tr[i] = dot3_ref(sc.R[3 * i], Cr[0], sc.R[3 * i + 1], Cr[1], sc.R[3 * i + 2], Cr[2]);

    const float rw = mufu_rcp(pl.w), rK0 = mufu_rcp(rc.K[0]), rK4 = mufu_rcp(rc.K[4]);
    float tmp[9];
    for (int r = 0; r < 3; ++r) {
        const float q0 = pl.x * tr[r], q1 = pl.y * tr[r], q2 = pl.z * tr[r];
        const float H0 = std::fmaf(-q0, rw, Rr[3 * r + 0]);
        const float H1 = std::fmaf(-q1, rw, Rr[3 * r + 1]);
        const float H2 = std::fmaf(-q2, rw, Rr[3 * r + 2]);
        tmp[3 * r + 0] = H0 * rK0;
        tmp[3 * r + 1] = H1 * rK4;
        const float a = H0 * -rc.K[2];
        const float b = (H1 * rc.K[5]) * rK4;
        tmp[3 * r + 2] = H2 + std::fmaf(a, rK0, -b);
    }
    for (int c = 0; c < 3; ++c) {
        const float m0 = sc.K[2] * tmp[6 + c], m1 = sc.K[5] * tmp[6 + c];
        H[c] = std::fmaf(sc.K[0], tmp[c], m0);
        H[3 + c] = std::fmaf(sc.K[4], tmp[3 + c], m1);
        H[6 + c] = sc.K[8] * tmp[6 + c];
    }
}

/* ComputeCorrespondingPoint of a single point, APD.cu:396-403, as built outside the patch loops:
 *   X = H2 + FFMA(H0, x, H1 * y) (same for Y, Z);  pt = (X * RCP(Z), Y * RCP(Z)) */
inline void project_xyz(const float *H, int px, int py, float &X, float &Y, float &Z) {
    const float x = (float)px, y = (float)py;
    const float m0 = H[1] * y, m1 = H[4] * y, m2 = H[7] * y;
    X = H[2] + std::fmaf(H[0], x, m0);
    Y = H[5] + std::fmaf(H[3], x, m1);
    Z = H[8] + std::fmaf(H[6], x, m2);
}
inline F2 corresponding_point(const float *H, int px, int py) {
    float X, Y, Z;
    project_xyz(H, px, py, X, Y, Z);
    const float rz = mufu_rcp(Z);
    return F2{X * rz, Y * rz};
}
/* texture coordinate of a single point: "pt + 0.5f" is contracted with the division, FFMA(X, RCP(Z), 0.5) (APD.cu:687-689) */
inline F2 sample_coord_point(const float *H, int px, int py) {
    float X, Y, Z;
    project_xyz(H, px, py, X, Y, Z);
    const float rz = mufu_rcp(Z);
    return F2{std::fmaf(X, rz, 0.5f), std::fmaf(Y, rz, 0.5f)};
}
/* inside the patch loops (i outer over x, j inner over y; APD.cu:629-634, 523-533) the build hoists the x products out of
 * the inner loop, which changes the contraction:  X = H2 + FFMA(H1, y, H0 * x);  u = FFMA(X, RCP(Z), 0.5) */
inline F2 sample_coord_loop(const float *H, int px, int py) {
    const float x = (float)px, y = (float)py;
    const float x0 = H[0] * x, x3 = H[3] * x, x6 = H[6] * x;
    const float X = H[2] + std::fmaf(H[1], y, x0);
    const float Y = H[5] + std::fmaf(H[4], y, x3);
    const float Z = H[8] + std::fmaf(H[7], y, x6);
    const float rz = mufu_rcp(Z);
    return F2{std::fmaf(X, rz, 0.5f), std::fmaf(Y, rz, 0.5f)};
}

inline F4 load4(const float *p, size_t i) { return F4{p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]}; }
inline void store4(float *p, size_t i, F4 v) { p[4 * i] = v.x; p[4 * i + 1] = v.y; p[4 * i + 2] = v.z; p[4 * i + 3] = v.w; }
inline S2 load_s2(const int16_t *p, size_t i) { return S2{p[2 * i], p[2 * i + 1]}; }
inline void store_s2(int16_t *p, size_t i, S2 v) { p[2 * i] = v.x; p[2 * i + 1] = v.y; }

inline void count(orc_problem *pb, int which) {
#pragma omp atomic
    pb->counters[which]++;
}

/* NCC of one patch around (cx, cy) with the given radius/increment (shared by Old and New).  APD.cu:622-662 */
/* "1.0f / bilateral_weight_sum" is one MUFU.RCP under --use_fast_math.  For the integer tap counts 1..36 it returns the
 * correctly rounded reciprocal except for 33 (0x3cf83e0f, one ulp below 1/33): measured on B200 with tools/probe_mufu_rcp.cu
 * (profiles/r01_mufu_rcp_counts.txt).  Counts other than 36 and 9 only occur with a segment-label map. */
inline float rcp_count(float n) {
    if (n == 33.0f) {
        const uint32_t bits = 0x3cf83e0fu;
        float r;
        std::memcpy(&r, &bits, 4);
        return r;
    }
    return 1.0f / n;
}

inline uint8_t sa_label(const orc_problem *pb, int x, int y) {
    /* the reference indexes sa_mask without a bounds test (APD.cu:494, 527); out-of-image taps are clamped here */
    const int W = pb->width, Hh = pb->height;
    x = x < 0 ? 0 : (x >= W ? W - 1 : x);
    y = y < 0 ? 0 : (y >= Hh ? Hh - 1 : y);
    return pb->sa_mask[x + (size_t)y * W];
}

/* filter_label >= 0: only taps of that segment label take part (NCC-New with use_sa_mask, APD.cu:526-530) */
inline float patch_ncc(const orc_problem *pb, const float *H, int src_idx, int cx, int cy, int radius, int increment,
                       int filter_label = -1) {
    const int W = pb->width, Hh = pb->height;
    const float *ref = pb->images[0];
    const float *src = pb->images[src_idx];
    float sum_ref = 0.0f, sum_ref_ref = 0.0f, sum_src = 0.0f, sum_src_src = 0.0f, sum_ref_src = 0.0f, wsum = 0.0f;
    for (int i = -radius; i <= radius; i += increment) {
        for (int j = -radius; j <= radius; j += increment) {
            const int rx = cx + i, ry = cy + j;
            if (filter_label >= 0 && sa_label(pb, rx, ry) != filter_label) continue;
            const float ref_pix = tex_point(ref, W, Hh, rx, ry);
            const F2 sp = sample_coord_loop(H, rx, ry);
            const float src_pix = tex_linear(src, W, Hh, sp.x, sp.y, pb->tex_mode);
            const float weight = 1.0f; /* quirk 1: "bilateral" weight == 1 (APD.cu:534,635) */
            /* nvcc (-fmad=true, the reference's build) folds weight == 1 and contracts every "sum += a * b" into one
             * FMA.  This matters: on weak texture var = E[x^2] - E[x]^2 is a difference of two ~16384 numbers whose
             * fp32 rounding noise (~2e-3) dwarfs the 1e-5 variance floor, so the contraction pattern decides costs. */
            sum_ref += weight * ref_pix;
            sum_ref_ref = std::fmaf(ref_pix, ref_pix, sum_ref_ref);
            sum_src += weight * src_pix;
            sum_src_src = std::fmaf(src_pix, src_pix, sum_src_src);
            sum_ref_src = std::fmaf(ref_pix, src_pix, sum_ref_src);
            wsum += weight;
        }
    }
    if (wsum == 0.0f) return -1.0f; /* only reachable in NCC-New with a label map: the caller skips the patch (APD.cu:543-545) */
    /* Epilogue exactly as nvcc compiles APD.cu:644-661 for sm_100 with the reference flags (SASS of the reference build:
     * MUFU.RCP; 3x FMUL; FMUL mean^2; FFMA(inv, sum_xx, -mean^2); FFMA(-mean_r, mean_s, E_rs); FFMA(-covar, 1/sqrt, 1)).
     * MUFU.RCP(36) and MUFU.RCP(9) return the correctly rounded reciprocal (tools/probe_mufu_rcp.cu). */
    const float inv = rcp_count(wsum);
    const float mean_ref = inv * sum_ref, mean_src = inv * sum_src, e_rs = inv * sum_ref_src;
    const float var_ref = std::fmaf(inv, sum_ref_ref, -(mean_ref * mean_ref));
    const float var_src = std::fmaf(inv, sum_src_src, -(mean_src * mean_src));
    const float kMinVar = 1e-5f;
    if (var_ref < kMinVar || var_src < kMinVar) return 2.0f;
    const float covar = std::fmaf(-mean_ref, mean_src, e_rs);
    const float denom = mufu_sqrt(var_ref * var_src);
    /* max(0, min(2, v)) with CUDA fminf/fmaxf NaN semantics (quirk 10): NaN -> 2 */
    const float v = std::fmaf(-covar, mufu_rcp(denom), 1.0f);
    return std::fmax(0.0f, std::fmin(2.0f, v));
}

/* NCC-Old branch B, APD.cu:664-719: the 36 odd offsets walked quadrant by quadrant; a tap outside the image is skipped, the
 * first tap of another segment label ends its quadrant.  Same sums and epilogue as branch A. */
inline float patch_ncc_quadrants(const orc_problem *pb, const float *H, int src_idx, int px, int py) {
    static const int sign[] = {1, 1, -1, -1, 1, -1, -1, 1};
    static const int offset[] = {1, 1, 3, 1, 1, 3, 1, 5, 3, 3, 5, 1, 5, 3, 3, 5, 5, 5};
    const int W = pb->width, Hh = pb->height;
    const float *ref = pb->images[0];
    const float *src = pb->images[src_idx];
    const uint8_t center_id = pb->sa_mask[px + (size_t)py * W];
    float sum_ref = 0.0f, sum_ref_ref = 0.0f, sum_src = 0.0f, sum_src_src = 0.0f, sum_ref_src = 0.0f, wsum = 0.0f;
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 9; ++j) {
            const int rx = px + offset[j * 2] * sign[i * 2], ry = py + offset[j * 2 + 1] * sign[i * 2 + 1];
            if (rx < 0 || rx >= W || ry < 0 || ry >= Hh) continue;
            if (pb->sa_mask[rx + (size_t)ry * W] != center_id) break;
            const float ref_pix = tex_point(ref, W, Hh, rx, ry);
            const F2 sp = sample_coord_point(H, rx, ry);
            const float src_pix = tex_linear(src, W, Hh, sp.x, sp.y, pb->tex_mode);
            sum_ref += ref_pix;
            sum_ref_ref = std::fmaf(ref_pix, ref_pix, sum_ref_ref);
            sum_src += src_pix;
            sum_src_src = std::fmaf(src_pix, src_pix, sum_src_src);
            sum_ref_src = std::fmaf(ref_pix, src_pix, sum_ref_src);
            wsum += 1.0f;
        }
    }
    if (wsum == 0.0f) return 2.0f; /* 1/0 = inf, 0 * inf = NaN, every comparison false, max(0, min(2, NaN)) = 2 (quirk 10) */
    const float inv = rcp_count(wsum);
    const float mean_ref = inv * sum_ref, mean_src = inv * sum_src, e_rs = inv * sum_ref_src;
    const float var_ref = std::fmaf(inv, sum_ref_ref, -(mean_ref * mean_ref));
    const float var_src = std::fmaf(inv, sum_src_src, -(mean_src * mean_src));
    const float kMinVar = 1e-5f;
    if (var_ref < kMinVar || var_src < kMinVar) return 2.0f;
    const float covar = std::fmaf(-mean_ref, mean_src, e_rs);
    const float denom = mufu_sqrt(var_ref * var_src);
    const float v = std::fmaf(-covar, mufu_rcp(denom), 1.0f);
    return std::fmax(0.0f, std::fmin(2.0f, v));
}

/* APD.cu:596-721 */
float ncc_old(orc_problem *pb, I2 p, int src_idx, F4 plane) {
    count(pb, 0);
    const orc_camera &rc = pb->cameras[0];
    const orc_camera &sc = pb->cameras[src_idx];
    float H[9];
    homography(rc, sc, plane, H);
    const F2 pt = corresponding_point(H, p.x, p.y);
    if (pt.x >= sc.width || pt.x < 0.0f || pt.y >= sc.height || pt.y < 0.0f) return 2.0f;
    if (pb->sa_mask) {
        /* "const int center = pt.y * src_camera.width + pt.x; if (sa_mask[center] == 0)" (APD.cu:619-621): the REFERENCE
         * view's label map indexed by the projected point, float arithmetic (one FFMA under -fmad) truncated to int */
        const long long c = (long long)std::fmaf(pt.y, (float)sc.width, pt.x);
        /* for points in the last source row the index of FLOAT coordinates runs past the map (up to W - 1 elements): the
         * reference reads past its buffer there (undefined); defined here as label 0 -> branch A, which is what the reference
         * binary returned for every such tuple of tests/golden/ref_costs_sa.npz */
        if (c < (long long)pb->width * pb->height && pb->sa_mask[c] != 0) return patch_ncc_quadrants(pb, H, src_idx, p.x, p.y);
    }
    return patch_ncc(pb, H, src_idx, p.x, p.y, pb->params.strong_radius, pb->params.strong_increment);
}

inline void softmax(float *c, int n) { /* APD.cu:431-446; exp(x) = EX2(x * log2e), "/ sum" = "* RCP(sum)" as built */
    float mx = -1e10f;
    for (int i = 0; i < n; i++) if (c[i] > mx) mx = c[i];
    float sum = 0.0f;
    for (int i = 0; i < n; i++) { c[i] = mufu_ex2((c[i] - mx) * 1.4426950216293334961f); sum += c[i]; }
    const float rs = mufu_rcp(sum);
    for (int i = 0; i < n; i++) c[i] *= rs;
}

/* APD.cu:448-593 */
float ncc_new(orc_problem *pb, I2 p, int src_idx, F4 plane) {
    count(pb, 1);
    const orc_camera &rc = pb->cameras[0];
    const orc_camera &sc = pb->cameras[src_idx];
    const int W = pb->width, Hh = pb->height;
    const size_t center = p.x + (size_t)p.y * W;
    const float cost_max = 2.0f;
    float H[9];
    homography(rc, sc, plane, H);
    const F2 pt = corresponding_point(H, p.x, p.y);
    if (pt.x >= sc.width || pt.x < 0.0f || pt.y >= sc.height || pt.y < 0.0f) return cost_max;
    float cost = 0.0f;
    float strong_costs[9];
    int strong_num = 0;
    if (pb->weak_info[center] != ORC_WEAK) return cost; /* reference prints "error" and returns 0 (APD.cu:590) */
    float center_cost = 0.0f, strong_weight = 0.0f;
    const int center_id = pb->sa_mask ? pb->sa_mask[center] : 0;
    const bool use_sa_mask = center_id != 0; /* APD.cu:463-465 */
    for (int k = 0; k < ORC_ANCHOR_NUM; ++k) {
        const S2 a = load_s2(pb->anchors, center * ORC_ANCHOR_NUM + k);
        if (a.x == -1 || a.y == -1) continue;
        if (use_sa_mask && pb->sa_mask[a.x + (size_t)a.y * W] != center_id) continue; /* APD.cu:493-497 */
        const F2 asp = corresponding_point(H, a.x, a.y);
        if (asp.x < 0 || asp.y < 0 || asp.x >= W || asp.y >= Hh) {
            if (k != 0) {
                const uint32_t vi = pb->selected_views[a.x + (size_t)a.y * W];
                if (is_set(vi, src_idx - 1)) { strong_costs[strong_num++] = cost_max; strong_weight += 1; }
                continue;
            } else {
                return cost_max;
            }
        }
        const int radius = (k == 0 ? pb->params.strong_radius : pb->params.weak_radius);
        const int increment = (k == 0 ? pb->params.strong_increment : pb->params.weak_increment);
        const float tc = patch_ncc(pb, H, src_idx, a.x, a.y, radius, increment, use_sa_mask ? center_id : -1);
        if (tc < 0.0f) continue; /* no tap of the centre's segment: bilateral_weight_sum == 0 (APD.cu:543-545) */
        if (k == 0) center_cost = tc;
        else { strong_costs[strong_num++] = tc; strong_weight += 1; }
    }
    if (strong_weight <= 1e-6f) {
        cost = center_cost;
    } else {
        float wts[9];
        for (int i = 0; i < strong_num; ++i) wts[i] = strong_costs[i];
        softmax(wts, strong_num);
        float strong_cost = 0.0f;
        for (int i = 0; i < strong_num; ++i) strong_cost = std::fmaf(wts[i], strong_costs[i], strong_cost);
        strong_cost = ORC_MIN(strong_cost, cost_max);
        cost = (float)(0.25 * (double)center_cost + 0.75 * (double)strong_cost); /* double literals, APD.cu:586 */
    }
    return cost;
}

inline F3 point_on_world(float x, float y, float depth, const orc_camera &cam) { /* APD.cu:831-851 */
    F3 X, T;
    X.x = ((x - cam.K[2]) * depth) * mufu_rcp(cam.K[0]); /* as built */
    X.y = ((y - cam.K[5]) * depth) * mufu_rcp(cam.K[4]);
    X.z = depth;
    T.x = dot3_ref(cam.R[0], X.x, cam.R[3], X.y, cam.R[6], X.z);
    T.y = dot3_ref(cam.R[1], X.x, cam.R[4], X.y, cam.R[7], X.z);
    T.z = dot3_ref(cam.R[2], X.x, cam.R[5], X.y, cam.R[8], X.z);
    X.x = T.x + cam.c[0];
    X.y = T.y + cam.c[1];
    X.z = T.z + cam.c[2];
    return X;
}

/* APD.cu:853-863 as built; num = the two numerators, rdepth = RCP(depth) (callers contract "p - num / depth" into one FFMA) */
inline void project_on_camera_parts(F3 P, const orc_camera &cam, F2 &num, float &depth, float &rdepth) {
    F3 t;
    t.x = dot3_ref(cam.R[0], P.x, cam.R[1], P.y, cam.R[2], P.z) + cam.t[0];
    t.y = dot3_ref(cam.R[3], P.x, cam.R[4], P.y, cam.R[5], P.z) + cam.t[1];
    t.z = dot3_ref(cam.R[6], P.x, cam.R[7], P.y, cam.R[8], P.z) + cam.t[2];
    depth = dot3_ref(cam.K[6], t.x, cam.K[7], t.y, cam.K[8], t.z);
    rdepth = mufu_rcp(depth);
    num.x = dot3_ref(cam.K[0], t.x, cam.K[1], t.y, cam.K[2], t.z);
    num.y = dot3_ref(cam.K[3], t.x, cam.K[4], t.y, cam.K[5], t.z);
}
inline void project_on_camera(F3 P, const orc_camera &cam, F2 &pt, float &depth) {
    F2 num; float rd;
    project_on_camera_parts(P, cam, num, depth, rd);
    pt.x = num.x * rd;
    pt.y = num.y * rd;
}

/* nearest-texel depth fetch "(int)x + 0.5f" through a clamped linear texture (APD.cu:885, 2319) */
inline float depth_fetch(const orc_problem *pb, int idx, F2 pt) {
    const int ix = clampi(cuda_f2i(pt.x), 0, pb->width - 1);
    const int iy = clampi(cuda_f2i(pt.y), 0, pb->height - 1);
    return pb->depths[idx][iy * (size_t)pb->width + ix];
}

/* APD.cu:865-902 */
float geom_cost(orc_problem *pb, I2 p, int src_idx, F4 plane) {
    count(pb, 2);
    const orc_camera &rc = pb->cameras[0];
    const orc_camera &sc = pb->cameras[src_idx];
    const float max_cost = 3.0f;
    const float depth = depth_from_plane(rc, plane, p);
    const F3 fwd = point_on_world((float)p.x, (float)p.y, depth, rc);
    F2 sp; float sd;
    project_on_camera(fwd, sc, sp, sd);
    const float src_depth = depth_fetch(pb, src_idx, sp);
    if (src_depth == 0.0f) return max_cost;
    const F3 s3 = point_on_world(sp.x, sp.y, src_depth, sc);
    F2 bn; float rd, rrd;
    project_on_camera_parts(s3, rc, bn, rd, rrd);
    /* as built: diff = FFMA(-num, RCP(depth), p); cost = min(3, SQRT(FFMA(dc, dc, dr * dr))) */
    const float dc = std::fmaf(-bn.x, rrd, (float)p.x), dr = std::fmaf(-bn.y, rrd, (float)p.y);
    const float drr = dr * dr;
    const float cc = mufu_sqrt(std::fmaf(dc, dc, drr));
    return std::fmin(max_cost, cc);
}

/* APD.cu:723-774 */
float initial_cost_and_views(orc_problem *pb, I2 p) {
    const orc_params &prm = pb->params;
    const size_t center = p.x + (size_t)p.y * pb->width;
    const F4 plane = load4(pb->planes, center);
    const float cost_max = 2.0f;
    float cv[32] = {2.0f};
    float cvc[32] = {2.0f};
    int cost_count = 0, num_valid = 0;
    for (int i = 1; i < prm.num_images; ++i) {
        float c;
        if (prm.use_APD && pb->weak_info[center] == ORC_WEAK) c = ncc_new(pb, p, i, plane);
        else c = ncc_old(pb, p, i, plane);
        cv[i - 1] = c; cvc[i - 1] = c;
        cost_count++;
        if (c < cost_max) num_valid++;
    }
    sort_small(cv, cost_count);
    uint32_t sel = 0;
    const int top_k = std::min(num_valid, prm.top_k);
    float ret = cost_max;
    if (top_k > 0) {
        float cost = 0.0f;
        for (int i = 0; i < top_k; ++i) cost += cv[i];
        const float thr = cv[top_k - 1];
        for (int i = 0; i < prm.num_images - 1; ++i) if (cvc[i] <= thr) sel |= (1u << i);
        ret = cost / top_k;
    }
    pb->selected_views[center] = sel;
    return ret;
}

/* rows a half-grid launch covers (quirk 7, APD.cu:2676-2678) */
inline int half_rows_limit(int H) { return 32 * (((H / 2) + 15) / 16); }
/* Black <=> (x+y) even (APD.cu:1619-1626, 1656-1663) */
inline bool is_color(int x, int y, int color) { return ((x + y) & 1) == color; }

/* APD.cu:1127-1314: the 8 adaptive-checkerboard candidates */
void checkerboard_candidates(const float *costs, int width, int height, I2 p, int positions[8], bool flag[8]) {
    const int center = p.y * width + p.x;
    for (int i = 0; i < 8; ++i) flag[i] = false;
    float costMin; int costMinPoint;
    int left_near = center - 1, left_far = center - 3, right_near = center + 1, right_far = center + 3;
    int up_near = center - width, up_far = center - 3 * width, down_near = center + width, down_far = center + 3 * width;
    if (p.y > 2) {
        flag[1] = true; costMin = costs[up_far]; costMinPoint = up_far;
        for (int i = 1; i < 11; ++i) if (p.y > 2 + 2 * i) {
            int pt = up_far - 2 * i * width;
            if (costs[pt] < costMin) { costMin = costs[pt]; costMinPoint = pt; }
        }
        up_far = costMinPoint;
    }
    if (p.y < height - 3) {
        flag[3] = true; costMin = costs[down_far]; costMinPoint = down_far;
        for (int i = 1; i < 11; ++i) if (p.y < height - 3 - 2 * i) {
            int pt = down_far + 2 * i * width;
            if (costs[pt] < costMin) { costMin = costs[pt]; costMinPoint = pt; }
        }
        down_far = costMinPoint;
    }
    if (p.x > 2) {
        flag[5] = true; costMin = costs[left_far]; costMinPoint = left_far;
        for (int i = 1; i < 11; ++i) if (p.x > 2 + 2 * i) {
            int pt = left_far - 2 * i;
            if (costs[pt] < costMin) { costMin = costs[pt]; costMinPoint = pt; }
        }
        left_far = costMinPoint;
    }
    if (p.x < width - 3) {
        flag[7] = true; costMin = costs[right_far]; costMinPoint = right_far;
        for (int i = 1; i < 11; ++i) if (p.x < width - 3 - 2 * i) {
            int pt = right_far + 2 * i;
            if (costs[pt] < costMin) { costMin = costs[pt]; costMinPoint = pt; }
        }
        right_far = costMinPoint;
    }
    if (p.y > 0) {
        flag[0] = true; costMin = costs[up_near]; costMinPoint = up_near;
        for (int i = 0; i < 3; ++i) {
            if (p.y > 1 + i && p.x > i) {
                int pt = up_near - (1 + i) * width - (i + 1);
                if (costs[pt] < costMin) { costMin = costs[pt]; costMinPoint = pt; }
            }
            if (p.y > 1 + i && p.x < width - 1 - i) {
                int pt = up_near - (1 + i) * width + (i + 1);
                if (costs[pt] < costMin) { costMin = costs[pt]; costMinPoint = pt; }
            }
        }
        up_near = costMinPoint;
    }
    if (p.y < height - 1) {
        flag[2] = true; costMin = costs[down_near]; costMinPoint = down_near;
        for (int i = 0; i < 3; ++i) {
            if (p.y < height - 2 - i && p.x > i) {
                int pt = down_near + (1 + i) * width - (i + 1);
                if (costs[pt] < costMin) { costMin = costs[pt]; costMinPoint = pt; }
            }
            if (p.y < height - 2 - i && p.x < width - 1 - i) {
                int pt = down_near + (1 + i) * width + (i + 1);
                if (costs[pt] < costMin) { costMin = costs[pt]; costMinPoint = pt; }
            }
        }
        down_near = costMinPoint;
    }
    if (p.x > 0) {
        flag[4] = true; costMin = costs[left_near]; costMinPoint = left_near;
        for (int i = 0; i < 3; ++i) {
            if (p.x > 1 + i && p.y > i) {
                int pt = left_near - (1 + i) - (i + 1) * width;
                if (costs[pt] < costMin) { costMin = costs[pt]; costMinPoint = pt; }
            }
            if (p.x > 1 + i && p.y < height - 1 - i) {
                int pt = left_near - (1 + i) + (i + 1) * width;
                if (costs[pt] < costMin) { costMin = costs[pt]; costMinPoint = pt; }
            }
        }
        left_near = costMinPoint;
    }
    if (p.x < width - 1) {
        flag[6] = true; costMin = costs[right_near]; costMinPoint = right_near;
        for (int i = 0; i < 3; ++i) {
            if (p.x < width - 2 - i && p.y > i) {
                int pt = right_near + (1 + i) - (i + 1) * width;
                if (costs[pt] < costMin) { costMin = costs[pt]; costMinPoint = pt; }
            }
            if (p.x < width - 2 - i && p.y < height - 1 - i) {
                int pt = right_near + (1 + i) + (i + 1) * width;
                if (costs[pt] < costMin) { costMin = costs[pt]; costMinPoint = pt; }
            }
        }
        right_near = costMinPoint;
    }
    const int pos[8] = {up_near, up_far, down_near, down_far, left_near, left_far, right_near, right_far};
    for (int i = 0; i < 8; ++i) positions[i] = pos[i];
}

/* joint view selection shared by strong/weak propagation (APD.cu:1339-1386, 1505-1552) */
void select_views(const orc_problem *pb, float cost_array[8][32], const float *priors, int iter, Rng &rng,
                  uint8_t *view_weights, uint32_t *temp_selected, float *weight_norm) {
    const int nv = pb->params.num_images - 1;
    float sampling_probs[32] = {0.0f};
    const float cost_threshold = (float)(0.8 * (double)std::exp((float)(iter * iter) / (-90.0f)));
    for (int i = 0; i < nv; i++) {
        float cnt = 0; int cnt_false = 0; float tmpw = 0;
        for (int j = 0; j < 8; j++) {
            if (cost_array[j][i] < cost_threshold) {
                tmpw += std::exp(cost_array[j][i] * cost_array[j][i] / (-0.18f));
                cnt++;
            }
            if (cost_array[j][i] > 1.2f) cnt_false++;
        }
        if (cnt > 2 && cnt_false < 3) sampling_probs[i] = tmpw / cnt;
        else if (cnt_false < 3) sampling_probs[i] = std::exp(cost_threshold * cost_threshold / (-0.32f));
        sampling_probs[i] = sampling_probs[i] * priors[i];
    }
    pdf_to_cdf(sampling_probs, nv);
    for (int s = 0; s < 15; ++s) {
        const float r = rng.uniform() - FLT_EPSILON;
        for (int v = 0; v < nv; ++v) {
            if (sampling_probs[v] > r) { view_weights[v] += 1; break; }
        }
    }
    uint32_t sel = 0; float wn = 0;
    for (int i = 0; i < nv; ++i) if (view_weights[i] > 0) { sel |= (1u << i); wn += view_weights[i]; }
    *temp_selected = sel; *weight_norm = wn;
}

/* APD.cu:950-1006 */
void refine_strong(orc_problem *pb, F4 *plane, float *depth, float *cost, Rng &rng, const uint8_t *vw,
                   float weight_norm, I2 p) {
    const orc_camera &cam = pb->cameras[0];
    const orc_params &prm = pb->params;
    const float depth_perturbation = 0.02f, normal_perturbation = 0.02f;
    const float dmin = prm.depth_min, dmax = prm.depth_max;
    const float depth_rand = rng.uniform() * (dmax - dmin) + dmin;
    const F4 n_rand = random_normal(cam, p, rng, *depth);
    float depth_pert = *depth;
    const float lo = (1 - depth_perturbation) * depth_pert, hi = (1 + depth_perturbation) * depth_pert;
    depth_pert = rng.uniform() * (hi - lo) + lo; /* do-while can never repeat (quirk 6) */
    const F4 n_pert = perturbed_normal(cam, p, *plane, rng, (float)((double)normal_perturbation * kPiD));
    const float depths[5] = {depth_rand, *depth, depth_rand, *depth, depth_pert};
    const F4 normals[5] = {*plane, n_rand, n_rand, n_pert, *plane};
    for (int i = 0; i < 5; ++i) {
        F4 tp = normals[i];
        tp.w = distance_to_origin(cam, p, depths[i], tp);
        float cvv[32] = {2.0f};
        for (int v = 1; v < prm.num_images; ++v) cvv[v - 1] = ncc_old(pb, p, v, tp);
        float tc = 0.0f;
        for (int j = 0; j < prm.num_images - 1; ++j) {
            if (prm.geom_consistency && prm.use_impetus)
                tc += vw[j] * (cvv[j] + prm.geom_factor * geom_cost(pb, p, j + 1, tp));
            else
                tc += vw[j] * cvv[j];
        }
        tc /= weight_norm;
        const float db = depth_from_plane(cam, tp, p);
        if (db >= dmin && db <= dmax && tc < *cost) { *depth = db; *plane = tp; *cost = tc; }
    }
}

/* APD.cu:1098-1440 */
void propagate_strong_pixel(orc_problem *pb, I2 p, int iter) {
    const int width = pb->width, height = pb->height;
    const orc_params &prm = pb->params;
    const orc_camera &cam = pb->cameras[0];
    const int nv = prm.num_images - 1;
    const int center = p.y * width + p.x;
    float cost_array[8][32];
    std::memset(cost_array, 0, sizeof(cost_array));
    cost_array[0][0] = 2.0f; /* quirk 2: "= {2.0f}" sets only [0][0] */
    bool flag[8]; int positions[8];
    checkerboard_candidates(pb->costs, width, height, p, positions, flag);
    /* evaluation order of the reference: 1,3,5,7,0,2,4,6 -- order is irrelevant for the values */
    for (int k = 0; k < 8; ++k) if (flag[k]) {
        const F4 pl = load4(pb->planes, positions[k]);
        for (int v = 1; v < prm.num_images; ++v) cost_array[k][v - 1] = ncc_old(pb, p, v, pl);
    }
    uint8_t *vw = &pb->view_weight[(size_t)center * ORC_MAX_IMAGES];
    for (int i = 0; i < ORC_MAX_IMAGES; ++i) vw[i] = 0;
    float priors[32] = {0.0f};
    const int nbr[4] = {center - width, center + width, center - 1, center + 1};
    for (int i = 0; i < 4; ++i) if (flag[2 * i]) {
        for (int j = 0; j < nv; ++j) {
            if (is_set(pb->selected_views[nbr[i]], j) == 1) priors[j] += 0.9f;
            else priors[j] += 0.1f;
        }
    }
    Rng rng(pb->seed, pb->stream, (uint32_t)center, SITE_STRONG + iter);
    uint32_t temp_sel; float weight_norm;
    select_views(pb, cost_array, priors, iter, rng, vw, &temp_sel, &weight_norm);

    float final_costs[8] = {0.0f};
    for (int i = 0; i < 8; ++i) {
        for (int j = 0; j < nv; ++j) if (vw[j] > 0) final_costs[i] += vw[j] * cost_array[i][j];
        final_costs[i] /= weight_norm;
    }
    const int min_idx = find_min_cost_index(final_costs, 8);

    const F4 plane_c = load4(pb->planes, center);
    float cvn[32] = {2.0f};
    for (int v = 1; v < prm.num_images; ++v) cvn[v - 1] = ncc_old(pb, p, v, plane_c);
    float cost_now = 0.0f;
    for (int i = 0; i < nv; ++i) {
        if (prm.geom_consistency && prm.use_impetus)
            cost_now += vw[i] * (cvn[i] + prm.geom_factor * geom_cost(pb, p, i + 1, plane_c));
        else
            cost_now += vw[i] * cvn[i];
    }
    cost_now /= weight_norm;
    pb->costs[center] = cost_now;
    float depth_now = depth_from_plane(cam, plane_c, p);
    F4 plane_now = plane_c;
    if (flag[min_idx]) {
        const F4 cand = load4(pb->planes, positions[min_idx]);
        const float db = depth_from_plane(cam, cand, p);
        if (db >= prm.depth_min && db <= prm.depth_max && final_costs[min_idx] < cost_now) {
            depth_now = db; plane_now = cand; cost_now = final_costs[min_idx];
            pb->selected_views[center] = temp_sel;
        }
    }
    refine_strong(pb, &plane_now, &depth_now, &cost_now, rng, vw, weight_norm, p);
    if (prm.state == ORC_REFINE_INIT) {
        if ((double)cost_now < (double)pb->costs[center] - 0.1) { /* double literal, APD.cu:1431 */
            pb->costs[center] = cost_now; store4(pb->planes, center, plane_now);
        }
    } else {
        pb->costs[center] = cost_now; store4(pb->planes, center, plane_now);
    }
}

/* APD.cu:1008-1096 */
void refine_weak(orc_problem *pb, F4 *plane, float *depth, float *cost, Rng &rng, const uint8_t *vw,
                 float weight_norm, I2 p) {
    const orc_camera &cam = pb->cameras[0];
    const orc_params &prm = pb->params;
    const float depth_perturbation = 0.02f, normal_perturbation = 0.02f;
    const float dmin = prm.depth_min, dmax = prm.depth_max;
    const size_t center = p.x + (size_t)p.y * pb->width;
    {
        const F4 fit = load4(pb->fit_planes, center);
        if (fit.x == 0 && fit.y == 0 && fit.z == 0) return;
        float cvv[32] = {2.0f};
        for (int v = 1; v < prm.num_images; ++v) cvv[v - 1] = ncc_new(pb, p, v, fit);
        float tc = 0.0f;
        for (int j = 0; j < prm.num_images - 1; ++j) if (vw[j] > 0) {
            if (prm.geom_consistency) tc += vw[j] * (cvv[j] + prm.geom_factor * geom_cost(pb, p, j + 1, fit));
            else tc += vw[j] * cvv[j];
        }
        tc /= weight_norm;
        const float db = depth_from_plane(cam, fit, p);
        if (db >= dmin && db <= dmax && tc < *cost) { *depth = db; *plane = fit; *cost = tc; }
    }
    const float depth_rand = rng.uniform() * (dmax - dmin) + dmin;
    const F4 n_rand = random_normal(cam, p, rng, *depth);
    float depth_pert = *depth;
    const float lo = (1 - depth_perturbation) * depth_pert, hi = (1 + depth_perturbation) * depth_pert;
    depth_pert = rng.uniform() * (hi - lo) + lo;
    const F4 n_pert = perturbed_normal(cam, p, *plane, rng, (float)((double)normal_perturbation * kPiD));
    const float depths[5] = {depth_rand, *depth, depth_rand, *depth, depth_pert};
    const F4 normals[5] = {*plane, n_rand, n_rand, n_pert, *plane};
    for (int i = 0; i < 5; ++i) {
        F4 tp = normals[i];
        tp.w = distance_to_origin(cam, p, depths[i], tp);
        float cvv[32] = {2.0f};
        for (int v = 1; v < prm.num_images; ++v) cvv[v - 1] = ncc_new(pb, p, v, tp);
        float tc = 0.0f;
        for (int j = 0; j < prm.num_images - 1; ++j) if (vw[j] > 0) {
            if (prm.geom_consistency) tc += vw[j] * (cvv[j] + prm.geom_factor * geom_cost(pb, p, j + 1, tp));
            else tc += vw[j] * cvv[j];
        }
        tc /= weight_norm;
        const float db = depth_from_plane(cam, tp, p);
        if (db >= dmin && db <= dmax && tc < *cost) { *depth = db; *plane = tp; *cost = tc; }
    }
}

/* APD.cu:1442-1615 */
void propagate_weak_pixel(orc_problem *pb, I2 p, int iter) {
    const int width = pb->width;
    const orc_params &prm = pb->params;
    const orc_camera &cam = pb->cameras[0];
    const int nv = prm.num_images - 1;
    const size_t center = p.y * (size_t)width + p.x;
    float cost_array[8][32];
    std::memset(cost_array, 0, sizeof(cost_array));
    cost_array[0][0] = 2.0f;
    bool flag[8]; size_t positions[8] = {0}; F4 new_planes[8];
    for (int i = 0; i < 8; ++i) {
        const S2 a = load_s2(pb->anchors, center * ORC_ANCHOR_NUM + i + 1);
        if (a.x == -1 || a.y == -1 || pb->weak_info[a.x + (size_t)a.y * width] != ORC_STRONG) { flag[i] = false; continue; }
        positions[i] = a.x + (size_t)a.y * width;
        flag[i] = true;
        new_planes[i] = load4(pb->planes, positions[i]);
        for (int v = 1; v < prm.num_images; ++v) cost_array[i][v - 1] = ncc_new(pb, p, v, new_planes[i]);
    }
    uint8_t *vw = &pb->view_weight[center * ORC_MAX_IMAGES];
    for (int i = 0; i < ORC_MAX_IMAGES; ++i) vw[i] = 0;
    float priors[32] = {0.0f};
    for (int i = 0; i < 8; ++i) {
        const S2 a = load_s2(pb->anchors, center * ORC_ANCHOR_NUM + i + 1);
        if (a.x == -1 || a.y == -1) continue;
        for (int j = 0; j < nv; ++j) {
            if (is_set(pb->selected_views[a.x + (size_t)a.y * width], j) == 1) priors[j] += 0.9f;
            else priors[j] += 0.1f;
        }
    }
    Rng rng(pb->seed, pb->stream, (uint32_t)center, SITE_WEAK + iter);
    uint32_t temp_sel; float weight_norm;
    select_views(pb, cost_array, priors, iter, rng, vw, &temp_sel, &weight_norm);

    float final_costs[8] = {0.0f};
    for (int i = 0; i < 8; ++i) {
        for (int j = 0; j < nv; ++j) if (vw[j] > 0) {
            if (prm.geom_consistency) {
                if (flag[i]) final_costs[i] += vw[j] * (cost_array[i][j] + prm.geom_factor * geom_cost(pb, p, j + 1, load4(pb->planes, positions[i])));
                else final_costs[i] += vw[j] * (cost_array[i][j] + prm.geom_factor * 3.0f);
            } else {
                final_costs[i] += vw[j] * cost_array[i][j];
            }
        }
        final_costs[i] /= weight_norm;
    }
    const int min_idx = find_min_cost_index(final_costs, 8);
    const F4 plane_c = load4(pb->planes, center);
    float cvn[32] = {2.0f};
    for (int v = 1; v < prm.num_images; ++v) cvn[v - 1] = ncc_new(pb, p, v, plane_c);
    float cost_now = 0.0f;
    for (int i = 0; i < nv; ++i) {
        if (prm.geom_consistency) cost_now += vw[i] * (cvn[i] + prm.geom_factor * geom_cost(pb, p, i + 1, plane_c));
        else cost_now += vw[i] * cvn[i];
    }
    cost_now /= weight_norm;
    pb->costs[center] = cost_now;
    float depth_now = depth_from_plane(cam, plane_c, p);
    F4 plane_now = plane_c;
    if (flag[min_idx]) {
        const float db = depth_from_plane(cam, new_planes[min_idx], p);
        if (db >= prm.depth_min && db <= prm.depth_max && final_costs[min_idx] < cost_now) {
            depth_now = db; plane_now = new_planes[min_idx]; cost_now = final_costs[min_idx];
            pb->selected_views[center] = temp_sel;
        }
    }
    refine_weak(pb, &plane_now, &depth_now, &cost_now, rng, vw, weight_norm, p);
    if (prm.state == ORC_REFINE_INIT) {
        if ((double)cost_now < (double)pb->costs[center] - 0.1) { pb->costs[center] = cost_now; store4(pb->planes, center, plane_now); }
    } else {
        pb->costs[center] = cost_now; store4(pb->planes, center, plane_now);
    }
}

/* APD.cu:1711-1821 */
void median_pixel(orc_problem *pb, I2 p) {
    const int width = pb->width, height = pb->height;
    const int center = p.y * width + p.x;
    float *pl = pb->planes;
    const uint8_t *wk = pb->weak_info;
    float filter[21]; int index = 0;
    filter[index++] = pl[4 * center + 3];
    const int left = center - 1, leftleft = center - 3, up = center - width, upup = center - 3 * width;
    const int down = center + width, downdown = center + 3 * width, right = center + 1, rightright = center + 3;
    if (pb->costs[center] < 0.001f) return;
#define TAP(cond, idx) if ((cond) && wk[(idx)] == ORC_STRONG) filter[index++] = pl[4 * (idx) + 3];
    TAP(p.y > 0, up)
    TAP(p.y > 2, upup)
    TAP(p.y > 4, upup - width * 2)
    TAP(p.y < height - 1, down)
    TAP(p.y < height - 3, downdown)
    TAP(p.y < height - 5, downdown + width * 2)
    TAP(p.x > 0, left)
    TAP(p.x > 2, leftleft)
    TAP(p.x > 4, leftleft - 2)
    TAP(p.x < width - 1, right)
    TAP(p.x < width - 3, rightright)
    TAP(p.x < width - 5, rightright + 2)
    TAP(p.y > 0 && p.x < width - 2, up + 2)
    TAP(p.y < height - 1 && p.x < width - 2, down + 2)
    TAP(p.y > 0 && p.x > 1, up - 2)
    TAP(p.y < height - 1 && p.x > 1, down - 2)
    TAP(p.x > 0 && p.y > 2, left - width * 2)
    TAP(p.x < width - 1 && p.y > 2, right - width * 2)
    TAP(p.x > 0 && p.y < height - 2, left + width * 2)
    TAP(p.x < width - 1 && p.y < height - 2, right + width * 2)
#undef TAP
    sort_small(filter, index);
    const int m = index / 2;
    if (index % 2 == 0) pl[4 * center + 3] = (filter[m - 1] + filter[m]) / 2;
    else pl[4 * center + 3] = filter[m];
}

inline float baseline_len(const orc_camera &a, const orc_camera &b) { /* APD.cu:2142-2147 */
    const float d0 = a.c[0] - b.c[0], d1 = a.c[1] - b.c[1], d2 = a.c[2] - b.c[2];
    return mufu_sqrt(dot3_ref(d0, d0, d1, d1, d2, d2)); /* as built: FFMA(d2, d2, FFMA(d0, d0, d1 * d1)), MUFU.SQRT */
}

bool point_in_triangle(S2 A, S2 B, S2 C, I2 P) { /* APD.cu:122-143 */
    const F2 AB{(float)(B.x - A.x), (float)(B.y - A.y)};
    const F2 BC{(float)(C.x - B.x), (float)(C.y - B.y)};
    const F2 CA{(float)(A.x - C.x), (float)(A.y - C.y)};
    const float ab = std::sqrt(AB.x * AB.x + AB.y * AB.y);
    const float bc = std::sqrt(BC.x * BC.x + BC.y * BC.y);
    const float ca = std::sqrt(CA.x * CA.x + CA.y * CA.y);
    if (ab <= 2 || bc <= 2 || ca <= 2) return false;
    if (!(ab + bc > ca && bc + ca > ab && ab + ca > bc)) return false;
    const F2 PA{(float)(A.x - P.x), (float)(A.y - P.y)};
    const F2 PB{(float)(B.x - P.x), (float)(B.y - P.y)};
    const F2 PC{(float)(C.x - P.x), (float)(C.y - P.y)};
    const float t1 = PA.x * PB.y - PA.y * PB.x;
    const float t2 = PB.x * PC.y - PB.y * PC.x;
    const float t3 = PC.x * PA.y - PC.y * PA.x;
    return t1 * t2 >= 0 && t1 * t3 >= 0;
}

void sort_small_weighted(S2 *pts, float *w, int n) { /* APD.cu:25-38 */
    int j;
    for (int i = 1; i < n; i++) {
        S2 tmp = pts[i]; float tw = w[i];
        for (j = i; j >= 1 && tw < w[j - 1]; j--) { pts[j] = pts[j - 1]; w[j] = w[j - 1]; }
        pts[j] = tmp; w[j] = tw;
    }
}

/* "(curand() % 2 == 0 ? 1 : -1) * curand() % shift_range" (APD.cu:1921): int * unsigned -> unsigned arithmetic */
inline int rand_shift(Rng &rng, int shift_range) {
    const int sgn = (rng.next() % 2 == 0 ? 1 : -1);
    const uint32_t r = rng.next();
    const uint32_t prod = (uint32_t)sgn * r;
    return (int)(prod % (uint32_t)shift_range);
}

int set_threads(const orc_problem *pb) {
    int n = pb->num_threads > 0 ? pb->num_threads : 1;
#ifdef _OPENMP
    omp_set_num_threads(n);
#endif
    return n;
}

}  // namespace

/* ================================================================================================
 * extern "C" API
 * ============================================================================================== */
extern "C" {

void orc_philox(uint32_t seed, uint32_t stream, uint32_t pixel, uint32_t site, uint32_t block, uint32_t out[4]) {
    out[0] = pixel; out[1] = site; out[2] = block; out[3] = 0;
    philox4x32_10(out, seed, stream);
}

void orc_homography(const orc_camera *ref, const orc_camera *src, const float plane[4], float H[9]) {
    homography(*ref, *src, F4{plane[0], plane[1], plane[2], plane[3]}, H);
}

float orc_tex2d(const float *img, int w, int h, float x, float y, int tex_mode) {
    return tex_linear(img, w, h, x, y, tex_mode);
}

float orc_ncc_old(orc_problem *pb, int x, int y, int src_idx, const float plane[4]) {
    return ncc_old(pb, I2{x, y}, src_idx, F4{plane[0], plane[1], plane[2], plane[3]});
}
float orc_ncc_new(orc_problem *pb, int x, int y, int src_idx, const float plane[4]) {
    return ncc_new(pb, I2{x, y}, src_idx, F4{plane[0], plane[1], plane[2], plane[3]});
}
float orc_geom_cost(orc_problem *pb, int x, int y, int src_idx, const float plane[4]) {
    return geom_cost(pb, I2{x, y}, src_idx, F4{plane[0], plane[1], plane[2], plane[3]});
}

void orc_eval_costs(orc_problem *pb, int n, const int32_t *tuples, const float *planes, int mode, float *out) {
    set_threads(pb);
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < n; ++i) {
        const I2 p{tuples[3 * i], tuples[3 * i + 1]};
        const int v = tuples[3 * i + 2];
        const F4 pl = load4(planes, i);
        out[i] = mode == 0 ? ncc_old(pb, p, v, pl) : (mode == 1 ? ncc_new(pb, p, v, pl) : geom_cost(pb, p, v, pl));
    }
}

void orc_checkerboard_candidates(const float *costs, int w, int h, int x, int y, int32_t positions[8], uint8_t flags[8]) {
    bool f[8]; int pos[8];
    checkerboard_candidates(costs, w, h, I2{x, y}, pos, f);
    for (int i = 0; i < 8; ++i) { positions[i] = pos[i]; flags[i] = f[i] ? 1 : 0; }
}

/* K5: APD.cu:919-948 */
void orc_random_init(orc_problem *pb) {
    set_threads(pb);
    const int W = pb->width, H = pb->height;
    const orc_camera &cam = pb->cameras[0];
    /* The deformable cost of a WEAK pixel reads selected_views of its STRONG anchors (APD.cu:502-503) while the reference's
     * kernel is still writing them (a race).  Serialisation chosen here and in the CUDA product: all non-WEAK pixels first,
     * then the WEAK ones. */
    for (int phase = 0; phase < (pb->params.use_APD ? 2 : 1); ++phase)
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        const I2 p{x, y};
        const size_t center = (size_t)y * W + x;
        if (pb->params.use_APD && (int)(pb->weak_info[center] == ORC_WEAK) != phase) continue;
        if (pb->params.state == ORC_FIRST_INIT) {
            Rng rng(pb->seed, pb->stream, (uint32_t)center, SITE_INIT);
            store4(pb->planes, center, random_plane(cam, p, rng, pb->params.depth_min, pb->params.depth_max));
        } else {
            F4 pl = load4(pb->planes, center);
            pl = normal_to_refcam(cam, pl);
            const float depth = pl.w;
            pl.w = distance_to_origin(cam, p, depth, pl);
            store4(pb->planes, center, pl);
        }
        pb->costs[center] = initial_cost_and_views(pb, p);
    }
}

/* K6: APD.cu:1654-1692 */
void orc_propagate_strong(orc_problem *pb, int iter, int color) {
    set_threads(pb);
    const int W = pb->width, H = pb->height;
    const int ylim = std::min(H, half_rows_limit(H));
#pragma omp parallel for schedule(dynamic, 2)
    for (int y = 0; y < ylim; ++y) for (int x = 0; x < W; ++x) {
        if (!is_color(x, y, color)) continue;
        if (pb->weak_info[(size_t)y * W + x] == ORC_WEAK) continue;
        propagate_strong_pixel(pb, I2{x, y}, iter);
    }
}

/* K8: APD.cu:1617-1652 */
void orc_propagate_weak(orc_problem *pb, int iter, int color) {
    set_threads(pb);
    const int W = pb->width, H = pb->height;
    const int ylim = std::min(H, half_rows_limit(H));
#pragma omp parallel for schedule(dynamic, 2)
    for (int y = 0; y < ylim; ++y) for (int x = 0; x < W; ++x) {
        if (!is_color(x, y, color)) continue;
        if (pb->weak_info[(size_t)y * W + x] != ORC_WEAK) continue;
        propagate_weak_pixel(pb, I2{x, y}, iter);
    }
}

/* K9: APD.cu:1694-1709 */
void orc_depth_normal(orc_problem *pb) {
    const int W = pb->width, H = pb->height;
    const orc_camera &cam = pb->cameras[0];
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        const size_t c = (size_t)y * W + x;
        F4 pl = load4(pb->planes, c);
        pl.w = depth_from_plane(cam, pl, I2{x, y});
        store4(pb->planes, c, normal_to_world(cam, pl));
    }
}

/* K10: APD.cu:1823-1855 */
void orc_median_filter(orc_problem *pb, int color) {
    const int W = pb->width, H = pb->height;
    const int ylim = std::min(H, half_rows_limit(H));
    for (int y = 0; y < ylim; ++y) for (int x = 0; x < W; ++x) {
        if (!is_color(x, y, color)) continue;
        if (pb->weak_info[(size_t)y * W + x] != ORC_WEAK) median_pixel(pb, I2{x, y});
    }
}

/* K11: APD.cu:2103-2250 */
void orc_depth_to_weak(orc_problem *pb, float *curve) {
    set_threads(pb);
    const int W = pb->width, H = pb->height;
    const orc_params &prm = pb->params;
    const orc_camera *cams = pb->cameras;
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        const I2 pt{x, y};
        const int min_margin = 6;
        const size_t center = (size_t)y * W + x;
        if (x < min_margin || y < min_margin || x >= W - min_margin || y >= H - min_margin) { pb->weak_info[center] = ORC_UNKNOWN; continue; }
        const uint8_t *vw = &pb->view_weight[ORC_MAX_IMAGES * center];
        F4 opl = normal_to_refcam(cams[0], load4(pb->planes, center));
        const float origin_depth = opl.w;
        if (origin_depth == 0) { pb->weak_info[center] = ORC_UNKNOWN; continue; }
        float base_line = 0; int valid_src = 0; float weight_normal = 0.0f;
        const uint32_t sel = pb->selected_views[center];
        for (int s = 1; s < prm.num_images; ++s) if (is_set(sel, s - 1)) {
            weight_normal += vw[s - 1];
            base_line += baseline_len(cams[0], cams[s]);
            valid_src++;
        }
        if (valid_src == 0) { pb->weak_info[center] = ORC_UNKNOWN; continue; }
        /* as built (APD.cu:2155-2165): base_line * RCP(valid); fb = base_line * K0; disp = fb * RCP(depth); p_depth = fb * RCP(disp + pd) */
        base_line = mufu_rcp((float)valid_src) * base_line;
        const float fb = base_line * cams[0].K[0];
        const float disp = fb * mufu_rcp(origin_depth);
        const float rwn = mufu_rcp(weight_normal);
        const int radius = 30, n = 2 * radius + 1;
        float pc[61];
        for (int pd = -radius; pd <= radius; ++pd) {
            const float p_depth = fb * mufu_rcp(disp + (float)pd);
            if (p_depth < prm.depth_min || p_depth > prm.depth_max) { pc[pd + radius] = 2.0f; continue; }
            F4 tp = opl;
            tp.w = distance_to_origin(cams[0], pt, p_depth, tp);
            float p_cost = 0.0f;
            for (int s = 1; s < prm.num_images; ++s) {
                float tc = 0.0f;
                if (is_set(sel, s - 1)) {
                    tc += ncc_old(pb, pt, s, tp);
                    if (prm.geom_consistency) tc = std::fmaf(prm.geom_factor, geom_cost(pb, pt, s, tp), tc); /* FFMA as built */
                    p_cost = std::fmaf((float)vw[s - 1], tc, p_cost);
                }
            }
            p_cost = p_cost * rwn;
            pc[pd + radius] = ORC_MIN(2.0f, p_cost);
        }
        if (curve) for (int i = 0; i < n; ++i) curve[center * n + i] = pc[i];
        bool is_peak[61];
        for (int i = 0; i < n; ++i) is_peak[i] = false;
        int peak_count = 0, min_peak = 0; float min_cost = 2.0f;
        for (int i = 2; i < n - 2; ++i) {
            if (pc[i - 1] > pc[i] && pc[i + 1] > pc[i]) {
                is_peak[i] = true; peak_count++;
                if (pc[i] < min_cost) { min_peak = i; min_cost = pc[i]; }
            }
        }
        if (std::abs(min_peak - radius) > prm.weak_peak_radius || pc[min_peak] > 0.5f) { pb->weak_info[center] = ORC_WEAK; continue; }
        if (peak_count == 1) { pb->weak_info[center] = (pc[min_peak] <= 0.15f) ? ORC_STRONG : ORC_WEAK; continue; }
        float var = 0.0f;
        for (int i = 2; i < n - 2; ++i) if (is_peak[i] && i != min_peak) { const float d = pc[i] - min_cost; var += d * d; }
        var = std::sqrt(var);
        var /= (peak_count - 1);
        pb->weak_info[center] = (var > 0.2f) ? ORC_STRONG : ORC_WEAK;
    }
}

/* K12: APD.cu:2282-2344 */
void orc_confidence(orc_problem *pb) {
    const int W = pb->width, H = pb->height;
    const orc_camera &rc = pb->cameras[0];
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        const size_t center = (size_t)y * W + x;
        pb->confidence[center] = 0;
        const uint32_t sel = pb->selected_views[center];
        const float ref_depth = pb->planes[4 * center + 3];
        if (ref_depth <= 0.0f) { pb->weak_info[center] = ORC_UNKNOWN; continue; }
        const F3 fwd = point_on_world((float)x, (float)y, ref_depth, rc);
        int nc = 1;
        for (int i = 0; i < pb->params.num_images - 1; ++i) {
            if (!is_set(sel, i)) continue;
            const int s = i + 1;
            const orc_camera &sc = pb->cameras[s];
            F2 sp; float sd;
            project_on_camera(fwd, sc, sp, sd);
            const float src_depth = depth_fetch(pb, s, sp);
            if (src_depth <= 0.0f) continue;
            nc += 1;
            const F3 s3 = point_on_world(sp.x, sp.y, src_depth, sc);
            F2 bn; float rd, rrd;
            project_on_camera_parts(s3, rc, bn, rd, rrd);
            const float dc = std::fmaf(-bn.x, rrd, (float)x), dr = std::fmaf(-bn.y, rrd, (float)y); /* as built */
            const float drr = dr * dr;
            const float pd = mufu_sqrt(std::fmaf(dc, dc, drr));
            if (pd <= 2.0f) nc += 2;
            const float rdd = std::fabs(ref_depth - rd) * mufu_rcp(ref_depth);
            if (rdd <= 0.02f) nc += 2;
        }
        if (nc > 255) nc = 255;
        pb->confidence[center] = (uint8_t)nc;
    }
}

/* K13: APD.cu:2346-2432 */
void orc_local_refine(orc_problem *pb) {
    set_threads(pb);
    const int W = pb->width, H = pb->height;
    const orc_params &prm = pb->params;
    const orc_camera *cams = pb->cameras;
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        const I2 pt{x, y};
        const size_t center = (size_t)y * W + x;
        const uint8_t *vw = &pb->view_weight[ORC_MAX_IMAGES * center];
        F4 opl = normal_to_refcam(cams[0], load4(pb->planes, center));
        const float origin_depth = opl.w;
        if (origin_depth == 0) continue;
        const uint32_t sel = pb->selected_views[center];
        float cost_now = 0.0f, base_line = 0; int valid_src = 0; float weight_normal = 0.0f;
        for (int s = 1; s < prm.num_images; ++s) if (is_set(sel, s - 1)) {
            F4 tp = opl;
            tp.w = distance_to_origin(cams[0], pt, origin_depth, tp);
            float tc = ncc_old(pb, pt, s, tp);
            if (prm.geom_consistency) tc = std::fmaf(prm.geom_factor, geom_cost(pb, pt, s, tp), tc); /* FFMA as built */
            cost_now = std::fmaf((float)vw[s - 1], tc, cost_now);
            weight_normal += vw[s - 1];
            base_line += baseline_len(cams[0], cams[s]);
            valid_src++;
        }
        if (weight_normal == 0 || valid_src == 0) continue;
        const float rwn = mufu_rcp(weight_normal);
        base_line = mufu_rcp((float)valid_src) * base_line; /* as built, APD.cu:2399-2407 */
        const float fb = base_line * cams[0].K[0];
        const float disp = fb * mufu_rcp(origin_depth);
        const int radius = 5;
        float min_cost = 2.0f, best_depth = origin_depth;
        for (int pd = -radius; pd <= radius; ++pd) {
            const float p_depth = fb * mufu_rcp(disp + (float)pd);
            if (p_depth < prm.depth_min || p_depth > prm.depth_max) continue;
            F4 tp = opl;
            tp.w = distance_to_origin(cams[0], pt, p_depth, tp);
            float tc = 0.0f;
            for (int s = 1; s < prm.num_images; ++s) if (is_set(sel, s - 1)) {
                tc = std::fmaf((float)vw[s - 1], ncc_old(pb, pt, s, tp), tc); /* FFMA as built, APD.cu:2417-2419 */
                if (prm.geom_consistency) { const float g = prm.geom_factor * geom_cost(pb, pt, s, tp); tc = std::fmaf((float)vw[s - 1], g, tc); }
            }
            tc = tc * rwn;
            if (tc < min_cost) { min_cost = tc; best_depth = p_depth; }
        }
        /* "cost_now /= weight_normal; ... cost_now - min_cost > 0.1" is ONE FFMA in the reference build, compared in double */
        if ((double)std::fmaf(rwn, cost_now, -min_cost) > 0.1) pb->planes[4 * center + 3] = best_depth;
    }
}

/* K2: APD.cu:2434-2484 */
void orc_nearest_strong(orc_problem *pb) {
    set_threads(pb);
    const int W = pb->width, H = pb->height;
    const uint8_t *wk = pb->weak_info, *cf = pb->confidence;
#pragma omp parallel for schedule(dynamic, 2)
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        const size_t center = (size_t)y * W + x;
        S2 out{-1, -1};
        const uint8_t cc = cf[center];
        if (wk[center] == ORC_WEAK || wk[center] == ORC_UNKNOWN) {
            uint8_t best_conf = 0; S2 best{-1, -1}; float min_dist = FLT_MAX;
            const int radius = 100;
            for (int dx = -radius; dx <= radius; ++dx) for (int dy = -radius; dy <= radius; ++dy) {
                const int tx = x + dx, ty = y + dy;
                if (tx < 0 || tx >= W || ty < 0 || ty >= H) continue;
                const size_t tc = (size_t)ty * W + tx;
                if (wk[tc] != ORC_STRONG) continue;
                if (cf[tc] < cc) continue;
                const float d = std::sqrt((float)(dx * dx + dy * dy));
                if (d < min_dist) { min_dist = d; best = S2{(int16_t)tx, (int16_t)ty}; best_conf = cf[tc]; }
                else if (d == min_dist) { if (cf[tc] > best_conf) { best = S2{(int16_t)tx, (int16_t)ty}; best_conf = cf[tc]; } }
            }
            out = best;
        } else if (wk[center] == ORC_STRONG) {
            out = S2{(int16_t)x, (int16_t)y};
        }
        store_s2(pb->nearest_strong, center, out);
    }
}

/* K3: APD.cu:1857-2082 */
void orc_gen_anchors(orc_problem *pb) {
    set_threads(pb);
    const int width = pb->width, height = pb->height;
    const orc_params &prm = pb->params;
    const orc_camera &camera = pb->cameras[0];
#pragma omp parallel for schedule(dynamic, 2)
    for (int y = 0; y < height; ++y) for (int x = 0; x < width; ++x) {
        const I2 point{x, y};
        const size_t center = (size_t)y * width + x;
        if (pb->weak_info[center] != ORC_WEAK) continue;
        const int min_margin = 6;
        const float depth_diff = prm.depth_max - prm.depth_min;
        Rng rng(pb->seed, pb->stream, (uint32_t)center, SITE_ANCHOR);
        int16_t *anchors = &pb->anchors[center * ORC_ANCHOR_NUM * 2];
        for (int i = 0; i < ORC_ANCHOR_NUM; ++i) { anchors[2 * i] = -1; anchors[2 * i + 1] = -1; }
        anchors[0] = (int16_t)x; anchors[1] = (int16_t)y;
        S2 strong_points[32]; bool dir_valid[32];
        for (int i = 0; i < 32; ++i) { strong_points[i] = S2{-1, -1}; dir_valid[i] = false; }
        int origin_direction_index = -1, strong_point_size = 0;
        const int rotate_time = prm.rotate_time;
        const float angle = 45.0f / rotate_time;
        const float cos_angle = (float)std::cos((double)angle * kPiD / (double)180.f);
        const float sin_angle = (float)std::sin((double)angle * kPiD / (double)180.f);
        const float threshhold = (float)std::cos((double)(angle / 2.0f) * kPiD / 180.0);
        const int shift_range = ORC_MAX((int)(std::tan((double)(angle / 2.0f) * kPiD / 180.0) * 20), 1);
        const float ransac_threshold = prm.ransac_threshold;
        for (int odx = -1; odx <= 1; ++odx) for (int ody = -1; ody <= 1; ++ody) {
            if (odx == 0 && ody == 0) continue;
            F2 od{(float)odx, (float)ody};
            normalize2(&od);
            origin_direction_index++;
            for (int rot = 0; rot < rotate_time; ++rot) {
                const int dir_index = origin_direction_index * 4 + rot;
                for (int radius = 2; radius <= 4096; radius = ORC_MIN(radius * 2, radius + 25)) {
                    const F2 test_pt{point.x + od.x * radius, point.y + od.y * radius};
                    if (test_pt.x < 0 || test_pt.y < 0 || test_pt.x >= width || test_pt.y >= height) break;
                    for (int ri = 0; ri < 4; ++ri) {
                        const int rxs = rand_shift(rng, shift_range);
                        const int rys = rand_shift(rng, shift_range);
                        F2 dir{od.x * 20 + rxs, od.y * 20 + rys};
                        normalize2(&dir);
                        S2 ap{(int16_t)(int)(point.x + dir.x * radius), (int16_t)(int)(point.y + dir.y * radius)};
                        if (ap.x < min_margin || ap.y < min_margin || ap.x >= width - min_margin || ap.y >= height - min_margin) continue;
                        ap = load_s2(pb->nearest_strong, ap.x + (size_t)ap.y * width);
                        if (ap.x == -1 || ap.y == -1) continue;
                        F2 td{(float)(ap.x - point.x), (float)(ap.y - point.y)};
                        normalize2(&td);
                        const float ca = td.x * od.x + td.y * od.y;
                        if (ca > threshhold) { strong_points[dir_index] = ap; dir_valid[dir_index] = true; strong_point_size++; break; }
                    }
                    if (dir_valid[dir_index]) break;
                }
                F2 rd{od.x * cos_angle - od.y * sin_angle, od.x * sin_angle + od.y * cos_angle};
                normalize2(&rd);
                od = rd;
            }
        }
        if (strong_point_size <= 3) { pb->weak_reliable[center] = 0; continue; }
        F4 best_plane{0, 0, 0, 0};
        int ua = -1, ub = -1, uc = -1; bool has_valid = false;
        S2 spv[32]; F3 spv3[32]; int valid_count = 0; float X[3];
        get_3d_point(camera, (float)point.x, (float)point.y, pb->planes[4 * center + 3], X);
        const F3 cpw{X[0], X[1], X[2]};
        for (int i = 0; i < 32; ++i) {
            spv[i] = S2{-1, -1};
            if (dir_valid[i]) {
                const S2 sp = strong_points[i];
                const size_t spc = sp.x + (size_t)sp.y * width;
                spv[valid_count] = sp;
                get_3d_point(camera, (float)sp.x, (float)sp.y, pb->planes[4 * spc + 3], X);
                spv3[valid_count] = F3{X[0], X[1], X[2]};
                valid_count++;
            }
        }
        {
            int iteration = 50; float min_cost = FLT_MAX; int max_count = 3;
            while (iteration--) {
                const int a = rng.next() % valid_count, b = rng.next() % valid_count, c = rng.next() % valid_count;
                if (a == b || b == c || a == c) continue;
                if (!point_in_triangle(spv[a], spv[b], spv[c], point)) continue;
                const F3 &A = spv3[a], &B = spv3[b], &C = spv3[c];
                const F3 AC{A.x - C.x, A.y - C.y, A.z - C.z}, BC{B.x - C.x, B.y - C.y, B.z - C.z};
                F4 cv;
                cv.x = AC.y * BC.z - BC.y * AC.z;
                cv.y = -(AC.x * BC.z - BC.x * AC.z);
                cv.z = AC.x * BC.y - BC.x * AC.y;
                cv.w = 0;
                if ((cv.x == 0 && cv.y == 0 && cv.z == 0) || std::isnan(cv.x) || std::isnan(cv.y) || std::isnan(cv.z)) continue;
                normalize3(&cv);
                cv.w = -(cv.x * A.x + cv.y * A.y + cv.z * A.z);
                int tcnt = 0; float sdist = 0.0f;
                for (int si = 0; si < valid_count; ++si) {
                    const F3 &tp = spv3[si];
                    const float d = std::fabs(cv.x * tp.x + cv.y * tp.y + cv.z * tp.z + cv.w);
                    if (d / depth_diff < ransac_threshold) { tcnt++; sdist += d; }
                }
                if (tcnt < 6) continue;
                const float cd = std::fabs(cv.x * cpw.x + cv.y * cpw.y + cv.z * cpw.z + cv.w);
                if (tcnt > max_count) {
                    max_count = tcnt; min_cost = cd; best_plane = cv; has_valid = true; ua = a; ub = b; uc = c;
                } else if (tcnt == max_count) {
                    if (cd < min_cost) { max_count = tcnt; min_cost = cd; best_plane = cv; ua = a; ub = b; uc = c; }
                }
            }
        }
        if (!has_valid) { pb->weak_reliable[center] = 0; continue; }
        float weight[32];
        for (int i = 0; i < valid_count; ++i) {
            const F3 &tp = spv3[i];
            float d = std::fabs(best_plane.x * tp.x + best_plane.y * tp.y + best_plane.z * tp.z + best_plane.w);
            if (d / depth_diff >= ransac_threshold) { spv[i] = S2{-1, -1}; weight[i] = FLT_MAX; continue; }
            if (i == ua || i == ub || i == uc) d -= 1;
            weight[i] = d;
        }
        sort_small_weighted(spv, weight, valid_count);
        /* strong_points_valid has 32 initialised slots; anchors 1..8 <- slots 0..7 (APD.cu:2075-2080) */
        for (int i = 1; i < ORC_ANCHOR_NUM; ++i) { anchors[2 * i] = spv[i - 1].x; anchors[2 * i + 1] = spv[i - 1].y; }
        pb->weak_reliable[center] = 1;
    }
}

/* K4: APD.cu:2084-2100 */
void orc_neighbour_update(orc_problem *pb) {
    const size_t P = (size_t)pb->width * pb->height;
    for (size_t c = 0; c < P; ++c)
        if (pb->weak_info[c] == ORC_WEAK && pb->weak_reliable[c] != 1) pb->weak_info[c] = ORC_UNKNOWN;
}

/* K7: APD.cu:2486-2598 */
void orc_ransac_fit(orc_problem *pb, int iter) {
    set_threads(pb);
    const int width = pb->width, height = pb->height;
    const orc_camera &camera = pb->cameras[0];
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < height; ++y) for (int x = 0; x < width; ++x) {
        const I2 point{x, y};
        const size_t center = (size_t)y * width + x;
        if (pb->weak_info[center] != ORC_WEAK) { store4(pb->fit_planes, center, load4(pb->planes, center)); continue; }
        Rng rng(pb->seed, pb->stream, (uint32_t)center, SITE_FIT + iter);
        S2 sp[8]; F3 sp3[8]; int sc = 0; float X[3];
        for (int i = 1; i < ORC_ANCHOR_NUM; ++i) {
            const S2 tp = load_s2(pb->anchors, center * ORC_ANCHOR_NUM + i);
            if (tp.x == -1 || tp.y == -1) continue;
            sp[sc] = tp;
            const size_t tc = tp.x + (size_t)tp.y * width;
            const float depth = depth_from_plane(camera, load4(pb->planes, tc), I2{tp.x, tp.y});
            get_3d_point(camera, (float)tp.x, (float)tp.y, depth, X);
            sp3[sc] = F3{X[0], X[1], X[2]};
            sc++;
        }
        if (sc < 3) { store4(pb->fit_planes, center, load4(pb->planes, center)); continue; }
        int iteration = 50; float min_cost = FLT_MAX; F4 best{0, 0, 0, 0}; bool has_best = false;
        while (iteration--) {
            const int a = rng.next() % sc, b = rng.next() % sc, c = rng.next() % sc;
            if (a == b || b == c || a == c) continue;
            if (!point_in_triangle(sp[a], sp[b], sp[c], point)) continue;
            const F3 &A = sp3[a], &B = sp3[b], &C = sp3[c];
            const F3 AC{A.x - C.x, A.y - C.y, A.z - C.z}, BC{B.x - C.x, B.y - C.y, B.z - C.z};
            F4 cv;
            cv.x = AC.y * BC.z - BC.y * AC.z;
            cv.y = -(AC.x * BC.z - BC.x * AC.z);
            cv.z = AC.x * BC.y - BC.x * AC.y;
            cv.w = 0;
            if ((cv.x == 0 && cv.y == 0 && cv.z == 0) || std::isnan(cv.x) || std::isnan(cv.y) || std::isnan(cv.z)) continue;
            normalize3(&cv);
            cv.w = -(cv.x * A.x + cv.y * A.y + cv.z * A.z);
            float tcst = 0.0f;
            for (int si = 0; si < sc; ++si) {
                if (si == a || si == b || si == c) continue;
                const F3 &tp = sp3[si];
                tcst += std::fabs(cv.x * tp.x + cv.y * tp.y + cv.z * tp.z + cv.w);
            }
            if (tcst < min_cost) { min_cost = tcst; best = cv; has_best = true; }
            if (min_cost == 0) break;
        }
        if (has_best) {
            const float depth = depth_from_plane(camera, load4(pb->planes, center), point);
            const F4 vd = view_direction(camera, point, depth);
            const float dp = best.x * vd.x + best.y * vd.y + best.z * vd.z;
            if (dp > 0) { best.x = -best.x; best.y = -best.y; best.z = -best.z; best.w = -best.w; }
            store4(pb->fit_planes, center, best);
        } else {
            store4(pb->fit_planes, center, F4{0, 0, 0, 0});
        }
    }
}

/* APD::RunPatchMatch, APD.cu:2663-2737 */
void orc_run_pass(orc_problem *pb) {
    const orc_params &prm = pb->params;
    if (prm.use_APD) { orc_nearest_strong(pb); orc_gen_anchors(pb); orc_neighbour_update(pb); }
    orc_random_init(pb);
    for (int i = 0; i < prm.max_iterations; ++i) {
        orc_propagate_strong(pb, i, 0);
        orc_propagate_strong(pb, i, 1);
        if (prm.use_APD) { orc_ransac_fit(pb, i); orc_propagate_weak(pb, i, 0); orc_propagate_weak(pb, i, 1); }
    }
    orc_depth_normal(pb);
    orc_median_filter(pb, 0);
    orc_median_filter(pb, 1);
    orc_depth_to_weak(pb, nullptr);
    if (prm.geom_consistency || prm.use_APD) orc_confidence(pb);
    orc_local_refine(pb);
}

/* ================================================================================================
 * fusion (APD.cpp:844-910, 962-1227)
 * ============================================================================================== */
namespace {
/* ProjectCamera of the HOST code, APD.cpp:891-900: plain IEEE fp32, no contraction (g++ build of the reference) */
void f_project_on_camera(F3 P, const orc_camera &cam, F2 &pt, float &depth) {
    F3 t;
    t.x = cam.R[0] * P.x + cam.R[1] * P.y + cam.R[2] * P.z + cam.t[0];
    t.y = cam.R[3] * P.x + cam.R[4] * P.y + cam.R[5] * P.z + cam.t[1];
    t.z = cam.R[6] * P.x + cam.R[7] * P.y + cam.R[8] * P.z + cam.t[2];
    depth = cam.K[6] * t.x + cam.K[7] * t.y + cam.K[8] * t.z;
    pt.x = (cam.K[0] * t.x + cam.K[1] * t.y + cam.K[2] * t.z) / depth;
    pt.y = (cam.K[3] * t.x + cam.K[4] * t.y + cam.K[5] * t.z) / depth;
}
F3 f_point_on_world(int x, int y, float depth, const orc_camera &cam) { /* APD.cpp:866-889 */
    F3 X, T;
    X.x = depth * (x - cam.K[2]) / cam.K[0];
    X.y = depth * (y - cam.K[5]) / cam.K[4];
    X.z = depth;
    T.x = cam.R[0] * X.x + cam.R[3] * X.y + cam.R[6] * X.z;
    T.y = cam.R[1] * X.x + cam.R[4] * X.y + cam.R[7] * X.z;
    T.z = cam.R[2] * X.x + cam.R[5] * X.y + cam.R[8] * X.z;
    F3 C;
    C.x = -(cam.R[0] * cam.t[0] + cam.R[3] * cam.t[1] + cam.R[6] * cam.t[2]);
    C.y = -(cam.R[1] * cam.t[0] + cam.R[4] * cam.t[1] + cam.R[7] * cam.t[2]);
    C.z = -(cam.R[2] * cam.t[0] + cam.R[5] * cam.t[1] + cam.R[8] * cam.t[2]);
    return F3{T.x + C.x, T.y + C.y, T.z + C.z};
}
float f_angle(const float *a, const float *b) { /* APD.cpp:902-910; cv::norm is double */
    const float dot = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    const double na = std::sqrt((double)a[0] * a[0] + (double)a[1] * a[1] + (double)a[2] * a[2]);
    const double nb = std::sqrt((double)b[0] * b[0] + (double)b[1] * b[1] + (double)b[2] * b[2]);
    const float ang = std::acos((float)(dot / (na * nb)));
    if (ang != ang) return 0.0f;
    return ang;
}
/* confidences[..].at<float>(r, c) on a CV_8UC1 map (quirk 14, APD.cpp:1010-1011): 4 raw bytes at r*W + 4c;
 * bytes past the end of the map read as 0 (the reference reads out of bounds there). */
float conf_as_float(const uint8_t *conf, size_t P, int W, int r, int c) {
    uint8_t b[4] = {0, 0, 0, 0};
    const size_t off = (size_t)r * W + 4 * (size_t)c;
    for (int i = 0; i < 4; ++i) if (off + i < P) b[i] = conf[off + i];
    float f; std::memcpy(&f, b, 4);
    return f;
}
}  // namespace

/* APD.cpp:962-1049 */
void orc_weak_vis_filter(const orc_fusion_input *in, uint8_t *skip) {
    const int V = in->num_views, W = in->width, H = in->height;
    const size_t P = (size_t)W * H;
#ifdef _OPENMP
    omp_set_num_threads(in->num_threads > 0 ? in->num_threads : 1);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int ref = 0; ref < V; ++ref) {
        const orc_camera &rc = in->cameras[ref];
        for (int r = 0; r < H; ++r) for (int c = 0; c < W; ++c) {
            if (in->weaks[ref * P + (size_t)r * W + c] != ORC_WEAK) continue;
            const float ref_depth = in->depths[ref * P + (size_t)r * W + c];
            const F3 X = f_point_on_world(c, r, ref_depth, rc);
            int so = 0, wo = 0;
            for (int s = 0; s < V; ++s) {
                if (s == ref) continue;
                const orc_camera &sc = in->cameras[s];
                const float a[3] = {rc.c[0] - X.x, rc.c[1] - X.y, rc.c[2] - X.z};
                const float b[3] = {sc.c[0] - X.x, sc.c[1] - X.y, sc.c[2] - X.z};
                float ang = f_angle(a, b);
                ang = (float)((double)(ang * 180.0f) / kPiD);
                if (ang > 80.0f) continue;
                F2 pt; float pd;
                f_project_on_camera(X, sc, pt, pd);
                if (pd <= 0.0f) continue;
                const int sr = (int)(pt.y + 0.5f), scn = (int)(pt.x + 0.5f);
                if (scn >= 0 && scn < W && sr >= 0 && sr < H) {
                    const float sd = in->depths[s * P + (size_t)sr * W + scn];
                    const uint8_t wk = in->weaks[s * P + (size_t)sr * W + scn];
                    if (wk == ORC_STRONG) {
                        if (pd < sd - 0.01f * sd) so++;
                    } else if (wk == ORC_WEAK) {
                        if (conf_as_float(in->confidences + s * P, P, W, sr, scn) < conf_as_float(in->confidences + ref * P, P, W, r, c)) {
                            if (pd < sd - 0.01f * sd) wo++;
                        }
                    }
                }
            }
            if (so >= 2 || wo >= 4) skip[ref * P + (size_t)r * W + c] = 1;
        }
    }
}

/* APD.cpp:1140-1224 (greedy, order dependent) */
int64_t orc_fuse(const orc_fusion_input *in, const uint8_t *skip, float *pts, float *cols, int64_t max_points) {
    const int V = in->num_views, W = in->width, H = in->height;
    const size_t P = (size_t)W * H;
    std::vector<uint8_t> masks((size_t)V * P, 0);
    int64_t n = 0;
    for (int ref = 0; ref < V; ++ref) {
        const int nb0 = in->src_offsets[ref], nb1 = in->src_offsets[ref + 1];
        const int num_ngb = nb1 - nb0;
        std::vector<I2> used(num_ngb);
        for (int r = 0; r < H; ++r) for (int c = 0; c < W; ++c) {
            const size_t px = (size_t)r * W + c;
            if (masks[ref * P + px] == 1) continue;
            if (skip && skip[ref * P + px] == 1) continue;
            const float ref_depth = in->depths[ref * P + px];
            if (ref_depth <= 0.0) continue;
            const float *rn = &in->normals[(ref * P + px) * 3];
            const F3 X = f_point_on_world(c, r, ref_depth, in->cameras[ref]);
            int num_consistent = 0; float dyn = 0.0f;
            for (int j = 0; j < num_ngb; ++j) used[j] = I2{-1, -1};
            for (int j = 0; j < num_ngb; ++j) {
                const int s = in->src_ids[nb0 + j];
                F2 pt; float pd;
                f_project_on_camera(X, in->cameras[s], pt, pd);
                const int sr = (int)(pt.y + 0.5f), sc = (int)(pt.x + 0.5f);
                if (sc >= 0 && sc < W && sr >= 0 && sr < H) {
                    const size_t spx = (size_t)sr * W + sc;
                    if (masks[s * P + spx] == 1) continue;
                    const float sd = in->depths[s * P + spx];
                    if (sd <= 0.0) continue;
                    const float *sn = &in->normals[(s * P + spx) * 3];
                    const F3 tX = f_point_on_world(sc, sr, sd, in->cameras[s]);
                    F2 tp;
                    f_project_on_camera(tX, in->cameras[ref], tp, pd);
                    const float re = (float)std::sqrt(std::pow((double)(c - tp.x), 2) + std::pow((double)(r - tp.y), 2));
                    const float rdd = std::fabs(pd - ref_depth) / ref_depth;
                    const float ang = f_angle(rn, sn);
                    if (re < 2.0f && rdd < 0.01f && ang < 0.174533f) {
                        used[j] = I2{sc, sr};
                        const float ti = re + 200 * rdd + ang * 10;
                        dyn += (float)std::exp((double)-ti);
                        num_consistent++;
                    }
                }
            }
            const float factor = (in->weaks[ref * P + px] == ORC_WEAK ? 0.45f : 0.3f);
            if (num_consistent >= 1 && (dyn > factor * num_consistent)) {
                float col[3] = {0, 0, 0};
                if (in->colors) for (int k = 0; k < 3; ++k) col[k] = (float)in->colors[(ref * P + px) * 3 + k];
                for (int j = 0; j < num_ngb; ++j) {
                    if (used[j].x == -1) continue;
                    const int s = in->src_ids[nb0 + j];
                    const size_t spx = (size_t)used[j].y * W + used[j].x;
                    masks[s * P + spx] = 1;
                    if (in->colors) for (int k = 0; k < 3; ++k) col[k] += in->colors[(s * P + spx) * 3 + k];
                }
                for (int k = 0; k < 3; ++k) col[k] /= (num_consistent + 1);
                if (n < max_points) {
                    pts[3 * n] = X.x; pts[3 * n + 1] = X.y; pts[3 * n + 2] = X.z;
                    if (cols) { cols[3 * n] = col[0]; cols[3 * n + 1] = col[1]; cols[3 * n + 2] = col[2]; }
                }
                n++;
            }
        }
    }
    return n;
}

}  // extern "C"

/* ================================================================================================
 * Tanks-and-Temples fusion variants (APD.cpp:1229-1431 RunFusion_TAT_I, 1433-1608 RunFusion_TAT_A)
 *   variant 1 = TAT_I (distance + depth + angle thresholds growing with k, colours averaged over the used neighbours)
 *   variant 2 = TAT_A (distance + depth thresholds, reference colour only)
 * Quirk 14: diff[] is declared per VIEW, not per pixel, so a neighbour that is skipped (out of bounds, masked, invalid
 * depth) keeps the values -- and the source coordinates -- of the last pixel that did update it.
 * ============================================================================================== */
extern "C" int64_t orc_fuse_tat(const orc_fusion_input *in, const uint8_t *skip, int variant, float *pts, float *cols,
                                int64_t max_points) {
    const int V = in->num_views, W = in->width, H = in->height;
    const size_t P = (size_t)W * H;
    const float dist_base = 0.25f;
    const float depth_base = variant == 1 ? 1.0f / 3500.0f : 1.0f / 3000.0f;
    const float angle_base = 0.06981317007977318f, angle_grad = 0.05235987755982988f;
    std::vector<uint8_t> masks((size_t)V * P, 0);
    struct CostData { float dist = FLT_MAX, depth = FLT_MAX, angle = FLT_MAX; int src_r = 0, src_c = 0; bool use = false; };
    int64_t n = 0;
    for (int ref = 0; ref < V; ++ref) {
        const int nb0 = in->src_offsets[ref], num_ngb = in->src_offsets[ref + 1] - nb0;
        std::vector<CostData> diff(num_ngb);
        for (int r = 0; r < H; ++r) for (int c = 0; c < W; ++c) {
            const size_t px = (size_t)r * W + c;
            if (skip && skip[ref * P + px] == 1) continue;
            const float ref_depth = in->depths[ref * P + px];
            if (ref_depth <= 0.0) continue;
            const float *rn = &in->normals[(ref * P + px) * 3];
            const F3 X = f_point_on_world(c, r, ref_depth, in->cameras[ref]);
            for (int j = 0; j < num_ngb; ++j) {
                const int s = in->src_ids[nb0 + j];
                F2 pt; float pd;
                f_project_on_camera(X, in->cameras[s], pt, pd);
                const int sr = (int)(pt.y + 0.5f), sc = (int)(pt.x + 0.5f);
                if (sc >= 0 && sc < W && sr >= 0 && sr < H) {
                    const size_t spx = (size_t)sr * W + sc;
                    if (masks[s * P + spx] == 1) continue;
                    const float sd = in->depths[s * P + spx];
                    if (sd <= 0.0) continue;
                    const float *sn = &in->normals[(s * P + spx) * 3];
                    const F3 tX = f_point_on_world(sc, sr, sd, in->cameras[s]);
                    F2 tp;
                    f_project_on_camera(tX, in->cameras[ref], tp, pd);
                    diff[j].dist = (float)std::sqrt(std::pow((double)(c - tp.x), 2) + std::pow((double)(r - tp.y), 2));
                    diff[j].depth = std::fabs(pd - ref_depth) / ref_depth;
                    diff[j].angle = f_angle(rn, sn);
                    diff[j].src_r = sr; diff[j].src_c = sc;
                }
            }
            for (int k = 2; k <= num_ngb; ++k) {
                int count = 0;
                for (int j = 0; j < num_ngb; ++j) {
                    diff[j].use = false;
                    bool ok = diff[j].dist < k * dist_base && diff[j].depth < k * depth_base;
                    if (variant == 1) ok = ok && diff[j].angle < (k * angle_grad + angle_base);
                    if (ok) { count++; diff[j].use = true; }
                }
                if (count >= k) {
                    float col[3] = {0, 0, 0};
                    if (in->colors) for (int q = 0; q < 3; ++q) col[q] = (float)in->colors[(ref * P + px) * 3 + q];
                    if (variant == 1 && in->colors) {
                        for (int j = 0; j < num_ngb; ++j) if (diff[j].use) {
                            const int s = in->src_ids[nb0 + j];
                            const size_t spx = (size_t)diff[j].src_r * W + diff[j].src_c;
                            for (int q = 0; q < 3; ++q) col[q] += (float)in->colors[(s * P + spx) * 3 + q];
                        }
                        for (int q = 0; q < 3; ++q) col[q] /= (count + 1.0f);
                    }
                    if (n < max_points) {
                        pts[3 * n] = X.x; pts[3 * n + 1] = X.y; pts[3 * n + 2] = X.z;
                        if (cols) { cols[3 * n] = col[0]; cols[3 * n + 1] = col[1]; cols[3 * n + 2] = col[2]; }
                    }
                    n++;
                    masks[ref * P + px] = 1;
                    break;
                }
            }
        }
    }
    return n;
}
