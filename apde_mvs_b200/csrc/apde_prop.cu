// apde_prop.cu -- red/black propagation (strong AND weak pixels) as a pipeline of balanced kernels.  sm_100a.
//
// The reference runs CheckerboardPropagation{Strong,Weak} + PlaneHypothesisRefinement{Strong,Weak} (APD.cu:1098-1615,
// 950-1096) as one fat thread per pixel: 8N + N + 5N sequential cost evaluations, 1.7-2 KB of local arrays, and (measured
// on B200, r01) two problems for a one-thread-per-pixel port: warps wait for their pixel with the most selected views,
// and the weak kernel has so few pixels that a launch is a handful of long serial threads (4 % of the texture roofline).
// Here a half-sweep is
//   P0  k_prop_candidates   per pixel : the 8 candidate positions (adaptive checkerboard / anchors)
//   P1  k_prop_eval1        per (pixel, view) COLUMN, all N views: 8 candidate costs + the current plane's cost
//   P2  k_prop_select       per pixel : joint view selection, candidate choice, refinement hypotheses (counter RNG)
//       prefix sum + scatter: number the (pixel, selected view) columns view-major
//   P3  k_prop_eval3        per (pixel, SELECTED view) column: the refinement hypotheses (+ geometric term)
//   P4  k_prop_final        per pixel : weighted sums in the reference order, accept / write back
// Every evaluation kernel has one thread per column: all lanes of a warp work on the SAME source view, on neighbouring
// pixels, with identical trip counts; the parallelism is N x larger than one thread per pixel.  Sums over views are formed
// per pixel in ascending view order, so results are identical to the per-pixel formulation (zero-weight views contribute
// exactly 0 in the reference's sums: costs are finite by construction).
#include <cub/cub.cuh>

#include <cfloat>

#include "apde_common.cuh"
#include "apde_kernels.h"

namespace apde {

constexpr int kH1 = 9;  // phase-1 hypotheses per column: 8 candidates + the current plane

__device__ __forceinline__ bool prop_pixel(const PassK &K, const PropK &B, int pix, int &px, int &py, int &center) {
    if (pix >= *B.count) return false;
    center = B.list[pix];
    px = center % K.W;
    py = center / K.W;
    return true;
}

// ------------------------------------------------------------------------------------------------ P0 candidates
template <bool WEAK, bool U>
__device__ __forceinline__ void k_prop_candidates_body(const PassK &K, const PropK &B) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    int px, py, center;
    if (!prop_pixel(K, B, pix, px, py, center)) return;
    int pos[8];
    unsigned flags = 0, valid = 0;
    if (!WEAK) {
        flags = checkerboard_candidates(K.costs, K.W, K.H, px, py, pos);
    } else {
        // candidates = the planes of the 8 anchors that are still STRONG (APD.cu:1471-1482)
        const short2 *anc = K.anchors + (size_t)center * APDE_ANCHOR_NUM;
#pragma unroll
        for (int h = 0; h < 8; ++h) {
            const short2 a = anc[h + 1];
            pos[h] = 0;
            if (!(a.x == -1 || a.y == -1)) {
                valid |= 1u << h;
                pos[h] = a.x + a.y * K.W;
                if (K.weak[pos[h]] == APDE_STRONG) flags |= 1u << h;
            }
        }
    }
#pragma unroll
    for (int h = 0; h < 8; ++h) B.cand_pos[(size_t)h * B.cap + pix] = pos[h];
    B.cand_flags[pix] = flags | (valid << 8);
    // the reference patch is gathered ONCE per pixel and half-sweep; the column kernels re-load it with coalesced LDGs
    // (one column per thread would otherwise spend 36 texture fetches per column on it)
    RefPatch rp;
    load_ref_patch<U>(K, px, py, rp);
#pragma unroll
    for (int k = 0; k < kPatch; ++k) B.refpatch[(size_t)k * B.cap + pix] = rp.r[k];
    B.refpatch[(size_t)kPatch * B.cap + pix] = rp.mean;
    B.refpatch[(size_t)(kPatch + 1) * B.cap + pix] = rp.var;
    if (WEAK) {
        // the reference side of the anchor patches depends on the pixel only: gathered once here (72 texture fetches), re-loaded
        // with coalesced LDGs by every (pixel, view) column of P1 and P3 instead of 72 fetches per column
        AnchorRef ar;
        load_anchor_ref<U>(K, K.anchors + (size_t)center * APDE_ANCHOR_NUM, ar);
#pragma unroll
        for (int k = 0; k < 72; ++k) B.anchorref[(size_t)k * B.cap + pix] = ar.r[k];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            B.anchorref[(size_t)(72 + k) * B.cap + pix] = ar.mean[k];
            B.anchorref[(size_t)(80 + k) * B.cap + pix] = ar.var[k];
            B.anchor_xy[(size_t)k * B.cap + pix] = ((int)ar.a[k].x & 0xffff) | ((int)ar.a[k].y << 16);
            if (B.akey_in) {  // anchor-sorted pipeline: key = 8x8 tile of the anchor in row-major tile order
                const bool on = ar.a[k].x != -1;
                B.akey_in[(size_t)pix * 8 + k] = on ? (unsigned)((ar.a[k].y >> 3) * ((K.W + 7) >> 3) + (ar.a[k].x >> 3)) : 0xffffffffu;
                B.aitem_in[(size_t)pix * 8 + k] = pix * 8 + k;
            }
        }
    }
}
template <bool WEAK>
__global__ void __launch_bounds__(128) k_prop_candidates(const __grid_constant__ PassK K, const __grid_constant__ PropK B) {
    if (K.tex_unorm > 0.0f) k_prop_candidates_body<WEAK, true>(K, B);
    else k_prop_candidates_body<WEAK, false>(K, B);
}


__device__ __forceinline__ void load_ref_patch_g(const PropK &B, int pix, RefPatch &rp) {
#pragma unroll
    for (int k = 0; k < kPatch; ++k) rp.r[k] = B.refpatch[(size_t)k * B.cap + pix];
    rp.mean = B.refpatch[(size_t)kPatch * B.cap + pix];
    rp.var = B.refpatch[(size_t)(kPatch + 1) * B.cap + pix];
}

__device__ __forceinline__ void load_anchor_ref_g(const PropK &B, int pix, AnchorRef &ar) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int xy = B.anchor_xy[(size_t)k * B.cap + pix];
        ar.a[k] = make_short2((short)(xy & 0xffff), (short)(xy >> 16));
        ar.mean[k] = B.anchorref[(size_t)(72 + k) * B.cap + pix];
        ar.var[k] = B.anchorref[(size_t)(80 + k) * B.cap + pix];
    }
#pragma unroll
    for (int k = 0; k < 72; ++k) ar.r[k] = B.anchorref[(size_t)k * B.cap + pix];
}

// ------------------------------------------------------------------------------------------------ P1 phase-1 columns
template <bool WEAK, bool U, bool SA>
__device__ __forceinline__ void k_prop_eval1_body(const PassK &K, const PropK &B) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y;
    int px, py, center;
    if (!prop_pixel(K, B, pix, px, py, center)) return;
    const int N = K.N;
    const ViewK &vk = K.v[v];  // warp-uniform: served from the constant bank
    RefPatch rp;
    load_ref_patch_g(B, pix, rp);
    typename SaTypes<SA>::Anchors ar;
    if (WEAK) load_anchor_ref_g(B, pix, ar);
    typename SaTypes<SA>::Info si;
    load_sa<SA>(K, px, py, rp, si);
    if constexpr (SA) { if (WEAK) anchor_ref_apply_labels(K, ar, si); }
    const unsigned flags = B.cand_flags[pix] & 0xffu;
    unsigned n_eval = 0;
#pragma unroll 1
    for (int h = 0; h < kH1; ++h) {
        float c;
        if (h == 8 || ((flags >> h) & 1u)) {
            const float4 pl = (h == 8) ? K.planes[center] : K.planes[B.cand_pos[(size_t)h * B.cap + pix]];
            const PlaneM m = plane_row(K, pl);
            c = WEAK ? ncc_new_x<U, SA>(K, vk, v, px, py, m, rp, ar, si) : ncc_old_x<U, SA>(K, vk, px, py, m, rp, si);
            n_eval++;
        } else {
            c = (h == 0 && v == 0) ? 2.0f : 0.0f;  // quirk 2: "cost_array[8][32] = {2.0f}" zero-fills all but [0][0]
        }
        B.cost1[((size_t)h * N + v) * B.cap + pix] = c;
    }
    count_evals(K, WEAK ? 0 : n_eval, WEAK ? n_eval : 0, 0);
}
template <bool WEAK, bool SA>
__global__ void __launch_bounds__(128) k_prop_eval1(const __grid_constant__ PassK K, const __grid_constant__ PropK B) {
    if (K.tex_unorm > 0.0f) k_prop_eval1_body<WEAK, true, SA>(K, B);
    else k_prop_eval1_body<WEAK, false, SA>(K, B);
}


// ------------------------------------------------------------------------------------------------ P2 selection
// joint view selection (APD.cu:1339-1386 / 1505-1552) on the stored phase-1 costs
template <int NBR>
__device__ __forceinline__ uint4 select_views_g(const PassK &K, const float *sc, size_t stride, const uint32_t (&nbr_sel)[NBR],
                                                unsigned nbr_valid, int iter, Rng &rng, uint32_t *selmask, float *wnorm) {
    const int N = K.N;
    const float cost_threshold = (float)(0.8 * (double)__expf((float)(iter * iter) / (-90.0f)));
    const float fallback = __expf(cost_threshold * cost_threshold / (-0.32f));
    float sp[kMaxSrc];
    float psum = 0.0f;
    for (int v = 0; v < N; ++v) {
        float cnt = 0.0f, tmpw = 0.0f;
        int cnt_false = 0;
#pragma unroll
        for (int h = 0; h < 8; ++h) {
            const float c = sc[((size_t)h * N + v) * stride];
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
        sp[v] = p;
        psum += p;
    }
    const float inv = 1.0f / psum;  // TransformPDFToCDF, APD.cu:174-188
    float cum = 0.0f;
    for (int v = 0; v < N; ++v) { cum += sp[v] * inv; sp[v] = cum; }
    uint4 w = make_uint4(0, 0, 0, 0);
    for (int s = 0; s < 15; ++s) {
        const float r = rng.uniform() - FLT_EPSILON;
        for (int v = 0; v < N; ++v) {
            if (sp[v] > r) { vw_inc(w, v); break; }
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

// the five random refinement hypotheses built from (plane_now, depth_now); draws come from the caller (APD.cu:968-980)
__device__ __forceinline__ void refinement_set(const PassK &K, int px, int py, float4 plane_now, float depth_now, float depth_rand,
                                               float4 n_rand, float u_pert, float u1, float u2, float u3, float4 *out /* [5] */) {
    const float lo = (1.0f - 0.02f) * depth_now, hi = (1.0f + 0.02f) * depth_now;
    const float depth_pert = u_pert * (hi - lo) + lo;  // the do-while can never repeat (quirk 6)
    // GeneratePerturbedNormal (APD.cu:270-305) with the three uniforms already drawn
    const float perturbation = (float)(0.02 * 3.14159265358979323846);
    const float3 vd = view_direction(K, px, py, 1.0f);
    const float a1 = (u1 - 0.5f) * perturbation, a2 = (u2 - 0.5f) * perturbation, a3 = (u3 - 0.5f) * perturbation;
    float s1, c1, s2, c2, s3, c3;
    sincosf(a1, &s1, &c1); sincosf(a2, &s2, &c2); sincosf(a3, &s3, &c3);
    const float r0 = c2 * c3, r1 = c3 * s1 * s2 - c1 * s3, r2 = s1 * s3 + c1 * c3 * s2;
    const float r3 = c2 * s3, r4 = c1 * c3 + s1 * s2 * s3, r5 = c1 * s2 * s3 - c3 * s1;
    const float r6 = -s2, r7 = c2 * s1, r8 = c1 * c2;
    float4 n_pert = make_float4(r0 * plane_now.x + r1 * plane_now.y + r2 * plane_now.z, r3 * plane_now.x + r4 * plane_now.y + r5 * plane_now.z,
                                r6 * plane_now.x + r7 * plane_now.y + r8 * plane_now.z, 0.0f);
    if (n_pert.x * vd.x + n_pert.y * vd.y + n_pert.z * vd.z >= 0.0f) n_pert = plane_now;
    normalize3(n_pert);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        float4 tp = (i == 0 || i == 4) ? plane_now : (i == 3 ? n_pert : n_rand);
        const float d = (i == 0 || i == 2) ? depth_rand : (i == 4 ? depth_pert : depth_now);
        tp.w = distance_to_origin(K, px, py, d, tp);
        out[i] = tp;
    }
}

template <bool WEAK>
__global__ void __launch_bounds__(128) k_prop_select(const __grid_constant__ PassK K, const __grid_constant__ PropK B, int iter) {
    extern __shared__ float smem[];
    const ViewK *s_vk = stage_views(K, smem);
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    int px, py, center;
    if (!prop_pixel(K, B, pix, px, py, center)) return;
    const int W = K.W, N = K.N;
    const size_t cap = B.cap;
    const float *sc = B.cost1 + pix;
    const unsigned cf = B.cand_flags[pix];
    const unsigned flags = cf & 0xffu;
    unsigned n_geom = 0;

    Rng rng(K.seed, K.stream, (uint32_t)center, (WEAK ? SITE_WEAK : SITE_STRONG) + iter);
    uint32_t wmask;
    float wnorm;
    uint4 w;
    if (!WEAK) {
        uint32_t nbr_sel[4];
        const unsigned nbr_valid = ((flags >> 0) & 1u) | (((flags >> 2) & 1u) << 1) | (((flags >> 4) & 1u) << 2) |
                                   (((flags >> 6) & 1u) << 3);
        nbr_sel[0] = (nbr_valid & 1u) ? K.sel[center - W] : 0u;
        nbr_sel[1] = (nbr_valid & 2u) ? K.sel[center + W] : 0u;
        nbr_sel[2] = (nbr_valid & 4u) ? K.sel[center - 1] : 0u;
        nbr_sel[3] = (nbr_valid & 8u) ? K.sel[center + 1] : 0u;
        w = select_views_g<4>(K, sc, cap, nbr_sel, nbr_valid, iter, rng, &wmask, &wnorm);
    } else {
        uint32_t nbr_sel[8];
        const unsigned anchor_valid = (cf >> 8) & 0xffu;
#pragma unroll
        for (int h = 0; h < 8; ++h) nbr_sel[h] = ((anchor_valid >> h) & 1u) ? K.sel[B.cand_pos[(size_t)h * cap + pix]] : 0u;
        w = select_views_g<8>(K, sc, cap, nbr_sel, anchor_valid, iter, rng, &wmask, &wnorm);
    }
    K.vw[center] = w;
    for (int v = 0; v < N; ++v) B.flags3[(size_t)v * cap + pix] = (wmask >> v) & 1u;

    // weighted candidate costs; weak candidates carry the geometric term whenever geom_consistency (quirk 4)
    const bool geom_cand = WEAK && K.geom;
    const bool geom_now = WEAK ? (K.geom != 0) : (K.geom && K.impetus);
    float fc_min = 0.0f;
    int min_idx = 0;
#pragma unroll 1
    for (int h = 0; h < 8; ++h) {
        const bool fl = (flags >> h) & 1u;
        float4 pl = make_float4(0, 0, 0, 0);
        if (geom_cand && fl) pl = K.planes[B.cand_pos[(size_t)h * cap + pix]];
        float acc = 0.0f;
        for (uint32_t mk = wmask; mk; mk &= mk - 1) {
            const int v = __ffs(mk) - 1;
            const float c = sc[((size_t)h * N + v) * cap];
            const float wv = (float)vw_get(w, v);
            if (geom_cand) {
                if (fl) { acc += wv * (c + K.geom_factor * geom_cost(K, s_vk[v], v, px, py, pl)); n_geom++; }
                else acc += wv * (c + K.geom_factor * 3.0f);
            } else {
                acc += wv * c;
            }
        }
        const float fc = acc / wnorm;
        if (h == 0 || fc <= fc_min) { fc_min = fc; min_idx = h; }  // FindMinCostIndex: ties -> last (quirk 3)
    }
    // current hypothesis on the selected views (its costs were evaluated as hypothesis 8 of phase 1)
    const float4 plane_c = K.planes[center];
    float cost_now;
    {
        float acc = 0.0f;
        for (uint32_t mk = wmask; mk; mk &= mk - 1) {
            const int v = __ffs(mk) - 1;
            float c = sc[((size_t)8 * N + v) * cap];
            if (geom_now) { c = c + K.geom_factor * geom_cost(K, s_vk[v], v, px, py, plane_c); n_geom++; }
            acc += (float)vw_get(w, v) * c;
        }
        cost_now = acc / wnorm;
    }
    const float cost_written = cost_now;  // costs[center] is overwritten before the REFINE_INIT test (quirk 5)
    float4 plane_now = plane_c;
    float depth_now = depth_from_plane(K, plane_c, px, py);
    if ((flags >> min_idx) & 1u) {
        const float4 cand = K.planes[B.cand_pos[(size_t)min_idx * cap + pix]];
        const float db = depth_from_plane(K, cand, px, py);
        if (db >= K.depth_min && db <= K.depth_max && fc_min < cost_now) {
            depth_now = db; plane_now = cand; cost_now = fc_min;
            K.sel[center] = wmask;
        }
    }
    B.plane_now[pix] = plane_now;
    B.depth_now[pix] = depth_now;
    B.cost_now[pix] = cost_now;
    B.cost_written[pix] = cost_written;
    B.wnorm[pix] = wnorm;
    B.wmask[pix] = wmask;

    // refinement hypotheses.  All draws happen here so that the counter RNG stream of (pixel, site) stays sequential.
    int nh = 5;
    float4 fitp = make_float4(0, 0, 0, 0);
    if (WEAK) {
        fitp = K.fit[center];
        if (fitp.x == 0 && fitp.y == 0 && fitp.z == 0) nh = 0;  // a zero fit plane skips the random refinement too (APD.cu:1028)
        else nh = 11;
    }
    B.nh[pix] = (uint8_t)nh;
    if (nh == 0) {
        for (int v = 0; v < N; ++v) B.flags3[(size_t)v * cap + pix] = 0;  // no phase-3 columns for this pixel
    } else {
        const float depth_rand = rng.uniform() * (K.depth_max - K.depth_min) + K.depth_min;
        const float4 n_rand = random_normal(K, px, py, rng, depth_now);  // direction only: independent of the depth's value
        const float u_pert = rng.uniform();
        const float u1 = rng.uniform(), u2 = rng.uniform(), u3 = rng.uniform();
        float4 hyp[5];
        if (!WEAK) {
            refinement_set(K, px, py, plane_now, depth_now, depth_rand, n_rand, u_pert, u1, u2, u3, hyp);
#pragma unroll
            for (int i = 0; i < 5; ++i) B.hyp[(size_t)i * cap + pix] = hyp[i];
        } else {
            // the five random hypotheses depend on whether the fit plane gets accepted first (APD.cu:1046-1067): both
            // variants are evaluated, P4 picks the set that matches the outcome of the fit test
            B.hyp[pix] = fitp;
            refinement_set(K, px, py, plane_now, depth_now, depth_rand, n_rand, u_pert, u1, u2, u3, hyp);
#pragma unroll
            for (int i = 0; i < 5; ++i) B.hyp[(size_t)(1 + i) * cap + pix] = hyp[i];
            const float dfit = depth_from_plane(K, fitp, px, py);
            refinement_set(K, px, py, fitp, dfit, depth_rand, n_rand, u_pert, u1, u2, u3, hyp);
#pragma unroll
            for (int i = 0; i < 5; ++i) B.hyp[(size_t)(6 + i) * cap + pix] = hyp[i];
        }
    }
    count_evals(K, 0, 0, n_geom);
}

// column number -> flat (view, pixel) index
__global__ void __launch_bounds__(256) k_prop_scatter(const int *__restrict__ flags, const int *__restrict__ colidx, size_t nflat,
                                                      int *__restrict__ colmap) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nflat && flags[i]) colmap[colidx[i]] = (int)i;
}

// ------------------------------------------------------------------------------------------------ P3 phase-3 columns
// mode (weak only): 0 = all hypotheses in one launch (slot 8 duplicates slot 3 and is skipped); 1 = the fit plane only
// (slot 0); 2 = the five random hypotheses that follow the outcome of the fit test (slots base3 .. base3 + 4)
template <bool WEAK, bool U, bool SA>
__device__ __forceinline__ void k_prop_eval3_body(const PassK &K, const PropK &B, int mode) {
    extern __shared__ float smem[];
    const ViewK *s_vk = stage_views(K, smem);
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t cap = B.cap;
    const size_t nflat = (size_t)K.N * cap;
    if (col >= B.colidx3[nflat]) return;  // total number of columns = the scan's last element
    const int flat = B.colmap3[col];
    const int v = flat / (int)cap, pix = flat % (int)cap;
    const int center = B.list[pix];
    const int px = center % K.W, py = center / K.W;
    const ViewK &vk = s_vk[v];
    RefPatch rp;
    load_ref_patch_g(B, pix, rp);
    typename SaTypes<SA>::Anchors ar;
    if (WEAK) load_anchor_ref_g(B, pix, ar);
    typename SaTypes<SA>::Info si;
    load_sa<SA>(K, px, py, rp, si);
    if constexpr (SA) { if (WEAK) anchor_ref_apply_labels(K, ar, si); }
    const bool geom = WEAK ? (K.geom != 0) : (K.geom && K.impetus);
    int i0 = 0, nh = WEAK ? 11 : 5;
    if (WEAK && mode == 1) nh = 1;
    if (WEAK && mode == 2) { i0 = B.base3[pix]; nh = i0 + 5; }
    unsigned n_eval = 0, n_geom = 0;
#pragma unroll 1
    for (int i = i0; i < nh; ++i) {
        if (WEAK && mode == 0 && i == 8) continue;  // (random depth, random normal): the same plane as slot 3
        const float4 tp = B.hyp[(size_t)i * cap + pix];
        const PlaneM m = plane_row(K, tp);
        float c = WEAK ? ncc_new_x<U, SA>(K, vk, v, px, py, m, rp, ar, si) : ncc_old_x<U, SA>(K, vk, px, py, m, rp, si);
        n_eval++;
        if (geom) { c = c + K.geom_factor * geom_cost(K, vk, v, px, py, tp); n_geom++; }
        B.cost3[(size_t)i * nflat + col] = c;
    }
    count_evals(K, WEAK ? 0 : n_eval, WEAK ? n_eval : 0, n_geom);
}
template <bool WEAK, bool SA>
__global__ void __launch_bounds__(128) k_prop_eval3(const __grid_constant__ PassK K, const __grid_constant__ PropK B, int mode) {
    if (K.tex_unorm > 0.0f) k_prop_eval3_body<WEAK, true, SA>(K, B, mode);
    else k_prop_eval3_body<WEAK, false, SA>(K, B, mode);
}

// weak pixels, between the two refinement rounds: does the fit plane get accepted (APD.cu:1046-1052)?  The answer selects
// which set of five random hypotheses is evaluated; k_prop_final repeats the same test on the same numbers.
__global__ void __launch_bounds__(128) k_prop_fit_test(const __grid_constant__ PassK K, const __grid_constant__ PropK B) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    int px, py, center;
    if (!prop_pixel(K, B, pix, px, py, center)) return;
    if (B.nh[pix] == 0) { B.base3[pix] = 1; return; }
    const size_t cap = B.cap;
    const uint32_t wmask = B.wmask[pix];
    const uint4 w = K.vw[center];
    float acc = 0.0f;
    for (uint32_t mk = wmask; mk; mk &= mk - 1) {
        const int v = __ffs(mk) - 1;
        acc = __fmaf_rn((float)vw_get(w, v), B.cost3[B.colidx3[(size_t)v * cap + pix]], acc);  // == weighted(0) of k_prop_final
    }
    const float tc = __fdividef(acc, B.wnorm[pix]);
    const float db = depth_from_plane(K, B.hyp[pix], px, py);
    B.base3[pix] = (db >= K.depth_min && db <= K.depth_max && tc < B.cost_now[pix]) ? 6 : 1;
}


// ------------------------------------------------------------------------------------------------ P4 final decision
template <bool WEAK>
__global__ void __launch_bounds__(128) k_prop_final(const __grid_constant__ PassK K, const __grid_constant__ PropK B, int one_round) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    int px, py, center;
    if (!prop_pixel(K, B, pix, px, py, center)) return;
    const size_t cap = B.cap;
    const size_t nflat = (size_t)K.N * cap;
    float4 plane_now = B.plane_now[pix];
    float depth_now = B.depth_now[pix], cost_now = B.cost_now[pix];
    const float cost_written = B.cost_written[pix], wnorm = B.wnorm[pix];
    const uint32_t wmask = B.wmask[pix];
    const uint4 w = K.vw[center];
    const float dmin = K.depth_min, dmax = K.depth_max;
    const int nh = B.nh[pix];
    auto weighted = [&](int i) {
        const int slot = (WEAK && one_round && i == 8) ? 3 : i;  // slot 8 is the same plane as slot 3 and is not evaluated twice
        float acc = 0.0f;
        for (uint32_t mk = wmask; mk; mk &= mk - 1) {
            const int v = __ffs(mk) - 1;
            acc = __fmaf_rn((float)vw_get(w, v), B.cost3[(size_t)slot * nflat + B.colidx3[(size_t)v * cap + pix]], acc);
        }
        return __fdividef(acc, wnorm);
    };
    auto try_hyp = [&](int i) {
        const float4 tp = B.hyp[(size_t)i * cap + pix];
        const float tc = weighted(i);
        const float db = depth_from_plane(K, tp, px, py);
        if (db >= dmin && db <= dmax && tc < cost_now) { depth_now = db; plane_now = tp; cost_now = tc; return true; }
        return false;
    };
    if (!WEAK) {
        for (int i = 0; i < 5; ++i) try_hyp(i);
    } else if (nh > 0) {
        const bool fit_taken = try_hyp(0);
        const int base = fit_taken ? 6 : 1;
        for (int i = 0; i < 5; ++i) try_hyp(base + i);
    }
    if (K.state == APDE_REFINE_INIT) {
        if ((double)cost_now < (double)cost_written - 0.1) { K.costs[center] = cost_now; K.planes[center] = plane_now; }
        else K.costs[center] = cost_written;
    } else {
        K.costs[center] = cost_now;
        K.planes[center] = plane_now;
    }
}

// ------------------------------------------------------------------------------------------------ weak pixels: anchor-sorted pipeline
// The deformable cost of a WEAK pixel (ComputeBilateralNCCNew, APD.cu:448-593) is one 6x6 centre patch plus up to eight 3x3
// anchor patches, and the anchors are STRONG pixels up to hundreds of pixels away in eight directions.  With one thread per
// (pixel, view) column the 32 lanes of a warp gather 32 unrelated anchor patches per texture instruction: measured r01 / r02,
// ~30 % of the gather rate on the anchors' samples, which are two thirds of all samples.  Neighbouring WEAK pixels share
// anchors (the nearest STRONG pixels around their blob), so here a half-sweep
//   * sorts its (pixel, anchor slot) pairs by the 8x8 image tile the anchor lies in (once per half-sweep, P0),
//   * evaluates the CENTRE patches per (pixel, view) as before               -> centre cost, or "centre outside the image",
//   * evaluates the ANCHOR patches per sorted (pixel, anchor) pair and view  -> one cost per (hypothesis, anchor, view, pixel),
//   * and combines them per (pixel, view) in anchor-slot order with the reference's softmax mix.
// Every cost is computed by the same device functions on the same operands as in ncc_new(), and combined in the same order:
// the maps are bit-identical to the column kernels above (tools/ab_r02_weak.sh).
constexpr float kAnchorAbsent = -1.0f;  // the anchor does not take part (outside the source image and not selected there)
constexpr float kCentreOutside = -1.0f; // the pixel itself projects outside the source image: the cost is 2 whatever the anchors

template <bool U>
__device__ __forceinline__ float weak_centre_cost(const PassK &K, const ViewK &vk, int px, int py, const PlaneM &m, const RefPatch &rp) {
    const Homog Hm = make_homography(K, vk, m);
    float ptx, pty;
    project_point(Hm.h, (float)px, (float)py, ptx, pty);
    if (ptx >= (float)K.W || ptx < 0.0f || pty >= (float)K.H || pty < 0.0f) return kCentreOutside;
    return patch_ncc36<U>(K, Hm, vk.layer, px, py, rp);
}
// one anchor of one hypothesis: APD.cu:499-512 (projection test, selected-view rule) and :514-563 (the 3x3 patch)
template <bool U>
__device__ __forceinline__ float weak_anchor_cost(const PassK &K, const ViewK &vk, int view_bit, short2 a, const PlaneM &m, const float *r9,
                                                  float mean_r, float var_r, unsigned &n_sampled) {
    const Homog Hm = make_homography(K, vk, m);
    float ax, ay;
    project_point(Hm.h, (float)a.x, (float)a.y, ax, ay);
    if (ax < 0.0f || ay < 0.0f || ax >= (float)K.W || ay >= (float)K.H)
        return ((K.sel[a.x + a.y * K.W] >> view_bit) & 1u) ? 2.0f : kAnchorAbsent;
    n_sampled += var_r >= 1e-5f;  // (a texture-less anchor patch decides its cost before any source sample, see patch_ncc9)
    return patch_ncc9<U>(K, Hm, vk.layer, a.x, a.y, r9, mean_r, var_r);
}
__device__ __forceinline__ float weak_mix(float centre, const float *ac /* [8] costs by anchor slot */, const int *xy /* [8] */) {
    if (centre < 0.0f) return 2.0f;
    float sc[8];
    int ns = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if ((short)(xy[k] & 0xffff) == -1) continue;
        const float c = ac[k];
        if (c < 0.0f) continue;
        sc[ns++] = c;
    }
    return focal_mix(centre, sc, ns);
}

// phase 1, centre patches: thread = (pixel, view), the nine hypotheses in turn
template <bool U>
__device__ __forceinline__ void k_weak_center1_body(const PassK &K, const PropK &B) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y;
    int px, py, center;
    if (!prop_pixel(K, B, pix, px, py, center)) return;
    const int N = K.N;
    const ViewK &vk = K.v[v];
    RefPatch rp;
    load_ref_patch_g(B, pix, rp);
    const unsigned flags = B.cand_flags[pix] & 0xffu;
    unsigned n_eval = 0;
#pragma unroll 1
    for (int h = 0; h < kH1; ++h) {
        float c;
        if (h == 8 || ((flags >> h) & 1u)) {
            const float4 pl = (h == 8) ? K.planes[center] : K.planes[B.cand_pos[(size_t)h * B.cap + pix]];
            c = weak_centre_cost<U>(K, vk, px, py, plane_row(K, pl), rp);
            n_eval++;
        } else {
            c = (h == 0 && v == 0) ? 2.0f : 0.0f;  // quirk 2 (see k_prop_eval1)
        }
        B.cost1[((size_t)h * N + v) * B.cap + pix] = c;
    }
    count_evals(K, 0, n_eval, 0);
}
__global__ void __launch_bounds__(128) k_weak_center1(const __grid_constant__ PassK K, const __grid_constant__ PropK B) {
    if (K.tex_unorm > 0.0f) k_weak_center1_body<true>(K, B);
    else k_weak_center1_body<false>(K, B);
}

// reference side of one anchor patch, gathered through the texture unit (neighbouring threads hold neighbouring anchors):
// the texels, sums and statistics of load_anchor_ref()
template <bool U>
__device__ __forceinline__ void anchor_ref9(const PassK &K, short2 a, float *r9, float &mean_r, float &var_r) {
    float sr = 0.0f, srr = 0.0f;
    int t = 0;
#pragma unroll
    for (int i = -5; i <= 5; i += 5) {
#pragma unroll
        for (int j = -5; j <= 5; j += 5) {
            const float r = fetch<U>(K, (float)(a.x + i) + 0.5f, (float)(a.y + j) + 0.5f, K.ref_layer);
            r9[t++] = r;
            sr = __fadd_rn(sr, r);
            srr = __fmaf_rn(r, r, srr);
        }
    }
    const float inv = 1.0f / 9.0f;
    mean_r = __fmul_rn(inv, sr);
    var_r = __fmaf_rn(inv, srr, -__fmul_rn(mean_r, mean_r));
}

// phase 1, anchor patches: thread = (sorted (pixel, anchor) pair, view)
template <bool U>
__device__ __forceinline__ void k_weak_anchor1_body(const PassK &K, const PropK &B, int nitems) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nitems || B.akey[idx] == 0xffffffffu) return;
    const int v = blockIdx.y, N = K.N;
    const int item = B.aitem[idx], pix = item >> 3, k = item & 7;
    if (pix >= *B.count) return;
    const int center = B.list[pix];
    const int xy = B.anchor_xy[(size_t)k * B.cap + pix];
    const short2 a = make_short2((short)(xy & 0xffff), (short)(xy >> 16));
    const ViewK &vk = K.v[v];
    float r9[9], mean_r, var_r;
    anchor_ref9<U>(K, a, r9, mean_r, var_r);
    const unsigned flags = B.cand_flags[pix] & 0xffu;
    unsigned n_anchor = 0;
#pragma unroll 1
    for (int h = 0; h < kH1; ++h) {
        if (!(h == 8 || ((flags >> h) & 1u))) continue;
        const float4 pl = (h == 8) ? K.planes[center] : K.planes[B.cand_pos[(size_t)h * B.cap + pix]];
        B.acost1[(((size_t)h * 8 + k) * N + v) * B.cap + pix] = weak_anchor_cost<U>(K, vk, v, a, plane_row(K, pl), r9, mean_r, var_r, n_anchor);
    }
    count_evals(K, 0, 0, 0, n_anchor);
}
__global__ void __launch_bounds__(128) k_weak_anchor1(const __grid_constant__ PassK K, const __grid_constant__ PropK B, int nitems) {
    if (K.tex_unorm > 0.0f) k_weak_anchor1_body<true>(K, B, nitems);
    else k_weak_anchor1_body<false>(K, B, nitems);
}

// phase 1, combination: thread = (pixel, view)
__global__ void __launch_bounds__(128) k_weak_combine1(const __grid_constant__ PassK K, const __grid_constant__ PropK B) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y, N = K.N;
    if (pix >= *B.count) return;
    const unsigned flags = B.cand_flags[pix] & 0xffu;
    int xy[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) xy[k] = B.anchor_xy[(size_t)k * B.cap + pix];
#pragma unroll 1
    for (int h = 0; h < kH1; ++h) {
        if (!(h == 8 || ((flags >> h) & 1u))) continue;
        float ac[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) ac[k] = ((short)(xy[k] & 0xffff) == -1) ? kAnchorAbsent : B.acost1[(((size_t)h * 8 + k) * N + v) * B.cap + pix];
        float *dst = &B.cost1[((size_t)h * N + v) * B.cap + pix];
        *dst = weak_mix(*dst, ac, xy);
    }
}

// refinement hypotheses of the weak pixels by column: slots [i0, i0 + nh) of B.hyp; mode 1 = the fit plane (slot 0),
// mode 2 = the five random hypotheses that follow the outcome of the fit test (slots base3 .. base3 + 4)
__device__ __forceinline__ void weak_slots(const PropK &B, int pix, int mode, int &i0, int &nh) {
    i0 = 0; nh = 1;
    if (mode == 2) { i0 = B.base3[pix]; nh = 5; }
}

template <bool U>
__device__ __forceinline__ void k_weak_center3_body(const PassK &K, const PropK &B, int mode) {
    extern __shared__ float smem[];
    const ViewK *s_vk = stage_views(K, smem);
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t cap = B.cap, nflat = (size_t)K.N * cap;
    if (col >= B.colidx3[nflat]) return;
    const int flat = B.colmap3[col];
    const int v = flat / (int)cap, pix = flat % (int)cap;
    const int center = B.list[pix];
    const int px = center % K.W, py = center / K.W;
    RefPatch rp;
    load_ref_patch_g(B, pix, rp);
    int i0, nh;
    weak_slots(B, pix, mode, i0, nh);
    unsigned n_eval = 0;
#pragma unroll 1
    for (int i = i0; i < i0 + nh; ++i) {
        const float4 tp = B.hyp[(size_t)i * cap + pix];
        B.cost3[(size_t)i * nflat + col] = weak_centre_cost<U>(K, s_vk[v], px, py, plane_row(K, tp), rp);
        n_eval++;
    }
    count_evals(K, 0, n_eval, 0);
}
__global__ void __launch_bounds__(128) k_weak_center3(const __grid_constant__ PassK K, const __grid_constant__ PropK B, int mode) {
    if (K.tex_unorm > 0.0f) k_weak_center3_body<true>(K, B, mode);
    else k_weak_center3_body<false>(K, B, mode);
}

template <bool U>
__device__ __forceinline__ void k_weak_anchor3_body(const PassK &K, const PropK &B, int nitems, int mode) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nitems || B.akey[idx] == 0xffffffffu) return;
    const int v = blockIdx.y;
    const size_t cap = B.cap, nflat = (size_t)K.N * cap;
    const int item = B.aitem[idx], pix = item >> 3, k = item & 7;
    if (pix >= *B.count || !B.flags3[(size_t)v * cap + pix]) return;  // only the views this pixel selected have a column
    const int col = B.colidx3[(size_t)v * cap + pix];
    const int xy = B.anchor_xy[(size_t)k * cap + pix];
    const short2 a = make_short2((short)(xy & 0xffff), (short)(xy >> 16));
    const ViewK &vk = K.v[v];
    float r9[9], mean_r, var_r;
    anchor_ref9<U>(K, a, r9, mean_r, var_r);
    int i0, nh;
    weak_slots(B, pix, mode, i0, nh);
    unsigned n_anchor = 0;
#pragma unroll 1
    for (int ii = 0; ii < nh; ++ii) {
        const float4 tp = B.hyp[(size_t)(i0 + ii) * cap + pix];
        B.acost3[((size_t)ii * 8 + k) * nflat + col] = weak_anchor_cost<U>(K, vk, v, a, plane_row(K, tp), r9, mean_r, var_r, n_anchor);
    }
    count_evals(K, 0, 0, 0, n_anchor);
}
__global__ void __launch_bounds__(128) k_weak_anchor3(const __grid_constant__ PassK K, const __grid_constant__ PropK B, int nitems, int mode) {
    if (K.tex_unorm > 0.0f) k_weak_anchor3_body<true>(K, B, nitems, mode);
    else k_weak_anchor3_body<false>(K, B, nitems, mode);
}

__global__ void __launch_bounds__(128) k_weak_combine3(const __grid_constant__ PassK K, const __grid_constant__ PropK B, int mode) {
    extern __shared__ float smem[];
    const ViewK *s_vk = stage_views(K, smem);
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t cap = B.cap, nflat = (size_t)K.N * cap;
    if (col >= B.colidx3[nflat]) return;
    const int flat = B.colmap3[col];
    const int v = flat / (int)cap, pix = flat % (int)cap;
    const int center = B.list[pix];
    const int px = center % K.W, py = center / K.W;
    int xy[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) xy[k] = B.anchor_xy[(size_t)k * cap + pix];
    int i0, nh;
    weak_slots(B, pix, mode, i0, nh);
    const bool geom = K.geom != 0;  // weak pixels carry the geometric term whenever geom_consistency (quirk 4)
    unsigned n_geom = 0;
#pragma unroll 1
    for (int ii = 0; ii < nh; ++ii) {
        float ac[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) ac[k] = ((short)(xy[k] & 0xffff) == -1) ? kAnchorAbsent : B.acost3[((size_t)ii * 8 + k) * nflat + col];
        float *dst = &B.cost3[(size_t)(i0 + ii) * nflat + col];
        float c = weak_mix(*dst, ac, xy);
        if (geom) { c = c + K.geom_factor * geom_cost(K, s_vk[v], v, px, py, B.hyp[(size_t)(i0 + ii) * cap + pix]); n_geom++; }
        *dst = c;
    }
    count_evals(K, 0, 0, n_geom);
}

// ------------------------------------------------------------------------------------------------ host side
#define PCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return e_; } while (0)

cudaError_t PropWorkspace::reserve(int cap_, int N) {
    if (cap_ <= cap && N <= views) return cudaSuccess;
    release();
    cap = cap_; views = N;
    const size_t c = (size_t)cap, nflat = (size_t)N * c;
    PCU(cudaMalloc(&cand_pos, 8 * c * sizeof(int)));
    PCU(cudaMalloc(&cand_flags, c * sizeof(uint32_t)));
    PCU(cudaMalloc(&cost1, (size_t)kH1 * nflat * sizeof(float)));
    PCU(cudaMalloc(&plane_now, c * sizeof(float4)));
    PCU(cudaMalloc(&depth_now, c * sizeof(float)));
    PCU(cudaMalloc(&cost_now, c * sizeof(float)));
    PCU(cudaMalloc(&cost_written, c * sizeof(float)));
    PCU(cudaMalloc(&wnorm, c * sizeof(float)));
    PCU(cudaMalloc(&wmask, c * sizeof(uint32_t)));
    PCU(cudaMalloc(&nh, c));
    PCU(cudaMalloc(&hyp, 11 * c * sizeof(float4)));
    PCU(cudaMalloc(&refpatch, (size_t)(kPatch + 2) * c * sizeof(float)));
    PCU(cudaMalloc(&base3, c));
    PCU(cudaMalloc(&flags3, (nflat + 1) * sizeof(int)));
    PCU(cudaMalloc(&colidx3, (nflat + 1) * sizeof(int)));
    PCU(cudaMalloc(&colmap3, nflat * sizeof(int)));
    PCU(cudaMalloc(&cost3, 11 * nflat * sizeof(float)));
    scan_bytes = 0;
    PCU(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, flags3, colidx3, (int)(nflat + 1)));
    PCU(cudaMalloc(&scan_tmp, scan_bytes));
    PCU(cudaMemset(flags3, 0, (nflat + 1) * sizeof(int)));
    return cudaSuccess;
}
// anchor-sorted weak pipeline: buffers sized by the row stride of the half-sweep (the weak list length), grow-only
cudaError_t PropWorkspace::reserve_anchor_pipeline(size_t stride, int N) {
    const size_t items = stride * 8;
    if (items > aitem_cap) {
        cudaFree(akey_in); cudaFree(akey); cudaFree(aitem_in); cudaFree(aitem); cudaFree(sort_tmp);
        akey_in = akey = nullptr; aitem_in = aitem = nullptr; sort_tmp = nullptr; aitem_cap = 0;
        const size_t cap2 = items + items / 4;
        PCU(cudaMalloc(&akey_in, cap2 * sizeof(unsigned)));
        PCU(cudaMalloc(&akey, cap2 * sizeof(unsigned)));
        PCU(cudaMalloc(&aitem_in, cap2 * sizeof(int)));
        PCU(cudaMalloc(&aitem, cap2 * sizeof(int)));
        sort_bytes = 0;
        PCU(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, akey_in, akey, aitem_in, aitem, (int)cap2));
        PCU(cudaMalloc(&sort_tmp, sort_bytes));
        aitem_cap = cap2;
    }
    const size_t need = (size_t)(kH1 + 5) * 8 * N * stride;  // acost1 [9][8][N][stride] + acost3 [5][8][N * stride]
    if (need > acost_cap) {
        cudaFree(acost1);
        acost1 = acost3 = nullptr; acost_cap = 0;
        const size_t cap2 = need + need / 4;
        PCU(cudaMalloc(&acost1, cap2 * sizeof(float)));
        acost_cap = cap2;
    }
    acost3 = acost1 + (size_t)kH1 * 8 * N * stride;
    return cudaSuccess;
}

void PropWorkspace::release() {
    cudaFree(akey_in); cudaFree(akey); cudaFree(aitem_in); cudaFree(aitem); cudaFree(sort_tmp); cudaFree(acost1);
    akey_in = akey = nullptr; aitem_in = aitem = nullptr; sort_tmp = nullptr; acost1 = acost3 = nullptr; aitem_cap = acost_cap = 0;
    cudaFree(cand_pos); cudaFree(cand_flags); cudaFree(cost1); cudaFree(plane_now); cudaFree(depth_now); cudaFree(cost_now);
    cudaFree(cost_written); cudaFree(wnorm); cudaFree(wmask); cudaFree(nh); cudaFree(hyp); cudaFree(flags3); cudaFree(colidx3);
    cudaFree(colmap3); cudaFree(cost3); cudaFree(scan_tmp); cudaFree(refpatch); cudaFree(anchorref); cudaFree(anchor_xy); cudaFree(base3);
    refpatch = anchorref = nullptr; anchor_xy = nullptr; base3 = nullptr;
    cand_pos = flags3 = colidx3 = colmap3 = nullptr; cand_flags = wmask = nullptr; cost1 = depth_now = cost_now = cost_written = wnorm = cost3 = nullptr;
    plane_now = hyp = nullptr; nh = nullptr; scan_tmp = nullptr; cap = 0; views = 0;
}

template <bool WEAK>
static cudaError_t run_half_sweep(const PassK &K, PropWorkspace &ws, const int *list, const int *count, int max_pixels, int iter,
                                  cudaStream_t st, uint64_t *launches) {
    if (max_pixels <= 0) return cudaSuccess;  // empty class: no launches at all
    PropK B;
    // row stride of every per-pixel array = the list length rounded up, not the capacity: the [N][stride] flag / scan / scatter
    // passes then touch what the half-sweep really holds
    B.list = list; B.count = count; B.cap = std::min(ws.cap, (max_pixels + 127) / 128 * 128);
    B.cand_pos = ws.cand_pos; B.cand_flags = ws.cand_flags; B.cost1 = ws.cost1; B.plane_now = ws.plane_now;
    B.depth_now = ws.depth_now; B.cost_now = ws.cost_now; B.cost_written = ws.cost_written; B.wnorm = ws.wnorm; B.wmask = ws.wmask;
    B.nh = ws.nh; B.hyp = ws.hyp; B.refpatch = ws.refpatch; B.flags3 = ws.flags3; B.colidx3 = ws.colidx3; B.colmap3 = ws.colmap3; B.cost3 = ws.cost3;
    if (WEAK && !ws.anchorref) {  // only the weak class needs the anchor cache (384 B per list slot)
        PCU(cudaMalloc(&ws.anchorref, (size_t)88 * ws.cap * sizeof(float)));
        PCU(cudaMalloc(&ws.anchor_xy, (size_t)8 * ws.cap * sizeof(int)));
    }
    B.anchorref = ws.anchorref; B.anchor_xy = ws.anchor_xy; B.base3 = ws.base3;
    const int N = K.N;
    // weak pixels without a label map: the anchor-sorted pipeline (APDE_WEAK_COLUMNS=1 keeps the column kernels: A/B and parity
    // twin).  Its anchor costs take (9 + 5) x 8 x N floats per weak pixel; beyond 8 GB the column kernels run instead.
    B.akey_in = B.akey = nullptr; B.aitem_in = B.aitem = nullptr; B.acost1 = B.acost3 = nullptr;
    bool sorted = false;
    if (WEAK && !K.sa) {
        const char *e = getenv("APDE_WEAK_COLUMNS");
        const size_t bytes = (size_t)(kH1 + 5) * 8 * N * B.cap * sizeof(float);
        if (!(e && e[0] == '1') && bytes <= ((size_t)8 << 30)) {
            PCU(ws.reserve_anchor_pipeline((size_t)B.cap, N));
            B.akey_in = ws.akey_in; B.akey = ws.akey; B.aitem_in = ws.aitem_in; B.aitem = ws.aitem; B.acost1 = ws.acost1; B.acost3 = ws.acost3;
            sorted = true;
        }
    }
    const int nitems = max_pixels * 8;
    const unsigned ib = (unsigned)((nitems + 127) / 128);
    const size_t nflat = (size_t)N * B.cap;
    const unsigned pb = (unsigned)((max_pixels + 127) / 128);
    const size_t vsm = sizeof(float) * views_smem_floats(N);
    if (pb == 0) return cudaSuccess;
    if (sorted) PCU(cudaMemsetAsync(ws.akey_in, 0xff, (size_t)nitems * sizeof(unsigned), st));  // slots past the list end sort last
    k_prop_candidates<WEAK><<<pb, 128, 0, st>>>(K, B);
    if (sorted) {
        PCU(cub::DeviceRadixSort::SortPairs(ws.sort_tmp, ws.sort_bytes, ws.akey_in, ws.akey, ws.aitem_in, ws.aitem, nitems, 0, 32, st));
        k_weak_center1<<<dim3(pb, N), 128, 0, st>>>(K, B);
        k_weak_anchor1<<<dim3(ib, N), 128, 0, st>>>(K, B, nitems);
        k_weak_combine1<<<dim3(pb, N), 128, 0, st>>>(K, B);
        if (launches) *launches += 3;
    }
    // problems with a segment-label map: the <SA> twins of the two evaluation kernels (labels re-read per column, the anchor
    // cache re-masked in registers; see "segment labels" in apde_device.cuh)
    else if (K.sa) k_prop_eval1<WEAK, true><<<dim3(pb, N), 128, 0, st>>>(K, B);
    else k_prop_eval1<WEAK, false><<<dim3(pb, N), 128, 0, st>>>(K, B);
    // flags3 of pixels beyond the list (previous, longer half-sweeps) must not create columns
    PCU(cudaMemsetAsync(ws.flags3, 0, (nflat + 1) * sizeof(int), st));
    k_prop_select<WEAK><<<pb, 128, vsm, st>>>(K, B, iter);
    PCU(cub::DeviceScan::ExclusiveSum(ws.scan_tmp, ws.scan_bytes, ws.flags3, ws.colidx3, (int)(nflat + 1), st));
    k_prop_scatter<<<(unsigned)((nflat + 255) / 256), 256, 0, st>>>(ws.flags3, ws.colidx3, nflat, ws.colmap3);
    // worst case: every (pixel, view) is a column; threads beyond the device-side count exit at once
    const size_t max_cols = (size_t)max_pixels * N;
    const unsigned cb = (unsigned)((max_cols + 127) / 128);
    const char *one = getenv("APDE_WEAK_ONE_ROUND");  // A/B switch: all 10 distinct hypotheses in one launch
    const int one_round = (!WEAK || (one && one[0] == '1')) ? 1 : 0;
    if (one_round) {
        if (K.sa) k_prop_eval3<WEAK, true><<<cb, 128, vsm, st>>>(K, B, 0);
        else k_prop_eval3<WEAK, false><<<cb, 128, vsm, st>>>(K, B, 0);
        if (launches) *launches += 8;
    } else {
        // the five random hypotheses of a weak pixel depend on whether its fit plane is accepted (APD.cu:1046-1067): evaluate
        // the fit plane, decide, then evaluate only the set that applies -- 6 deformable evaluations per column instead of 10
        auto eval3 = [&](int mode) {
            if (sorted) {
                k_weak_center3<<<cb, 128, vsm, st>>>(K, B, mode);
                k_weak_anchor3<<<dim3(ib, N), 128, 0, st>>>(K, B, nitems, mode);
                k_weak_combine3<<<cb, 128, vsm, st>>>(K, B, mode);
                if (launches) *launches += 2;
            } else if (K.sa) {
                k_prop_eval3<WEAK, true><<<cb, 128, vsm, st>>>(K, B, mode);
            } else {
                k_prop_eval3<WEAK, false><<<cb, 128, vsm, st>>>(K, B, mode);
            }
        };
        eval3(1);
        k_prop_fit_test<<<pb, 128, 0, st>>>(K, B);
        eval3(2);
        if (launches) *launches += 10;
    }
    k_prop_final<WEAK><<<pb, 128, 0, st>>>(K, B, one_round);
    return cudaGetLastError();
}

cudaError_t prop_half_sweep(const PassK &K, PropWorkspace &ws, bool weak, const int *list, const int *count, int max_pixels, int iter,
                            cudaStream_t st, uint64_t *launches) {
    return weak ? run_half_sweep<true>(K, ws, list, count, max_pixels, iter, st, launches)
                : run_half_sweep<false>(K, ws, list, count, max_pixels, iter, st, launches);
}

}  // namespace apde
