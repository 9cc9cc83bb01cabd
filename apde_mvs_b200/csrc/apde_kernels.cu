// apde_kernels.cu -- the PatchMatch pass kernels (sm_100a).  One thread per pixel; red/black half-sweeps update in
// place because every neighbour read is of the opposite colour (same invariant as the reference).
//
// Each kernel cites the reference kernel it replaces.  Integer / scheduling artefacts (colouring, candidate index
// sets, tie rules, zero-initialised cost rows of invalid neighbours) follow the reference bit for bit; see
// SURVEY.md section 9 for the quirk list.
#include "apde_device.cuh"
#include "apde_kernels.h"
#include "apde_common.cuh"

#include <cfloat>
#include <cstdlib>

namespace apde {

// -------------------------------------------------------------------------------------------- K5 init
// RandomInitialization + ComputeMultiViewInitialCostandSelectedViews, APD.cu:919-948, 723-774
//
// The deformable cost of a WEAK pixel reads selected_views of its (STRONG) anchors (APD.cu:502-503) -- in the reference
// while the same kernel is still writing them, a data race.  Here the kernel runs in two phases when use_APD is set:
// phase 0 = every pixel that is not WEAK, phase 1 = the WEAK pixels, which then see their anchors' final masks.  This is
// one of the serialisations the reference's race allows, and it makes the pass deterministic.  phase < 0 = all pixels.
template <bool U, bool SA>
__device__ __forceinline__ void k_init_body(const PassK &K, int tiles_x, int phase) {
    int px, py;
    if (!full_pixel(K, tiles_x, px, py)) return;
    const int center = py * K.W + px;
    if (phase >= 0 && (int)(K.weak[center] == APDE_WEAK) != phase) return;
    float4 pl;
    if (K.state == APDE_FIRST_INIT) {
        Rng rng(K.seed, K.stream, (uint32_t)center, SITE_INIT);
        const float depth = rng.uniform() * (K.depth_max - K.depth_min) + K.depth_min;
        pl = random_normal(K, px, py, rng, depth);
        pl.w = distance_to_origin(K, px, py, depth, pl);
    } else {
        pl = normal_to_refcam(K, K.planes[center]);
        pl.w = distance_to_origin(K, px, py, pl.w, pl);
    }
    K.planes[center] = pl;

    RefPatch rp;
    load_ref_patch<U>(K, px, py, rp);
    typename SaTypes<SA>::Info si;
    load_sa<SA>(K, px, py, rp, si);
    const PlaneM m = plane_row(K, pl);
    const bool weak = K.use_apd && K.weak[center] == APDE_WEAK;
    float cv[kMaxSrc], cvc[kMaxSrc];
    int num_valid = 0;
    typename SaTypes<SA>::Anchors ar;
    if (weak) load_anchor_ref_x<U, SA>(K, K.anchors + (size_t)center * APDE_ANCHOR_NUM, ar, si);
    for (int v = 0; v < K.N; ++v) {
        const float c = weak ? ncc_new_x<U, SA>(K, K.v[v], v, px, py, m, rp, ar, si) : ncc_old_x<U, SA>(K, K.v[v], px, py, m, rp, si);
        cv[v] = c; cvc[v] = c;
        if (c < 2.0f) num_valid++;
    }
    // sort_small, APD.cu:3-12
    for (int i = 1; i < K.N; i++) {
        const float tmp = cv[i];
        int j;
        for (j = i; j >= 1 && tmp < cv[j - 1]; j--) cv[j] = cv[j - 1];
        cv[j] = tmp;
    }
    uint32_t sel = 0;
    float cost = 2.0f;
    const int top_k = min(num_valid, K.top_k);
    if (top_k > 0) {
        float acc = 0.0f;
        for (int i = 0; i < top_k; ++i) acc = __fadd_rn(acc, cv[i]);
        const float thr = cv[top_k - 1];
        for (int i = 0; i < K.N; ++i) if (cvc[i] <= thr) sel |= (1u << i);
        cost = __fmul_rn(rcp_approx((float)top_k), acc);  // "cost / top_k" as built: MUFU.RCP(float(top_k)) * cost
    }
    K.sel[center] = sel;
    K.costs[center] = cost;
    count_evals(K, weak ? 0 : K.N, weak ? K.N : 0, 0);
}
__global__ void __launch_bounds__(128) k_init(const __grid_constant__ PassK K, int tiles_x, int phase) {
    if (K.tex_unorm > 0.0f) k_init_body<true, false>(K, tiles_x, phase);
    else k_init_body<false, false>(K, tiles_x, phase);
}
// the same kernel for a problem with a segment-label map (see "segment labels" in apde_device.cuh)
__global__ void __launch_bounds__(128) k_init_sa(const __grid_constant__ PassK K, int tiles_x, int phase) {
    if (K.tex_unorm > 0.0f) k_init_body<true, true>(K, tiles_x, phase);
    else k_init_body<false, true>(K, tiles_x, phase);
}


// -------------------------------------------------------------------------------------------- K6 strong propagation
// Black/RedPixelUpdateStrong -> CheckerboardPropagationStrong -> PlaneHypothesisRefinementStrong,
// APD.cu:1654-1692, 1098-1440, 950-1006.
//
// phase 1: 8 candidates x N views (uniform view loop, costs parked in shared memory)
// phase 2: joint view selection (15 Monte-Carlo draws from the CDF)
// phase 3: current hypothesis + 5 refinement candidates, evaluated ONLY on views with non-zero weight -- zero-weight
//          views contribute exactly 0 to the reference's weighted sums (costs are finite by construction), so
//          skipping them is bit-identical and removes (N - S) of every N evaluations here.
template <bool U, bool SA>
__device__ __forceinline__ void k_prop_strong_body(const PassK &K, int iter, int color, int tiles_x,
                                                     int ylimit) {
    extern __shared__ float smem[];
    const ViewK *s_vk = stage_views(K, smem);
    int px, py;
    if (!half_pixel_or_list(K, color, tiles_x, ylimit, px, py)) return;
    const int W = K.W, N = K.N;
    const int center = py * W + px;
    if (K.weak[center] == APDE_WEAK) return;
    const int stride = blockDim.x;
    float *sc = smem + views_smem_floats(N) + threadIdx.x;  // [8N][stride]
    float *sp = sc + 8 * N * stride;                        // [N][stride]

    RefPatch rp;
    load_ref_patch<U>(K, px, py, rp);
    typename SaTypes<SA>::Info si;
    load_sa<SA>(K, px, py, rp, si);
    unsigned n_old = 0, n_geom = 0;

    int pos[8];
    const unsigned flags = checkerboard_candidates(K.costs, W, K.H, px, py, pos);
#pragma unroll 1
    for (int h = 0; h < 8; ++h) {
        if ((flags >> h) & 1u) {
            const PlaneM m = plane_row(K, K.planes[pos[h]]);
#pragma unroll 1
            for (int v = 0; v < N; ++v) sc[(h * N + v) * stride] = ncc_old_x<U, SA>(K, K.v[v], px, py, m, rp, si);
            n_old += N;
        } else {
            // quirk 2: "float cost_array[8][32] = {2.0f}" leaves every entry 0 except [0][0]
            for (int v = 0; v < N; ++v) sc[(h * N + v) * stride] = (h == 0 && v == 0) ? 2.0f : 0.0f;
        }
    }

    uint32_t nbr_sel[4];
    const unsigned nbr_valid = ((flags >> 0) & 1u) | (((flags >> 2) & 1u) << 1) | (((flags >> 4) & 1u) << 2) |
                               (((flags >> 6) & 1u) << 3);
    nbr_sel[0] = (nbr_valid & 1u) ? K.sel[center - W] : 0u;
    nbr_sel[1] = (nbr_valid & 2u) ? K.sel[center + W] : 0u;
    nbr_sel[2] = (nbr_valid & 4u) ? K.sel[center - 1] : 0u;
    nbr_sel[3] = (nbr_valid & 8u) ? K.sel[center + 1] : 0u;

    Rng rng(K.seed, K.stream, (uint32_t)center, SITE_STRONG + iter);
    uint32_t wmask;
    float wnorm;
    const uint4 w = select_views<4>(K, sc, sp, stride, nbr_sel, nbr_valid, iter, rng, &wmask, &wnorm);
    K.vw[center] = w;

    float final_costs[8];
#pragma unroll
    for (int h = 0; h < 8; ++h) {
        float acc = 0.0f;
        for (uint32_t mk = wmask; mk; mk &= mk - 1) {
            const int v = __ffs(mk) - 1;
            acc += (float)vw_get(w, v) * sc[(h * N + v) * stride];
        }
        final_costs[h] = acc / wnorm;
    }
    int min_idx = 0;
    {
        float mc = final_costs[0];
#pragma unroll
        for (int h = 1; h < 8; ++h) if (final_costs[h] <= mc) { mc = final_costs[h]; min_idx = h; }  // ties -> last (quirk 3)
    }
    float fc_min = final_costs[0];
#pragma unroll
    for (int h = 1; h < 8; ++h) if (h == min_idx) fc_min = final_costs[h];

    const bool use_geom = K.geom && K.impetus;
    const float4 plane_c = K.planes[center];
    float4 plane_now = plane_c;
    float depth_now = 0.0f, cost_now = 0.0f, cost_written = 0.0f;
    float4 cand_n[2];  // 0: random normal, 1: perturbed normal
    float depth_rand = 0.0f, depth_pert = 0.0f;
    const float dmin = K.depth_min, dmax = K.depth_max;

    // current hypothesis on the selected views
    {
        const PlaneM m = plane_row(K, plane_c);
        float acc = 0.0f;
        for (uint32_t mk = wmask; mk; mk &= mk - 1) {
            const int v = __ffs(mk) - 1;
            float c = ncc_old_x<U, SA>(K, s_vk[v], px, py, m, rp, si);
            n_old++;
            if (use_geom) { c = __fmaf_rn(K.geom_factor, geom_cost(K, s_vk[v], v, px, py, plane_c), c); n_geom++; }
            acc += (float)vw_get(w, v) * c;
        }
        cost_now = acc / wnorm;
    }
    cost_written = cost_now;
    depth_now = depth_from_plane(K, plane_c, px, py);
    if ((flags >> min_idx) & 1u) {
        const float4 cand = K.planes[pos[min_idx]];
        const float db = depth_from_plane(K, cand, px, py);
        if (db >= dmin && db <= dmax && fc_min < cost_now) {
            depth_now = db; plane_now = cand; cost_now = fc_min;
            K.sel[center] = wmask;
        }
    }
    // PlaneHypothesisRefinementStrong: all five candidates are built from the state BEFORE the loop (APD.cu:968-980)
    depth_rand = rng.uniform() * (dmax - dmin) + dmin;
    cand_n[0] = random_normal(K, px, py, rng, depth_now);
    {
        const float lo = (1.0f - 0.02f) * depth_now, hi = (1.0f + 0.02f) * depth_now;
        depth_pert = rng.uniform() * (hi - lo) + lo;  // the do-while can never repeat (quirk 6)
    }
    cand_n[1] = perturbed_normal(K, px, py, plane_now, rng, (float)(0.02 * 3.14159265358979323846));
    const float4 base_n = plane_now;
    const float base_d = depth_now;
#pragma unroll 1
    for (int i = 0; i < 5; ++i) {
        float4 tp = (i == 0 || i == 4) ? base_n : (i == 3 ? cand_n[1] : cand_n[0]);
        const float d = (i == 0 || i == 2) ? depth_rand : (i == 4 ? depth_pert : base_d);
        tp.w = distance_to_origin(K, px, py, d, tp);
        const PlaneM m = plane_row(K, tp);
        float acc = 0.0f;
        for (uint32_t mk = wmask; mk; mk &= mk - 1) {
            const int v = __ffs(mk) - 1;
            float c = ncc_old_x<U, SA>(K, s_vk[v], px, py, m, rp, si);
            n_old++;
            if (use_geom) { c = __fmaf_rn(K.geom_factor, geom_cost(K, s_vk[v], v, px, py, tp), c); n_geom++; }
            acc += (float)vw_get(w, v) * c;
        }
        const float tc = acc / wnorm;
        const float db = depth_from_plane(K, tp, px, py);
        if (db >= dmin && db <= dmax && tc < cost_now) { depth_now = db; plane_now = tp; cost_now = tc; }
    }
    if (K.state == APDE_REFINE_INIT) {
        // costs[center] was overwritten with the recomputed current cost before this test (quirk 5)
        if ((double)cost_now < (double)cost_written - 0.1) { K.costs[center] = cost_now; K.planes[center] = plane_now; }
        else K.costs[center] = cost_written;
    } else {
        K.costs[center] = cost_now;
        K.planes[center] = plane_now;
    }
    count_evals(K, n_old, 0, n_geom);
}
__global__ void __launch_bounds__(128) k_prop_strong_v1(const __grid_constant__ PassK K, int iter, int color, int tiles_x,
                                                        int ylimit) {
    if (K.tex_unorm > 0.0f) k_prop_strong_body<true, false>(K, iter, color, tiles_x, ylimit);
    else k_prop_strong_body<false, false>(K, iter, color, tiles_x, ylimit);
}
__global__ void __launch_bounds__(128) k_prop_strong_sa(const __grid_constant__ PassK K, int iter, int color, int tiles_x,
                                                        int ylimit) {
    if (K.tex_unorm > 0.0f) k_prop_strong_body<true, true>(K, iter, color, tiles_x, ylimit);
    else k_prop_strong_body<false, true>(K, iter, color, tiles_x, ylimit);
}

// The same half-sweep with the phase-3 evaluations (current hypothesis, then the five refinement hypotheses) WARP-COMPACTED.
//
// In the kernel above every lane walks its OWN selected views in phase 3: at one instant the 32 lanes of a warp sample up
// to N different source views (different layers of the texture array -> no cache locality) and the warp runs as long as
// its lane with the most selected views.  Here the (pixel, selected view) pairs of a warp are listed VIEW-MAJOR in shared
// memory and dealt out 32 at a time: a warp-wide texture fetch then covers neighbouring pixels of ONE view, all lanes carry
// work, and every pair evaluates the five refinement hypotheses back to back.  The owner lane parks its reference patch and
// its five hypotheses in shared memory (the phase-1 cost columns are dead by then), any lane may evaluate them, and the owner
// gathers the costs afterwards and forms the weighted sums in ascending view order -- the arithmetic per cost and per sum
// is unchanged, so the results are bit-identical to the kernel above (tests/test_gpu_parity.py runs both).
//
// per-thread shared-memory column (slot * blockDim + thread):  phase 1-2: [8N] candidate costs, [N] view priors
//                                                              phase 3  : [38] ref patch, [20] 5 planes, [5N] costs
//   with 8-bit texels (U) the reference patch is 36 integers 0..255: parked as 9 packed words, so a column is 9N floats again
//   (N = 10: 90 instead of 108: 28 KB more L1 for the texture path at 3 resident blocks per SM)
constexpr int kStrongRefSlot = 0;
__host__ __device__ constexpr int strong_ref_words(bool U) { return (U ? kPatch / 4 : kPatch) + 2; }
__host__ __device__ inline int strong_column_floats(int N, bool U) { return max(9 * N, strong_ref_words(U) + 20 + 5 * N); }

template <bool U>
__device__ __forceinline__ void k_prop_strong_compact_body(const PassK &K, int iter, int color, int tiles_x, int ylimit) {
    extern __shared__ float smem[];
    const ViewK *s_vk = stage_views(K, smem);
    const int W = K.W, N = K.N;
    const int stride = blockDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *col0 = smem + views_smem_floats(N);                 // columns of all threads
    float *sc = col0 + threadIdx.x;                            // this thread's column
    float *sp = sc + 8 * N * stride;
    constexpr int kStrongPlaneSlot = strong_ref_words(U), kStrongCostSlot = strong_ref_words(U) + 20;
    unsigned short *items = reinterpret_cast<unsigned short *>(col0 + (size_t)strong_column_floats(N, U) * stride) + (size_t)warp * 32 * N;

    int px = 0, py = 0;
    bool active = half_pixel_or_list(K, color, tiles_x, ylimit, px, py);
    const int center = py * W + px;
    if (active && K.weak[center] == APDE_WEAK) active = false;

    unsigned n_old = 0, n_geom = 0;
    uint32_t wmask = 0;
    uint4 w = make_uint4(0, 0, 0, 0);
    float wnorm = 1.0f, cost_now = 0.0f, cost_written = 0.0f, depth_now = 0.0f, fc_min = 0.0f;
    float4 plane_now = make_float4(0.f, 0.f, 0.f, 0.f), plane_c = plane_now;
    int cand_center = -1;  // pixel whose plane won the candidate vote (-1: the winner was an invalid slot, quirk 2)
    const bool use_geom = K.geom && K.impetus;
    const float dmin = K.depth_min, dmax = K.depth_max;
    Rng rng(K.seed, K.stream, (uint32_t)center, SITE_STRONG + iter);

    // ---- phase 1 + 2 (owner lane): 8 candidates x N views, joint view selection, candidate vote
    if (active) {
        RefPatch rp;
        load_ref_patch<U>(K, px, py, rp);
        int pos[8];
        const unsigned flags = checkerboard_candidates(K.costs, W, K.H, px, py, pos);
#pragma unroll 1
        for (int h = 0; h < 8; ++h) {
            if ((flags >> h) & 1u) {
                const PlaneM m = plane_row(K, K.planes[pos[h]]);
#pragma unroll 1
                for (int v = 0; v < N; ++v) sc[(h * N + v) * stride] = ncc_old<U>(K, K.v[v], px, py, m, rp);
                n_old += N;
            } else {
                // quirk 2: "float cost_array[8][32] = {2.0f}" leaves every entry 0 except [0][0]
                for (int v = 0; v < N; ++v) sc[(h * N + v) * stride] = (h == 0 && v == 0) ? 2.0f : 0.0f;
            }
        }
        uint32_t nbr_sel[4];
        const unsigned nbr_valid = ((flags >> 0) & 1u) | (((flags >> 2) & 1u) << 1) | (((flags >> 4) & 1u) << 2) |
                                   (((flags >> 6) & 1u) << 3);
        nbr_sel[0] = (nbr_valid & 1u) ? K.sel[center - W] : 0u;
        nbr_sel[1] = (nbr_valid & 2u) ? K.sel[center + W] : 0u;
        nbr_sel[2] = (nbr_valid & 4u) ? K.sel[center - 1] : 0u;
        nbr_sel[3] = (nbr_valid & 8u) ? K.sel[center + 1] : 0u;

        w = select_views<4>(K, sc, sp, stride, nbr_sel, nbr_valid, iter, rng, &wmask, &wnorm);
        K.vw[center] = w;

        float final_costs[8];
#pragma unroll
        for (int h = 0; h < 8; ++h) {
            float acc = 0.0f;
            for (uint32_t mk = wmask; mk; mk &= mk - 1) {
                const int v = __ffs(mk) - 1;
                acc += (float)vw_get(w, v) * sc[(h * N + v) * stride];
            }
            final_costs[h] = acc / wnorm;
        }
        int min_idx = 0;
        {
            float mc = final_costs[0];
#pragma unroll
            for (int h = 1; h < 8; ++h) if (final_costs[h] <= mc) { mc = final_costs[h]; min_idx = h; }  // ties -> last (quirk 3)
        }
        fc_min = final_costs[0];
#pragma unroll
        for (int h = 1; h < 8; ++h) if (h == min_idx) fc_min = final_costs[h];
#pragma unroll
        for (int h = 0; h < 8; ++h) if (h == min_idx && ((flags >> h) & 1u)) cand_center = pos[h];

        plane_c = K.planes[center];
        // park the reference patch and the current plane in this thread's column (the phase-1 costs are dead now)
        if (U) {
#pragma unroll
            for (int k = 0; k < kPatch / 4; ++k) {
                const unsigned word = (unsigned)rp.r[4 * k] | ((unsigned)rp.r[4 * k + 1] << 8) | ((unsigned)rp.r[4 * k + 2] << 16) |
                                      ((unsigned)rp.r[4 * k + 3] << 24);
                sc[(kStrongRefSlot + k) * stride] = __uint_as_float(word);
            }
        } else {
#pragma unroll
            for (int k = 0; k < kPatch; ++k) sc[(kStrongRefSlot + k) * stride] = rp.r[k];
        }
        sc[(kStrongPlaneSlot - 2) * stride] = rp.mean;
        sc[(kStrongPlaneSlot - 1) * stride] = rp.var;
        sc[(kStrongPlaneSlot + 0) * stride] = plane_c.x;
        sc[(kStrongPlaneSlot + 1) * stride] = plane_c.y;
        sc[(kStrongPlaneSlot + 2) * stride] = plane_c.z;
        sc[(kStrongPlaneSlot + 3) * stride] = plane_c.w;
    }
    __syncwarp();

    // ---- (pixel, selected view) pairs of this warp, view-major
    int total = 0;
    for (int v = 0; v < N; ++v) {
        const unsigned m = __ballot_sync(0xffffffffu, active && ((wmask >> v) & 1u));
        if (active && ((wmask >> v) & 1u)) items[total + __popc(m & ((1u << lane) - 1u))] = (unsigned short)(lane | (v << 8));
        total += __popc(m);
    }
    __syncwarp();
    const int pxy = (px << 16) | py;
    // Two rounds over ONE instance of the evaluation code (the kernel's instruction footprint, not its arithmetic, is what a
    // second inlined copy would cost): round 0 = the current hypothesis, round 1 = the five refinement hypotheses.  In each
    // round every pair evaluates the nh planes parked in plane slots 0 .. nh-1 of its owner's column and leaves the costs in
    // cost slot [i][v]; the owner lane then consumes them.
#pragma unroll 1
    for (int round = 0; round < 2; ++round) {
        const int nh = round ? 5 : 1;
#pragma unroll 1
        for (int base = 0; base < total; base += 32) {
            const bool has = base + lane < total;
            const unsigned short it = has ? items[base + lane] : (unsigned short)0;
            const int src = it & 31, v = it >> 8;
            const int sxy = __shfl_sync(0xffffffffu, pxy, src);
            if (has) {
                const int spx = sxy >> 16, spy = sxy & 0xffff;
                float *scol = col0 + (warp * 32 + src);
                RefPatch rp;
                if (U) {
#pragma unroll
                    for (int k = 0; k < kPatch / 4; ++k) {
                        const unsigned word = __float_as_uint(scol[(kStrongRefSlot + k) * stride]);
                        rp.r[4 * k] = (float)(word & 255u);
                        rp.r[4 * k + 1] = (float)((word >> 8) & 255u);
                        rp.r[4 * k + 2] = (float)((word >> 16) & 255u);
                        rp.r[4 * k + 3] = (float)(word >> 24);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < kPatch; ++k) rp.r[k] = scol[(kStrongRefSlot + k) * stride];
                }
                rp.mean = scol[(kStrongPlaneSlot - 2) * stride];
                rp.var = scol[(kStrongPlaneSlot - 1) * stride];
                const ViewK &vk = s_vk[v];
#pragma unroll 1
                for (int i = 0; i < nh; ++i) {
                    const float4 tp = make_float4(scol[(kStrongPlaneSlot + 4 * i + 0) * stride], scol[(kStrongPlaneSlot + 4 * i + 1) * stride],
                                                  scol[(kStrongPlaneSlot + 4 * i + 2) * stride], scol[(kStrongPlaneSlot + 4 * i + 3) * stride]);
                    const PlaneM m = plane_row(K, tp);
                    float c = ncc_old<U>(K, vk, spx, spy, m, rp);
                    n_old++;
                    if (use_geom) { c = __fmaf_rn(K.geom_factor, geom_cost(K, vk, v, spx, spy, tp), c); n_geom++; }
                    scol[(kStrongCostSlot + i * N + v) * stride] = c;
                }
            }
        }
        __syncwarp();
        if (active && round == 0) {
            // ---- the current hypothesis on the selected views
            float acc = 0.0f;
            for (uint32_t mk = wmask; mk; mk &= mk - 1) {
                const int v = __ffs(mk) - 1;
                acc += (float)vw_get(w, v) * sc[(kStrongCostSlot + v) * stride];
            }
            cost_now = acc / wnorm;
            cost_written = cost_now;
            plane_now = plane_c;
            depth_now = depth_from_plane(K, plane_c, px, py);
            if (cand_center >= 0) {
                const float4 cand = K.planes[cand_center];
                const float db = depth_from_plane(K, cand, px, py);
                if (db >= dmin && db <= dmax && fc_min < cost_now) {
                    depth_now = db; plane_now = cand; cost_now = fc_min;
                    K.sel[center] = wmask;
                }
            }
            // PlaneHypothesisRefinementStrong: all five candidates are built from the state BEFORE the loop (APD.cu:968-980)
            const float depth_rand = rng.uniform() * (dmax - dmin) + dmin;
            float4 cand_n[2];
            cand_n[0] = random_normal(K, px, py, rng, depth_now);
            const float lo = (1.0f - 0.02f) * depth_now, hi = (1.0f + 0.02f) * depth_now;
            const float depth_pert = rng.uniform() * (hi - lo) + lo;  // the do-while can never repeat (quirk 6)
            cand_n[1] = perturbed_normal(K, px, py, plane_now, rng, (float)(0.02 * 3.14159265358979323846));
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                float4 tp = (i == 0 || i == 4) ? plane_now : (i == 3 ? cand_n[1] : cand_n[0]);
                const float d = (i == 0 || i == 2) ? depth_rand : (i == 4 ? depth_pert : depth_now);
                tp.w = distance_to_origin(K, px, py, d, tp);
                sc[(kStrongPlaneSlot + 4 * i + 0) * stride] = tp.x;
                sc[(kStrongPlaneSlot + 4 * i + 1) * stride] = tp.y;
                sc[(kStrongPlaneSlot + 4 * i + 2) * stride] = tp.z;
                sc[(kStrongPlaneSlot + 4 * i + 3) * stride] = tp.w;
            }
        } else if (active) {
            // ---- the five refinement hypotheses
#pragma unroll 1
            for (int i = 0; i < 5; ++i) {
                const float4 tp = make_float4(sc[(kStrongPlaneSlot + 4 * i + 0) * stride], sc[(kStrongPlaneSlot + 4 * i + 1) * stride],
                                              sc[(kStrongPlaneSlot + 4 * i + 2) * stride], sc[(kStrongPlaneSlot + 4 * i + 3) * stride]);
                float acc = 0.0f;
                for (uint32_t mk = wmask; mk; mk &= mk - 1) {
                    const int v = __ffs(mk) - 1;
                    acc += (float)vw_get(w, v) * sc[(kStrongCostSlot + i * N + v) * stride];
                }
                const float tc = acc / wnorm;
                const float db = depth_from_plane(K, tp, px, py);
                if (db >= dmin && db <= dmax && tc < cost_now) { depth_now = db; plane_now = tp; cost_now = tc; }
            }
            if (K.state == APDE_REFINE_INIT) {
                // costs[center] was overwritten with the recomputed current cost before this test (quirk 5)
                if ((double)cost_now < (double)cost_written - 0.1) { K.costs[center] = cost_now; K.planes[center] = plane_now; }
                else K.costs[center] = cost_written;
            } else {
                K.costs[center] = cost_now;
                K.planes[center] = plane_now;
            }
        }
        __syncwarp();
    }
    count_evals(K, n_old, 0, n_geom);
}
// 3 resident blocks per SM.  Measured on B200 (r02, 1920x1080): a 4th block (128 registers, 96 B of stack, 200 KB of shared
// memory per SM) leaves the texture path ~56 KB of L1 and costs 33 % (strong propagation 1852 -> 2456 ms per step)
__global__ void __launch_bounds__(128, 3) k_prop_strong(const __grid_constant__ PassK K, int iter, int color, int tiles_x,
                                                     int ylimit) {
    if (K.tex_unorm > 0.0f) k_prop_strong_compact_body<true>(K, iter, color, tiles_x, ylimit);
    else k_prop_strong_compact_body<false>(K, iter, color, tiles_x, ylimit);
}


// -------------------------------------------------------------------------------------------- K9
// GetDepthandNormal, APD.cu:1694-1709
__global__ void __launch_bounds__(256) k_depth_normal(const __grid_constant__ PassK K) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= K.W * K.H) return;
    const int px = idx % K.W, py = idx / K.W;
    float4 pl = K.planes[idx];
    pl.w = depth_from_plane(K, pl, px, py);
    K.planes[idx] = normal_to_world(K, pl);
}

// -------------------------------------------------------------------------------------------- K10 median
// Black/RedPixelFilterStrong -> CheckerboardFilterStrong, APD.cu:1711-1855
__global__ void __launch_bounds__(128) k_median(const __grid_constant__ PassK K, int color, int tiles_x, int ylimit) {
    int px, py;
    if (!half_pixel(K, color, tiles_x, ylimit, px, py)) return;
    const int width = K.W, height = K.H;
    const int center = py * width + px;
    if (K.weak[center] == APDE_WEAK) return;
    if (K.costs[center] < 0.001f) return;
    float f[21];
    int n = 0;
    f[n++] = K.planes[center].w;
    const int left = center - 1, leftleft = center - 3, up = center - width, upup = center - 3 * width;
    const int down = center + width, downdown = center + 3 * width, right = center + 1, rightright = center + 3;
#define APDE_TAP(cond, idx) if ((cond) && K.weak[(idx)] == APDE_STRONG) f[n++] = K.planes[(idx)].w;
    APDE_TAP(py > 0, up)
    APDE_TAP(py > 2, upup)
    APDE_TAP(py > 4, upup - width * 2)
    APDE_TAP(py < height - 1, down)
    APDE_TAP(py < height - 3, downdown)
    APDE_TAP(py < height - 5, downdown + width * 2)
    APDE_TAP(px > 0, left)
    APDE_TAP(px > 2, leftleft)
    APDE_TAP(px > 4, leftleft - 2)
    APDE_TAP(px < width - 1, right)
    APDE_TAP(px < width - 3, rightright)
    APDE_TAP(px < width - 5, rightright + 2)
    APDE_TAP(py > 0 && px < width - 2, up + 2)
    APDE_TAP(py < height - 1 && px < width - 2, down + 2)
    APDE_TAP(py > 0 && px > 1, up - 2)
    APDE_TAP(py < height - 1 && px > 1, down - 2)
    APDE_TAP(px > 0 && py > 2, left - width * 2)
    APDE_TAP(px < width - 1 && py > 2, right - width * 2)
    APDE_TAP(px > 0 && py < height - 2, left + width * 2)
    APDE_TAP(px < width - 1 && py < height - 2, right + width * 2)
#undef APDE_TAP
    for (int i = 1; i < n; i++) {
        const float tmp = f[i];
        int j;
        for (j = i; j >= 1 && tmp < f[j - 1]; j--) f[j] = f[j - 1];
        f[j] = tmp;
    }
    const int mi = n / 2;
    K.planes[center].w = (n % 2 == 0) ? (f[mi - 1] + f[mi]) / 2 : f[mi];
}

// -------------------------------------------------------------------------------------------- K11 DepthToWeak
// APD.cu:2103-2250: 61-sample disparity sweep of the view-weighted cost -> WEAK / STRONG / UNKNOWN
template <bool U, bool SA>
__device__ __forceinline__ void k_depth_to_weak_body(const PassK &K, int tiles_x, float *curve) {
    extern __shared__ float smem[];
    const ViewK *s_vk = stage_views(K, smem);
    int px, py;
    if (!full_pixel(K, tiles_x, px, py)) return;
    const int W = K.W, H = K.H;
    const int center = py * W + px;
    const int min_margin = 6;
    if (px < min_margin || py < min_margin || px >= W - min_margin || py >= H - min_margin) { K.weak[center] = APDE_UNKNOWN; return; }
    float4 opl = normal_to_refcam(K, K.planes[center]);
    const float origin_depth = opl.w;
    if (origin_depth == 0.0f) { K.weak[center] = APDE_UNKNOWN; return; }
    const uint32_t sel = K.sel[center];
    const uint4 w = K.vw[center];
    float base_line = 0.0f, weight_normal = 0.0f;
    int valid_src = 0;
    for (uint32_t mk = sel; mk; mk &= mk - 1) {
        const int v = __ffs(mk) - 1;
        weight_normal = __fadd_rn(weight_normal, (float)vw_get(w, v));
        base_line = __fadd_rn(base_line, view_baseline(K, s_vk[v]));
        valid_src++;
    }
    if (valid_src == 0) { K.weak[center] = APDE_UNKNOWN; return; }
    const SweepDepths sd = sweep_depths(K, base_line, valid_src, origin_depth);

    RefPatch rp;
    load_ref_patch<U>(K, px, py, rp);
    typename SaTypes<SA>::Info si;
    load_sa<SA>(K, px, py, rp, si);
    unsigned n_old = 0, n_geom = 0;
    const int radius = 30, n = 61;
    const float rwn = rcp_approx(weight_normal);
    float pc[61];
#pragma unroll 1
    for (int pd = -radius; pd <= radius; ++pd) {
        const float p_depth = sd.depth(pd);
        if (p_depth < K.depth_min || p_depth > K.depth_max) { pc[pd + radius] = 2.0f; continue; }
        float4 tp = opl;
        tp.w = distance_to_origin(K, px, py, p_depth, tp);
        const PlaneM m = plane_row(K, tp);
        float p_cost = 0.0f;
        for (uint32_t mk = sel; mk; mk &= mk - 1) {
            const int v = __ffs(mk) - 1;
            float tc = ncc_old_x<U, SA>(K, s_vk[v], px, py, m, rp, si);
            n_old++;
            if (K.geom) { tc = __fmaf_rn(K.geom_factor, geom_cost(K, s_vk[v], v, px, py, tp), tc); n_geom++; }
            p_cost = __fmaf_rn((float)vw_get(w, v), tc, p_cost);  // APD.cu:2179-2181 as built
        }
        p_cost = __fmul_rn(p_cost, rwn);
        pc[pd + radius] = (2.0f > p_cost) ? p_cost : 2.0f;  // OpenCV MIN(2.0f, p_cost): NaN -> 2
    }
    count_evals(K, n_old, 0, n_geom);
    if (curve) for (int i = 0; i < n; ++i) curve[(size_t)center * n + i] = pc[i];

    unsigned long long peaks = 0ull;
    int peak_count = 0, min_peak = 0;
    float min_cost = 2.0f;
    for (int i = 2; i < n - 2; ++i) {
        if (pc[i - 1] > pc[i] && pc[i + 1] > pc[i]) {
            peaks |= 1ull << i;
            peak_count++;
            if (pc[i] < min_cost) { min_peak = i; min_cost = pc[i]; }
        }
    }
    if (abs(min_peak - radius) > K.weak_peak_radius || pc[min_peak] > 0.5f) { K.weak[center] = APDE_WEAK; return; }
    if (peak_count == 1) { K.weak[center] = (pc[min_peak] <= 0.15f) ? APDE_STRONG : APDE_WEAK; return; }
    float var = 0.0f;
    for (int i = 2; i < n - 2; ++i) {
        if (((peaks >> i) & 1ull) && i != min_peak) { const float d = pc[i] - min_cost; var += d * d; }
    }
    var = sqrtf(var);
    var /= (peak_count - 1);
    K.weak[center] = (var > 0.2f) ? APDE_STRONG : APDE_WEAK;
}
__global__ void __launch_bounds__(128) k_depth_to_weak(const __grid_constant__ PassK K, int tiles_x, float *curve) {
    if (K.tex_unorm > 0.0f) k_depth_to_weak_body<true, false>(K, tiles_x, curve);
    else k_depth_to_weak_body<false, false>(K, tiles_x, curve);
}
__global__ void __launch_bounds__(128) k_depth_to_weak_sa(const __grid_constant__ PassK K, int tiles_x, float *curve) {
    if (K.tex_unorm > 0.0f) k_depth_to_weak_body<true, true>(K, tiles_x, curve);
    else k_depth_to_weak_body<false, true>(K, tiles_x, curve);
}


// -------------------------------------------------------------------------------------------- K12 confidence
// ConfidenceCompute, APD.cu:2282-2344
__global__ void __launch_bounds__(128) k_confidence(const __grid_constant__ PassK K, int tiles_x) {
    extern __shared__ float smem[];
    const ViewK *s_vk = stage_views(K, smem);
    int px, py;
    if (!full_pixel(K, tiles_x, px, py)) return;
    const int center = py * K.W + px;
    K.conf[center] = 0;
    const uint32_t sel = K.sel[center];
    const float ref_depth = K.planes[center].w;
    if (ref_depth <= 0.0f) { K.weak[center] = APDE_UNKNOWN; return; }
    const float fxp = (float)px, fyp = (float)py;
    const float3 Pw = point_to_world(K.Kr, K.R, K.c, fxp, fyp, ref_depth);
    int nc = 1;
    for (uint32_t mk = sel; mk; mk &= mk - 1) {
        const int v = __ffs(mk) - 1;
        const ViewK &vk = s_vk[v];
        const Proj s = project_to_camera(vk.K, vk.R, vk.t, Pw);
        const float sx = __fmul_rn(s.nx, s.rd), sy = __fmul_rn(s.ny, s.rd);
        const float sd = source_depth(K, v, sx, sy);
        if (sd <= 0.0f) continue;
        nc += 1;
        const float3 Ps = point_to_world(vk.K, vk.R, vk.c, sx, sy, sd);
        const Proj b = project_to_camera(K.Kr, K.R, K.t, Ps);
        const float dc = __fmaf_rn(-b.nx, b.rd, fxp), dr = __fmaf_rn(-b.ny, b.rd, fyp);
        if (sqrt_approx(__fmaf_rn(dc, dc, __fmul_rn(dr, dr))) <= 2.0f) nc += 2;
        if (__fmul_rn(fabsf(__fsub_rn(ref_depth, b.d)), rcp_approx(ref_depth)) <= 0.02f) nc += 2;
    }
    K.conf[center] = (uint8_t)min(nc, 255);
}

// -------------------------------------------------------------------------------------------- K13 local refine
// LocalRefine, APD.cu:2346-2432
template <bool U, bool SA>
__device__ __forceinline__ void k_local_refine_body(const PassK &K, int tiles_x) {
    extern __shared__ float smem[];
    const ViewK *s_vk = stage_views(K, smem);
    int px, py;
    if (!full_pixel(K, tiles_x, px, py)) return;
    const int center = py * K.W + px;
    float4 opl = normal_to_refcam(K, K.planes[center]);
    const float origin_depth = opl.w;
    if (origin_depth == 0.0f) return;
    const uint32_t sel = K.sel[center];
    if (sel == 0u) return;
    const uint4 w = K.vw[center];
    RefPatch rp;
    load_ref_patch<U>(K, px, py, rp);
    typename SaTypes<SA>::Info si;
    load_sa<SA>(K, px, py, rp, si);
    unsigned n_old = 0, n_geom = 0;
    float cost_now = 0.0f, base_line = 0.0f, weight_normal = 0.0f;
    int valid_src = 0;
    {
        float4 tp = opl;
        tp.w = distance_to_origin(K, px, py, origin_depth, tp);
        const PlaneM m = plane_row(K, tp);
        for (uint32_t mk = sel; mk; mk &= mk - 1) {
            const int v = __ffs(mk) - 1;
            float tc = ncc_old_x<U, SA>(K, s_vk[v], px, py, m, rp, si);
            n_old++;
            if (K.geom) { tc = __fmaf_rn(K.geom_factor, geom_cost(K, s_vk[v], v, px, py, tp), tc); n_geom++; }
            const float wv = (float)vw_get(w, v);
            cost_now = __fmaf_rn(wv, tc, cost_now);  // APD.cu:2380-2383 as built
            weight_normal = __fadd_rn(weight_normal, wv);
            base_line = __fadd_rn(base_line, view_baseline(K, s_vk[v]));
            valid_src++;
        }
    }
    if (weight_normal == 0.0f) { count_evals(K, n_old, 0, n_geom); return; }
    const float rwn = rcp_approx(weight_normal);
    const SweepDepths sd = sweep_depths(K, base_line, valid_src, origin_depth);
    float min_cost = 2.0f, best_depth = origin_depth;
#pragma unroll 1
    for (int pd = -5; pd <= 5; ++pd) {
        const float p_depth = sd.depth(pd);
        if (p_depth < K.depth_min || p_depth > K.depth_max) continue;
        float4 tp = opl;
        tp.w = distance_to_origin(K, px, py, p_depth, tp);
        const PlaneM m = plane_row(K, tp);
        float tc = 0.0f;
        for (uint32_t mk = sel; mk; mk &= mk - 1) {
            const int v = __ffs(mk) - 1;
            const float wv = (float)vw_get(w, v);
            tc = __fmaf_rn(wv, ncc_old_x<U, SA>(K, s_vk[v], px, py, m, rp, si), tc);  // APD.cu:2417-2419 as built
            n_old++;
            if (K.geom) { tc = __fmaf_rn(wv, __fmul_rn(K.geom_factor, geom_cost(K, s_vk[v], v, px, py, tp)), tc); n_geom++; }
        }
        tc = __fmul_rn(tc, rwn);
        if (tc < min_cost) { min_cost = tc; best_depth = p_depth; }
    }
    // "cost_now /= weight_normal; ... cost_now - min_cost > 0.1" is one FFMA in the reference build, compared in double
    if ((double)__fmaf_rn(rwn, cost_now, -min_cost) > 0.1) K.planes[center].w = best_depth;
    count_evals(K, n_old, 0, n_geom);
}
__global__ void __launch_bounds__(128) k_local_refine(const __grid_constant__ PassK K, int tiles_x) {
    if (K.tex_unorm > 0.0f) k_local_refine_body<true, false>(K, tiles_x);
    else k_local_refine_body<false, false>(K, tiles_x);
}
__global__ void __launch_bounds__(128) k_local_refine_sa(const __grid_constant__ PassK K, int tiles_x) {
    if (K.tex_unorm > 0.0f) k_local_refine_body<true, true>(K, tiles_x);
    else k_local_refine_body<false, true>(K, tiles_x);
}


// -------------------------------------------------------------------------------------------- parity hook
template <bool U, bool SA>
__device__ __forceinline__ void k_eval_costs_body(const PassK &K, int n, const int *__restrict__ tuples,
                                                    const float4 *__restrict__ planes, int mode, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int px = tuples[3 * i], py = tuples[3 * i + 1], v = tuples[3 * i + 2] - 1;
    const float4 pl = planes[i];
    float c;
    if (mode == 2) {
        c = geom_cost(K, K.v[v], v, px, py, pl);
    } else {
        RefPatch rp;
        load_ref_patch<U>(K, px, py, rp);
        typename SaTypes<SA>::Info si;
        load_sa<SA>(K, px, py, rp, si);
        const PlaneM m = plane_row(K, pl);
        if (mode == 0) {
            c = ncc_old_x<U, SA>(K, K.v[v], px, py, m, rp, si);
        } else {
            typename SaTypes<SA>::Anchors ar;
            load_anchor_ref_x<U, SA>(K, K.anchors + (size_t)(py * K.W + px) * APDE_ANCHOR_NUM, ar, si);
            c = ncc_new_x<U, SA>(K, K.v[v], v, px, py, m, rp, ar, si);
        }
    }
    out[i] = c;
}
__global__ void __launch_bounds__(128) k_eval_costs(const __grid_constant__ PassK K, int n, const int *__restrict__ tuples,
                                                    const float4 *__restrict__ planes, int mode, float *__restrict__ out) {
    if (K.tex_unorm > 0.0f) k_eval_costs_body<true, false>(K, n, tuples, planes, mode, out);
    else k_eval_costs_body<false, false>(K, n, tuples, planes, mode, out);
}
__global__ void __launch_bounds__(128) k_eval_costs_sa(const __grid_constant__ PassK K, int n, const int *__restrict__ tuples,
                                                       const float4 *__restrict__ planes, int mode, float *__restrict__ out) {
    if (K.tex_unorm > 0.0f) k_eval_costs_body<true, true>(K, n, tuples, planes, mode, out);
    else k_eval_costs_body<false, true>(K, n, tuples, planes, mode, out);
}


// -------------------------------------------------------------------------------------------- launchers
static size_t prop_smem_bytes(int N, int threads) { return sizeof(float) * ((size_t)views_smem_floats(N) + (size_t)9 * N * threads); }
int prop_block_threads(int N) {
    if (prop_smem_bytes(N, 128) <= 56 * 1024) return 128;
    if (prop_smem_bytes(N, 64) <= 64 * 1024) return 64;
    return 32;
}

cudaError_t launch_stage(const PassK &K, int stage, int iter, int color, cudaStream_t st, float *curve) {
    const int W = K.W, H = K.H, N = K.N;
    const bool sa = K.sa != nullptr;  // segment labels: the <SA = true> twins of the thread-per-pixel kernels
    const int ylimit = min(H, half_rows_limit(H));
    const int tiles8x = (W + 7) / 8;
    const size_t vsm = sizeof(float) * views_smem_floats(N);
    switch (stage) {
        case APDE_STAGE_INIT: {
            const int tiles = tiles8x * ((H + 3) / 4);
            if (sa) {  // a label map exists only in use_APD passes (APD.cpp:614, 641)
                k_init_sa<<<(tiles + 3) / 4, 128, 0, st>>>(K, tiles8x, 0);
                k_init_sa<<<(tiles + 3) / 4, 128, 0, st>>>(K, tiles8x, 1);
            } else if (K.use_apd) {
                k_init<<<(tiles + 3) / 4, 128, 0, st>>>(K, tiles8x, 0);
                k_init<<<(tiles + 3) / 4, 128, 0, st>>>(K, tiles8x, 1);
            } else {
                k_init<<<(tiles + 3) / 4, 128, 0, st>>>(K, tiles8x, -1);
            }
            break;
        }
        case APDE_STAGE_PROP_STRONG: {
            const char *v1env = getenv("APDE_STRONG_V1");  // read per launch: tests toggle it inside one process
            const bool v1 = v1env && v1env[0] == '1';
            const int tiles = half_tiles(K, ylimit);  // 32 same-colour pixels per tile == worst-case list length / 32
            const int tiles8x = half_tiles_x(K);
            if (v1 || sa) {  // per-lane refinement loop (kept for A/B measurements and as the parity twin of the compacted kernel)
                const int threads = prop_block_threads(N);
                const size_t smem = prop_smem_bytes(N, threads);
                static SmemOptIn configured;
                int dev = 0;
                if (configured.needed(smem, &dev)) {
                    cudaError_t e = cudaFuncSetAttribute(k_prop_strong_v1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_prop_strong_sa, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    if (e != cudaSuccess) return e;
                    configured.done(smem, dev);
                }
                const int wpb = threads / 32;
                if (sa) k_prop_strong_sa<<<(tiles + wpb - 1) / wpb, threads, smem, st>>>(K, iter, color, tiles8x, ylimit);
                else k_prop_strong_v1<<<(tiles + wpb - 1) / wpb, threads, smem, st>>>(K, iter, color, tiles8x, ylimit);
                break;
            }
            // column floats + the warp's pair list (N shorts per thread)
            const bool u8 = K.tex_unorm > 0.0f;
            auto bytes = [&](int threads) { return sizeof(float) * ((size_t)views_smem_floats(N) + (size_t)(strong_column_floats(N, u8) + (N + 1) / 2) * threads); };
            const int threads = bytes(128) <= 72 * 1024 ? 128 : (bytes(64) <= 100 * 1024 ? 64 : 32);
            const size_t smem = bytes(threads);
            static SmemOptIn configured;
            int dev = 0;
            if (configured.needed(smem, &dev)) {
                cudaError_t e = cudaFuncSetAttribute(k_prop_strong, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return e;
                configured.done(smem, dev);
                // experiment knob: shared-memory carve-out in per cent (less shared memory = fewer resident CTAs, more L1/TEX cache)
                if (const char *co = getenv("APDE_STRONG_CARVEOUT"))
                    cudaFuncSetAttribute(k_prop_strong, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(co));
            }
            const int wpb = threads / 32;
            k_prop_strong<<<(tiles + wpb - 1) / wpb, threads, smem, st>>>(K, iter, color, tiles8x, ylimit);
            break;
        }
        case APDE_STAGE_DEPTH_NORMAL:
            k_depth_normal<<<(W * H + 255) / 256, 256, 0, st>>>(K);
            break;
        case APDE_STAGE_MEDIAN: {
            const int tiles = half_tiles(K, ylimit);
            k_median<<<(tiles + 3) / 4, 128, 0, st>>>(K, color, half_tiles_x(K), ylimit);
            break;
        }
        case APDE_STAGE_DEPTH_TO_WEAK: {
            const int tiles = tiles8x * ((H + 3) / 4);
            if (sa) k_depth_to_weak_sa<<<(tiles + 3) / 4, 128, vsm, st>>>(K, tiles8x, curve);
            else k_depth_to_weak<<<(tiles + 3) / 4, 128, vsm, st>>>(K, tiles8x, curve);
            break;
        }
        case APDE_STAGE_CONFIDENCE: {
            const int tiles = tiles8x * ((H + 3) / 4);
            k_confidence<<<(tiles + 3) / 4, 128, vsm, st>>>(K, tiles8x);
            break;
        }
        case APDE_STAGE_LOCAL_REFINE: {
            const int tiles = tiles8x * ((H + 3) / 4);
            if (sa) k_local_refine_sa<<<(tiles + 3) / 4, 128, vsm, st>>>(K, tiles8x);
            else k_local_refine<<<(tiles + 3) / 4, 128, vsm, st>>>(K, tiles8x);
            break;
        }
        default:
            return launch_stage_apd(K, stage, iter, color, st);
    }
    return cudaGetLastError();
}

cudaError_t launch_eval_costs(const PassK &K, int n, const int *tuples, const float4 *planes, int mode, float *out,
                              cudaStream_t st) {
    if (K.sa) k_eval_costs_sa<<<(n + 127) / 128, 128, 0, st>>>(K, n, tuples, planes, mode, out);
    else k_eval_costs<<<(n + 127) / 128, 128, 0, st>>>(K, n, tuples, planes, mode, out);
    return cudaGetLastError();
}

}  // namespace apde
