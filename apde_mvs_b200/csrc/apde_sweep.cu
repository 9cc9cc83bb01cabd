// apde_sweep.cu -- DepthToWeak (APD.cu:2103-2250) and LocalRefine (APD.cu:2346-2432) as a balanced column workload.
//
// Both kernels of the reference sweep the depth of a pixel along the viewing ray (disparity steps of one pixel, +-30 and
// +-5) and sum, over the pixel's SELECTED views in ascending order, view-weighted NCC (+ geometric) costs.  One thread
// per pixel makes every warp wait for its pixel with the most selected views and mixes source views inside a warp.  Here
//   1. k_sweep_prepare   marks (pixel, view) columns;                         a prefix sum numbers them VIEW-MAJOR;
//   2. k_sweep_columns   one thread = one column: 61 sweep evaluations + the evaluation at the current depth;
//                        a warp holds 32 neighbouring pixels of the SAME source view, every lane does the same work;
//   3. k_sweep_classify  (DepthToWeak) / k_sweep_refine (LocalRefine): one thread per pixel adds its columns in the
//                        reference order and applies the reference's decision logic.
// LocalRefine's 11 sweep samples are a subset of DepthToWeak's 61 (same plane, same views, same weights; only
// ConfidenceCompute runs in between and it touches neither), so the second stage re-uses the stored column costs and
// only the evaluation at the current depth is new: 12 S evaluations per pixel become 0 extra.
#include <cub/cub.cuh>

#include "apde_common.cuh"
#include "apde_kernels.h"


namespace apde {

constexpr int kSweepR = 30, kSweepN = 61, kSlots = 62;
constexpr int kRefineR = 5, kRefineN = 11, kRefineSlots = 2 * kRefineN;  // LocalRefine's own steps: raw NCC [11] then raw geometric cost [11]  // slot 61 = the hypothesis at the current depth (LocalRefine's cost_now)

// class of a pixel for the sweep: 0 = none, 1 = LocalRefine only (+-5), 2 = DepthToWeak + LocalRefine (+-30)
__device__ __forceinline__ int sweep_class(const PassK &K, int px, int py, int center, bool dtw) {
    const float w = K.planes[center].w;  // rotating the normal does not touch w (APD.cu:2127-2129)
    if (w == 0.0f || K.sel[center] == 0u) return 0;
    const int m = 6;
    const bool border = px < m || py < m || px >= K.W - m || py >= K.H - m;
    return (dtw && !border) ? 2 : 1;
}

// All sweep kernels work on a BAND of whole rows starting at row y0; slots of the band are its pixels in row-major order,
// Pb = rows * W of them, and flags / colidx are [N][Pb].  (Measured on B200, r01: numbering the slots in 8x4 tiles instead
// made DepthToWeak 3 % slower -- with 8-bit texels a 32 B sector is a 32x1 strip, which a row-major warp covers best.)
__device__ __forceinline__ bool slot_pixel(const PassK &K, int y0, int rows, int loc, int &px, int &py) {
    px = loc % K.W;
    py = y0 + loc / K.W;
    return true;
}

__global__ void __launch_bounds__(256) k_sweep_prepare(const __grid_constant__ PassK K, int dtw, int y0, int rows, int Pb,
                                                       int *__restrict__ flags) {
    const int loc = blockIdx.x * blockDim.x + threadIdx.x;
    if (loc >= Pb) return;
    int px, py;
    uint32_t sel = 0u;
    if (slot_pixel(K, y0, rows, loc, px, py)) {
        const int idx = py * K.W + px;
        const int cls = sweep_class(K, px, py, idx, dtw != 0);
        if (dtw && cls != 2) K.weak[idx] = APDE_UNKNOWN;  // border / zero depth / no selected view (APD.cu:2114-2153)
        if (cls) sel = K.sel[idx];
    }
    for (int v = 0; v < K.N; ++v) flags[(size_t)v * Pb + loc] = (sel >> v) & 1u;
}

// column number -> flat (view, pixel) index
__global__ void __launch_bounds__(256) k_sweep_scatter(const int *__restrict__ flags, const int *__restrict__ colidx, size_t nflat,
                                                       int *__restrict__ colmap) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nflat && flags[i]) colmap[colidx[i]] = (int)i;
}

template <bool U, bool SA>
__device__ __forceinline__ void k_sweep_columns_body(const PassK &K, int dtw, int y0, int rows, int Pb,
                                                       const int *__restrict__ colmap, int ncols, float *__restrict__ ncc,
                                                       float *__restrict__ lr) {
    extern __shared__ float smem[];
    const ViewK *s_vk = stage_views(K, smem);
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncols) return;
    const int flat = colmap[col];
    const int v = flat / Pb;
    int px, py;
    slot_pixel(K, y0, rows, flat % Pb, px, py);
    const int center = py * K.W + px;
    const int cls = sweep_class(K, px, py, center, dtw != 0);
    const ViewK &vk = s_vk[v];
    const float4 opl = normal_to_refcam(K, K.planes[center]);
    const float origin_depth = opl.w;
    // mean baseline over the selected views, ascending (APD.cu:2138-2155)
    const uint32_t sel = K.sel[center];
    float base_line = 0.0f;
    int valid_src = 0;
    for (uint32_t mk = sel; mk; mk &= mk - 1) { base_line = __fadd_rn(base_line, view_baseline(K, s_vk[__ffs(mk) - 1])); valid_src++; }
    const SweepDepths sd = sweep_depths(K, base_line, valid_src, origin_depth);
    RefPatch rp;
    load_ref_patch<U>(K, px, py, rp);
    typename SaTypes<SA>::Info si;
    load_sa<SA>(K, px, py, rp, si);
    unsigned n_old = 0, n_geom = 0;
    const int r = (cls == 2) ? kSweepR : 5;
    // one instance of the evaluation code: the sweep steps, then (pd == r + 1) the hypothesis at the current depth in slot 61
#pragma unroll 1
    for (int pd = -r; pd <= r + 1; ++pd) {
        const bool current = pd > r;
        const float p_depth = current ? origin_depth : sd.depth(pd);
        if (!current && (p_depth < K.depth_min || p_depth > K.depth_max)) continue;
        float4 tp = opl;
        tp.w = distance_to_origin(K, px, py, p_depth, tp);
        const PlaneM m = plane_row(K, tp);
        const size_t o = (size_t)(current ? kSweepN : pd + kSweepR) * ncols + col;
        const float cn = ncc_old_x<U, SA>(K, vk, px, py, m, rp, si);
        n_old++;
        if (K.geom) {
            // geometric passes store what the consumers add up: the combined cost FFMA(geom_factor, geom, ncc) of every step
            // (DepthToWeak's curve, LocalRefine's cost_now, APD.cu:2179-2181, 2380-2383) and, for the 11 steps LocalRefine
            // sweeps, the two terms on their own (it weights them separately, APD.cu:2417-2419): 62 + 22 floats per column
            // instead of 124
            const float cg = geom_cost(K, vk, v, px, py, tp);
            n_geom++;
            ncc[o] = __fmaf_rn(K.geom_factor, cg, cn);
            if (!current && pd >= -kRefineR && pd <= kRefineR) {
                lr[(size_t)(pd + kRefineR) * ncols + col] = cn;
                lr[(size_t)(kRefineN + pd + kRefineR) * ncols + col] = cg;
            }
        } else {
            ncc[o] = cn;
        }
    }
    count_evals(K, n_old, 0, n_geom);
}
// 4 resident blocks per SM (128 registers, 48 B of stack): measured on B200 at 1920x1080, DepthToWeak per schedule step
// 1879 ms with 3 blocks (142 registers, the compiler's own choice), 1834 ms with 4, 2093 ms with 5 (96 registers), 2072 ms with 6
__global__ void __launch_bounds__(128, 4) k_sweep_columns(const __grid_constant__ PassK K, int dtw, int y0, int rows, int Pb,
                                                       const int *__restrict__ colmap, int ncols, float *__restrict__ ncc,
                                                       float *__restrict__ geo) {
    if (K.tex_unorm > 0.0f) k_sweep_columns_body<true, false>(K, dtw, y0, rows, Pb, colmap, ncols, ncc, geo);
    else k_sweep_columns_body<false, false>(K, dtw, y0, rows, Pb, colmap, ncols, ncc, geo);
}
// the same kernel for a problem with a segment-label map (see "segment labels" in apde_device.cuh)
__global__ void __launch_bounds__(128) k_sweep_columns_sa(const __grid_constant__ PassK K, int dtw, int y0, int rows, int Pb,
                                                          const int *__restrict__ colmap, int ncols, float *__restrict__ ncc,
                                                          float *__restrict__ geo) {
    if (K.tex_unorm > 0.0f) k_sweep_columns_body<true, true>(K, dtw, y0, rows, Pb, colmap, ncols, ncc, geo);
    else k_sweep_columns_body<false, true>(K, dtw, y0, rows, Pb, colmap, ncols, ncc, geo);
}


// DepthToWeak decision logic, APD.cu:2157-2249
__global__ void __launch_bounds__(128) k_sweep_classify(const __grid_constant__ PassK K, int y0, int rows, int Pb,
                                                        const int *__restrict__ colidx, int ncols,
                                                        const float *__restrict__ ncc, float *curve) {
    extern __shared__ float smem[];
    const ViewK *s_vk = stage_views(K, smem);
    const int loc = blockIdx.x * blockDim.x + threadIdx.x;
    int px, py;
    if (loc >= Pb || !slot_pixel(K, y0, rows, loc, px, py)) return;
    const int idx = py * K.W + px;
    if (sweep_class(K, px, py, idx, true) != 2) return;
    const float origin_depth = K.planes[idx].w;
    const uint32_t sel = K.sel[idx];
    const uint4 w = K.vw[idx];
    float base_line = 0.0f, weight_normal = 0.0f;
    int valid_src = 0;
    for (uint32_t mk = sel; mk; mk &= mk - 1) {
        const int v = __ffs(mk) - 1;
        weight_normal = __fadd_rn(weight_normal, (float)vw_get(w, v));
        base_line = __fadd_rn(base_line, view_baseline(K, s_vk[v]));
        valid_src++;
    }
    const SweepDepths sd = sweep_depths(K, base_line, valid_src, origin_depth);
    const float rwn = rcp_approx(weight_normal);
    const int radius = kSweepR, n = kSweepN;
    // The reference sums, per depth step, over the selected views in ascending order.  Here the views are the OUTER loop and
    // the 61 steps are unrolled: every pc[i] still receives its FFMAs in ascending view order (bit-identical), but the 61
    // (+61) loads of a view are independent and in flight together, and pc[] lives in registers (static indices only below).
    float pc[kSweepN];
#pragma unroll
    for (int i = 0; i < kSweepN; ++i) pc[i] = 0.0f;
    // (the view loop is UNIFORM over the warp, lanes that did not select view v sit the iteration out: the lanes that take part
    // then read consecutive columns of the same view -- whole sectors.  Walking each lane's own selected views in turn put 8.8
    // sectors behind every request and read 14.3 GB instead of 3.5 GB per launch, profiles/r02_ncu_top_kernels.md)
    for (int v = 0; v < K.N; ++v) {
        if (!((sel >> v) & 1u)) continue;
        const size_t c0 = (size_t)colidx[(size_t)v * Pb + loc];
        const float wv = (float)vw_get(w, v);
        // loads in batches of kBatch steps (independent loads in flight), then the FFMAs of the batch
        // (in geometric passes a slot already holds FFMA(geom_factor, geom, ncc), APD.cu:2179-2181 as built)
        constexpr int kBatch = 32;
#pragma unroll
        for (int b = 0; b < kSweepN; b += kBatch) {
            float tn[kBatch];
#pragma unroll
            for (int i = 0; i < kBatch; ++i)
                if (b + i < kSweepN) tn[i] = __ldg(ncc + (size_t)(b + i) * ncols + c0);
#pragma unroll
            for (int i = 0; i < kBatch; ++i)
                if (b + i < kSweepN) pc[b + i] = __fmaf_rn(wv, tn[i], pc[b + i]);
        }
    }
#pragma unroll
    for (int i = 0; i < kSweepN; ++i) {
        const float p_depth = sd.depth(i - radius);
        // (steps outside the depth range were never evaluated: their slots hold stale values, replaced here)
        const float p_cost = __fmul_rn(pc[i], rwn);
        pc[i] = (p_depth < K.depth_min || p_depth > K.depth_max) ? 2.0f : ((2.0f > p_cost) ? p_cost : 2.0f);  // OpenCV MIN(2.0f, p_cost): NaN -> 2
    }
    if (curve) {
#pragma unroll
        for (int i = 0; i < n; ++i) curve[(size_t)idx * n + i] = pc[i];
    }
    unsigned long long peaks = 0ull;
    int peak_count = 0, min_peak = 0;
    float min_cost = 2.0f, cost_at_min_peak = pc[0];  // pc[min_peak] without a dynamic index
#pragma unroll
    for (int i = 2; i < n - 2; ++i) {
        if (pc[i - 1] > pc[i] && pc[i + 1] > pc[i]) {
            peaks |= 1ull << i;
            peak_count++;
            if (pc[i] < min_cost) { min_peak = i; min_cost = pc[i]; cost_at_min_peak = pc[i]; }
        }
    }
    if (abs(min_peak - radius) > K.weak_peak_radius || cost_at_min_peak > 0.5f) { K.weak[idx] = APDE_WEAK; return; }
    if (peak_count == 1) { K.weak[idx] = (cost_at_min_peak <= 0.15f) ? APDE_STRONG : APDE_WEAK; return; }
    float var = 0.0f;
#pragma unroll
    for (int i = 2; i < n - 2; ++i)
        if (((peaks >> i) & 1ull) && i != min_peak) { const float d = pc[i] - min_cost; var += d * d; }
    var = sqrtf(var);
    var /= (peak_count - 1);
    K.weak[idx] = (var > 0.2f) ? APDE_STRONG : APDE_WEAK;
}

// LocalRefine decision logic, APD.cu:2368-2431
__global__ void __launch_bounds__(128) k_sweep_refine(const __grid_constant__ PassK K, int y0, int rows, int Pb,
                                                      const int *__restrict__ colidx, int ncols,
                                                      const float *__restrict__ ncc, const float *__restrict__ lr) {
    extern __shared__ float smem[];
    const ViewK *s_vk = stage_views(K, smem);
    const int loc = blockIdx.x * blockDim.x + threadIdx.x;
    int px, py;
    if (loc >= Pb || !slot_pixel(K, y0, rows, loc, px, py)) return;
    const int idx = py * K.W + px;
    if (sweep_class(K, px, py, idx, false) == 0) return;
    const float origin_depth = K.planes[idx].w;
    const uint32_t sel = K.sel[idx];
    const uint4 w = K.vw[idx];
    float cost_now = 0.0f, base_line = 0.0f, weight_normal = 0.0f;
    int valid_src = 0;
    for (uint32_t mk = sel; mk; mk &= mk - 1) {
        const int v = __ffs(mk) - 1;
        const size_t o = (size_t)kSweepN * ncols + colidx[(size_t)v * Pb + loc];
        const float tc = ncc[o];  // geometric passes: already FFMA(geom_factor, geom, ncc)
        const float wv = (float)vw_get(w, v);
        cost_now = __fmaf_rn(wv, tc, cost_now);  // APD.cu:2380-2383 as built
        weight_normal = __fadd_rn(weight_normal, wv);
        base_line = __fadd_rn(base_line, view_baseline(K, s_vk[v]));
        valid_src++;
    }
    if (weight_normal == 0.0f) return;
    const float rwn = rcp_approx(weight_normal);
    const SweepDepths sd = sweep_depths(K, base_line, valid_src, origin_depth);
    float min_cost = 2.0f, best_depth = origin_depth;
#pragma unroll 1
    for (int pd = -5; pd <= 5; ++pd) {
        const float p_depth = sd.depth(pd);
        if (p_depth < K.depth_min || p_depth > K.depth_max) continue;
        float tc = 0.0f;
        for (uint32_t mk = sel; mk; mk &= mk - 1) {
            const int v = __ffs(mk) - 1;
            const size_t c0 = (size_t)colidx[(size_t)v * Pb + loc];
            const float wv = (float)vw_get(w, v);
            if (K.geom) {  // the two terms on their own (see k_sweep_columns)
                tc = __fmaf_rn(wv, lr[(size_t)(pd + kRefineR) * ncols + c0], tc);  // APD.cu:2417-2419 as built
                tc = __fmaf_rn(wv, __fmul_rn(K.geom_factor, lr[(size_t)(kRefineN + pd + kRefineR) * ncols + c0]), tc);
            } else {
                tc = __fmaf_rn(wv, ncc[(size_t)(pd + kSweepR) * ncols + c0], tc);
            }
        }
        tc = __fmul_rn(tc, rwn);
        if (tc < min_cost) { min_cost = tc; best_depth = p_depth; }
    }
    // "cost_now /= weight_normal; ... cost_now - min_cost > 0.1" is one FFMA in the reference build, compared in double
    if ((double)__fmaf_rn(rwn, cost_now, -min_cost) > 0.1) K.planes[idx].w = best_depth;
}

// ------------------------------------------------------------------------------------------------ host side
cudaError_t SweepWorkspace::reserve_flags(size_t n) {
    if (n <= flag_cap) return cudaSuccess;
    cudaFree(flags); cudaFree(colidx); cudaFree(scan_tmp);
    flags = colidx = nullptr; scan_tmp = nullptr;
    cudaError_t e = cudaMalloc(&flags, (n + 1) * sizeof(int));
    if (e != cudaSuccess) return e;
    if ((e = cudaMalloc(&colidx, (n + 1) * sizeof(int))) != cudaSuccess) return e;
    scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, flags, colidx, (int)(n + 1));
    if ((e = cudaMalloc(&scan_tmp, scan_bytes)) != cudaSuccess) return e;
    flag_cap = n;
    return cudaSuccess;
}
cudaError_t SweepWorkspace::reserve_columns(size_t ncols_, bool geom) {
    const size_t need = ncols_ * kSlots;
    if (ncols_ > map_cap) {
        cudaFree(colmap);
        colmap = nullptr;
        cudaError_t e = cudaMalloc(&colmap, (ncols_ + ncols_ / 8) * sizeof(int));
        if (e != cudaSuccess) return e;
        map_cap = ncols_ + ncols_ / 8;
    }
    if (need > col_cap) {
        cudaFree(ncc); cudaFree(geo);
        ncc = geo = nullptr; geo_cap = 0;
        const size_t cap = need + need / 8;
        cudaError_t e = cudaMalloc(&ncc, cap * sizeof(float));
        if (e != cudaSuccess) return e;
        col_cap = cap;
    }
    const size_t need_lr = ncols_ * kRefineSlots;
    if (geom && geo_cap < need_lr) {
        cudaFree(geo);
        geo = nullptr; geo_cap = 0;
        const size_t cap = need_lr + need_lr / 8;
        cudaError_t e = cudaMalloc(&geo, cap * sizeof(float));
        if (e != cudaSuccess) return e;
        geo_cap = cap;
    }
    return cudaSuccess;
}
void SweepWorkspace::release() {
    cudaFree(flags); cudaFree(colidx); cudaFree(scan_tmp); cudaFree(ncc); cudaFree(geo); cudaFree(colmap);
    flags = colidx = colmap = nullptr; scan_tmp = nullptr; ncc = geo = nullptr;
    flag_cap = col_cap = geo_cap = map_cap = 0; valid = false;
}

// Rows per band so that the WORST-CASE column storage of a band (every view selected at every pixel) stays under the
// budget: N * Pb * 62 floats (84 with the geometric term).  1920x1080 x 10 views fits one band (7.0 GB); 6048x4032 x 10
// views would need 82 GB in one piece and is processed in 2 bands.  APDE_SWEEP_BUDGET_MB overrides the 48 GB default.
static int band_rows(const PassK &K, const SweepWorkspace &ws) {
    static const size_t env_mb = [] {
        const char *e = getenv("APDE_SWEEP_BUDGET_MB");
        return e ? (size_t)atoll(e) : (size_t)0;
    }();
    const size_t mb = ws.budget_mb ? ws.budget_mb : (env_mb ? env_mb : (size_t)49152);
    const size_t budget = mb << 20;
    const size_t per_row = (size_t)K.N * K.W * (kSlots + (K.geom ? kRefineSlots : 0)) * sizeof(float);
    const size_t rows = budget / (per_row ? per_row : 1);
    return (int)std::max<size_t>(1, std::min<size_t>(rows, (size_t)K.H));
}

// columns of the band of rows [y0, y0 + rows) for the current problem state; dtw = 1: DepthToWeak (+-30 for interior pixels), 0: LocalRefine only
static cudaError_t sweep_build_band(const PassK &K, SweepWorkspace &ws, int dtw, int y0, int rows, cudaStream_t st, uint64_t *launches) {
    const int Pb = K.W * rows;
    const size_t nflat = (size_t)K.N * Pb;
    cudaError_t e = ws.reserve_flags(nflat);
    if (e != cudaSuccess) return e;
    k_sweep_prepare<<<(Pb + 255) / 256, 256, 0, st>>>(K, dtw, y0, rows, Pb, ws.flags);
    if ((e = cudaMemsetAsync(ws.flags + nflat, 0, sizeof(int), st)) != cudaSuccess) return e;
    if ((e = cub::DeviceScan::ExclusiveSum(ws.scan_tmp, ws.scan_bytes, ws.flags, ws.colidx, (int)(nflat + 1), st)) != cudaSuccess) return e;
    int ncols = 0;
    if ((e = cudaMemcpyAsync(&ncols, ws.colidx + nflat, sizeof(int), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    ws.ncols = ncols;
    ws.y0 = y0; ws.rows = rows; ws.Pb = Pb;
    if (launches) *launches += 2;
    if (ncols > 0) {
        if ((e = ws.reserve_columns((size_t)ncols, K.geom != 0)) != cudaSuccess) return e;
        const size_t vsm = sizeof(float) * views_smem_floats(K.N);
        k_sweep_scatter<<<(unsigned)((nflat + 255) / 256), 256, 0, st>>>(ws.flags, ws.colidx, nflat, ws.colmap);
        if (K.sa) k_sweep_columns_sa<<<(ncols + 127) / 128, 128, vsm, st>>>(K, dtw, y0, rows, Pb, ws.colmap, ncols, ws.ncc, ws.geo);
        else k_sweep_columns<<<(ncols + 127) / 128, 128, vsm, st>>>(K, dtw, y0, rows, Pb, ws.colmap, ncols, ws.ncc, ws.geo);
        if (launches) *launches += 2;
    }
    return cudaGetLastError();
}
static cudaError_t sweep_classify_band(const PassK &K, SweepWorkspace &ws, float *curve, cudaStream_t st) {
    if (ws.ncols == 0) return cudaSuccess;
    k_sweep_classify<<<(ws.Pb + 127) / 128, 128, sizeof(float) * views_smem_floats(K.N), st>>>(K, ws.y0, ws.rows, ws.Pb, ws.colidx, ws.ncols, ws.ncc,
                                                                                              curve);
    return cudaGetLastError();
}
static cudaError_t sweep_refine_band(const PassK &K, SweepWorkspace &ws, cudaStream_t st) {
    if (ws.ncols == 0) return cudaSuccess;
    k_sweep_refine<<<(ws.Pb + 127) / 128, 128, sizeof(float) * views_smem_floats(K.N), st>>>(K, ws.y0, ws.rows, ws.Pb, ws.colidx, ws.ncols, ws.ncc,
                                                                                            ws.geo);
    return cudaGetLastError();
}

// DepthToWeak.  With one band the columns stay valid for LocalRefine (ws.valid); with several bands they do not.
cudaError_t sweep_depth_to_weak(const PassK &K, SweepWorkspace &ws, float *curve, cudaStream_t st, uint64_t *launches) {
    const int rows = band_rows(K, ws);
    ws.valid = false;
    for (int y0 = 0; y0 < K.H; y0 += rows) {
        const int nr = std::min(rows, K.H - y0);
        cudaError_t e = sweep_build_band(K, ws, 1, y0, nr, st, launches);
        if (e != cudaSuccess) return e;
        if ((e = sweep_classify_band(K, ws, curve, st)) != cudaSuccess) return e;
        if (launches && y0 + nr < K.H) *launches += 1;
    }
    ws.valid = rows >= K.H;
    return cudaSuccess;
}
// LocalRefine: re-uses DepthToWeak's columns when they are still valid, otherwise builds the +-5 columns band by band.
cudaError_t sweep_local_refine(const PassK &K, SweepWorkspace &ws, cudaStream_t st, uint64_t *launches) {
    if (ws.valid) {
        ws.valid = false;
        return sweep_refine_band(K, ws, st);
    }
    const int rows = band_rows(K, ws);
    for (int y0 = 0; y0 < K.H; y0 += rows) {
        const int nr = std::min(rows, K.H - y0);
        cudaError_t e = sweep_build_band(K, ws, 0, y0, nr, st, launches);
        if (e != cudaSuccess) return e;
        if ((e = sweep_refine_band(K, ws, st)) != cudaSuccess) return e;
        if (launches && y0 + nr < K.H) *launches += 1;
    }
    return cudaSuccess;
}

}  // namespace apde
