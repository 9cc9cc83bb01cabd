// apde_apd.cu -- weak-texture ("APD") stages: nearest-strong search, anchor generation + RANSAC, per-iteration plane
// fit and the deformable (anchor-patch) propagation.  sm_100a.
#include "apde_device.cuh"
#include "apde_kernels.h"
#include "apde_common.cuh"

#include <cfloat>

namespace apde {

// -------------------------------------------------------------------------------------------- K2 nearest strong
// FindNearestStrongPoint, APD.cu:2434-2484.  The reference scans the whole 201x201 window (x outer, y inner) and keeps
// the minimum of (distance, -confidence, scan order).  Here the window is visited ring by ring (Chebyshev radius r,
// all of whose cells are at Euclidean distance >= r) and the search stops once r^2 exceeds the best squared distance:
// same winner, including the tie rules, in O(best_dist^2) instead of 40401 probes.
//
// Pruning: a first kernel summarises every 8x8 tile (0 = no STRONG pixel, else 1 + the highest confidence of its STRONG
// pixels).  A probe can only win if its pixel is STRONG with confidence >= the centre's, so while walking a ring edge a
// tile whose summary rules that out is skipped up to its far boundary -- pixels deep inside texture-less regions, which
// would otherwise run all 40401 probes, touch ~1/8 of them.  The winner is unchanged: only probes that cannot pass the
// reference's own test are skipped, and the comparison below is a total order independent of the visiting order.
__global__ void __launch_bounds__(64) k_ns_tiles(const __grid_constant__ PassK K, int tiles_x) {
    const int tile = blockIdx.x, lane = threadIdx.x;  // 64 threads = the 8x8 pixels of one tile
    const int px = (tile % tiles_x) * 8 + (lane & 7), py = (tile / tiles_x) * 8 + (lane >> 3);
    int v = 0;
    if (px < K.W && py < K.H) {
        const int c = py * K.W + px;
        if (K.weak[c] == APDE_STRONG) v = 1 + (int)K.conf[c];
    }
    v = __reduce_max_sync(0xffffffffu, v);
    __shared__ int part[2];
    if ((lane & 31) == 0) part[lane >> 5] = v;
    __syncthreads();
    if (lane == 0) K.ns_tiles[tile] = (uint16_t)max(part[0], part[1]);
}

__device__ __forceinline__ void ns_probe(const PassK &K, int tc, int dx, int dy, uint8_t cc, int &best_d2, int &best_conf,
                                         int &best_dx, int &best_dy) {
    if (K.weak[tc] != APDE_STRONG) return;
    const int cf = K.conf[tc];
    if (cf < cc) return;
    const int d2 = dx * dx + dy * dy;
    // total order of the reference scan: smaller distance; then higher confidence; then earlier (dx, dy)
    bool better = false;
    if (d2 < best_d2) better = true;
    else if (d2 == best_d2) {
        if (cf > best_conf) better = true;
        else if (cf == best_conf && (dx < best_dx || (dx == best_dx && dy < best_dy))) better = true;
    }
    if (better) { best_d2 = d2; best_conf = cf; best_dx = dx; best_dy = dy; }
}

__global__ void __launch_bounds__(128) k_nearest_strong(const __grid_constant__ PassK K, int tiles_x) {
    int px, py;
    if (!full_pixel(K, tiles_x, px, py)) return;
    const int W = K.W, H = K.H;
    const int center = py * W + px;
    const uint8_t wk = K.weak[center];
    short2 out = make_short2(-1, -1);
    if (wk == APDE_STRONG) {
        out = make_short2((short)px, (short)py);
    } else if (wk == APDE_WEAK || wk == APDE_UNKNOWN) {
        const uint8_t cc = K.conf[center];
        const int need = (int)cc + 1;  // a tile can hold a winner only if its summary is >= need
        const int tw = (W + 7) >> 3;
        int best_d2 = INT_MAX, best_conf = -1, best_dx = 0, best_dy = 0;
        for (int r = 1; r <= 100; ++r) {
            if (r * r > best_d2) break;
            // the two vertical edges dx = -r, +r (t = dy runs over the full side)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int dx = e ? r : -r, tx = px + dx;
                if (tx < 0 || tx >= W) continue;
                const uint16_t *trow = K.ns_tiles + (tx >> 3);
                for (int ty = max(py - r, 0), ty1 = min(py + r, H - 1); ty <= ty1;) {
                    if ((int)trow[(ty >> 3) * tw] < need) { ty = (ty | 7) + 1; continue; }
                    ns_probe(K, ty * W + tx, dx, ty - py, cc, best_d2, best_conf, best_dx, best_dy);
                    ++ty;
                }
            }
            // the two horizontal edges dy = -r, +r without the corners
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int dy = e ? r : -r, ty = py + dy;
                if (ty < 0 || ty >= H) continue;
                const uint16_t *trow = K.ns_tiles + (ty >> 3) * tw;
                for (int tx = max(px - r + 1, 0), tx1 = min(px + r - 1, W - 1); tx <= tx1;) {
                    if ((int)trow[tx >> 3] < need) { tx = (tx | 7) + 1; continue; }
                    ns_probe(K, ty * W + tx, tx - px, dy, cc, best_d2, best_conf, best_dx, best_dy);
                    ++tx;
                }
            }
        }
        if (best_d2 != INT_MAX) out = make_short2((short)(px + best_dx), (short)(py + best_dy));
    }
    K.nearest[center] = out;
}

// -------------------------------------------------------------------------------------------- K3 + K4 anchors
__device__ __forceinline__ void normalize2(float2 &v) {
    const float inv = rsqrtf(v.x * v.x + v.y * v.y);
    v.x *= inv; v.y *= inv;
}
__device__ __forceinline__ bool point_in_triangle(short2 A, short2 B, short2 C, int px, int py) {  // APD.cu:122-143
    const float2 AB = make_float2(B.x - A.x, B.y - A.y), BC = make_float2(C.x - B.x, C.y - B.y),
                 CA = make_float2(A.x - C.x, A.y - C.y);
    const float ab = sqrtf(AB.x * AB.x + AB.y * AB.y), bc = sqrtf(BC.x * BC.x + BC.y * BC.y),
                ca = sqrtf(CA.x * CA.x + CA.y * CA.y);
    if (ab <= 2 || bc <= 2 || ca <= 2) return false;
    if (!(ab + bc > ca && bc + ca > ab && ab + ca > bc)) return false;
    const float2 PA = make_float2(A.x - px, A.y - py), PB = make_float2(B.x - px, B.y - py), PC = make_float2(C.x - px, C.y - py);
    const float t1 = PA.x * PB.y - PA.y * PB.x, t2 = PB.x * PC.y - PB.y * PC.x, t3 = PC.x * PA.y - PC.y * PA.x;
    return t1 * t2 >= 0 && t1 * t3 >= 0;
}
__device__ __forceinline__ float3 back_project(const PassK &K, float x, float y, float depth) {  // Get3DPoint APD.cu:190
    return make_float3(depth * (x - K.cx) / K.fx, depth * (y - K.cy) / K.fy, depth);
}
// "(curand() % 2 == 0 ? 1 : -1) * curand() % shift_range" evaluates in unsigned arithmetic (APD.cu:1921)
__device__ __forceinline__ int rand_shift(Rng &rng, int shift_range) {
    const int sgn = (rng.next() % 2 == 0 ? 1 : -1);
    const uint32_t r = rng.next();
    return (int)(((uint32_t)sgn * r) % (uint32_t)shift_range);
}

// The WEAK-only kernels below walk the compacted (colour, weak) pixel lists when the caller provides one (K.list): WEAK
// pixels are a few per cent of the frame, and a full-grid launch would leave most lanes of most warps idle.
__device__ __forceinline__ bool full_pixel_or_list(const PassK &K, int tiles_x, int &px, int &py) {
    if (K.list == nullptr) return full_pixel(K, tiles_x, px, py);
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int *lst = K.list + (size_t)blockIdx.y * K.list_pair_stride;
    if (idx >= K.list_count[2 * blockIdx.y]) return false;
    const int center = lst[idx];
    px = center % K.W;
    py = center / K.W;
    return true;
}

// GenAnchors + NeigbourUpdate, APD.cu:1857-2100 (NeigbourUpdate only touches the pixel's own state, so it is fused)
__global__ void __launch_bounds__(128) k_gen_anchors(const __grid_constant__ PassK K, int tiles_x) {
    int px, py;
    if (!full_pixel_or_list(K, tiles_x, px, py)) return;
    const int width = K.W, height = K.H;
    const int center = py * width + px;
    if (K.weak[center] != APDE_WEAK) return;
    const int min_margin = 6;
    const float depth_diff = K.depth_max - K.depth_min;
    Rng rng(K.seed, K.stream, (uint32_t)center, SITE_ANCHOR);
    short2 *anchors = K.anchors + (size_t)center * APDE_ANCHOR_NUM;
    for (int i = 0; i < APDE_ANCHOR_NUM; ++i) anchors[i] = make_short2(-1, -1);
    anchors[0] = make_short2((short)px, (short)py);
    short2 sp[32];
    unsigned dir_valid = 0;
    int strong_point_size = 0, odi = -1;
    const int rotate_time = K.rotate_time;
    const float angle = 45.0f / rotate_time;
    const float cos_a = (float)cos((double)angle * 3.14159265358979323846 / (double)180.f);
    const float sin_a = (float)sin((double)angle * 3.14159265358979323846 / (double)180.f);
    const float thresh = (float)cos((double)(angle / 2.0f) * 3.14159265358979323846 / 180.0);
    const int shift_range = max((int)(tan((double)(angle / 2.0f) * 3.14159265358979323846 / 180.0) * 20), 1);
    for (int odx = -1; odx <= 1; ++odx) for (int ody = -1; ody <= 1; ++ody) {
        if (odx == 0 && ody == 0) continue;
        float2 od = make_float2((float)odx, (float)ody);
        normalize2(od);
        odi++;
        for (int rot = 0; rot < rotate_time; ++rot) {
            const int di = odi * 4 + rot;
            for (int radius = 2; radius <= 4096; radius = min(radius * 2, radius + 25)) {
                const float tx = px + od.x * radius, ty = py + od.y * radius;
                if (tx < 0 || ty < 0 || tx >= width || ty >= height) break;
                bool found = false;
                for (int ri = 0; ri < 4; ++ri) {
                    const int rxs = rand_shift(rng, shift_range);
                    const int rys = rand_shift(rng, shift_range);
                    float2 dir = make_float2(od.x * 20 + rxs, od.y * 20 + rys);
                    normalize2(dir);
                    short2 ap = make_short2((short)(int)(px + dir.x * radius), (short)(int)(py + dir.y * radius));
                    if (ap.x < min_margin || ap.y < min_margin || ap.x >= width - min_margin || ap.y >= height - min_margin) continue;
                    ap = K.nearest[ap.x + ap.y * width];
                    if (ap.x == -1 || ap.y == -1) continue;
                    float2 td = make_float2((float)(ap.x - px), (float)(ap.y - py));
                    normalize2(td);
                    if (td.x * od.x + td.y * od.y > thresh) { sp[di] = ap; dir_valid |= 1u << di; strong_point_size++; found = true; break; }
                }
                if (found) break;
            }
            float2 rd = make_float2(od.x * cos_a - od.y * sin_a, od.x * sin_a + od.y * cos_a);
            normalize2(rd);
            od = rd;
        }
    }
    if (strong_point_size <= 3) { K.reliable[center] = 0; K.weak[center] = APDE_UNKNOWN; return; }
    short2 spv[32];
    float3 spv3[32];
    int valid_count = 0;
    const float3 cpw = back_project(K, (float)px, (float)py, K.planes[center].w);
    for (int i = 0; i < 32; ++i) {
        spv[i] = make_short2(-1, -1);
        if ((dir_valid >> i) & 1u) {
            const short2 s = sp[i];
            spv[valid_count] = s;
            spv3[valid_count] = back_project(K, (float)s.x, (float)s.y, K.planes[s.x + s.y * width].w);
            valid_count++;
        }
    }
    float4 best_plane = make_float4(0, 0, 0, 0);
    int ua = -1, ub = -1, uc = -1;
    bool has_valid = false;
    {
        float min_cost = FLT_MAX;
        int max_count = 3;
        for (int iteration = 0; iteration < 50; ++iteration) {
            const int a = rng.next() % valid_count, b = rng.next() % valid_count, c = rng.next() % valid_count;
            if (a == b || b == c || a == c) continue;
            if (!point_in_triangle(spv[a], spv[b], spv[c], px, py)) continue;
            const float3 A = spv3[a], B = spv3[b], C = spv3[c];
            const float3 AC = make_float3(A.x - C.x, A.y - C.y, A.z - C.z), BC = make_float3(B.x - C.x, B.y - C.y, B.z - C.z);
            float4 cv = make_float4(AC.y * BC.z - BC.y * AC.z, -(AC.x * BC.z - BC.x * AC.z), AC.x * BC.y - BC.x * AC.y, 0.0f);
            if ((cv.x == 0 && cv.y == 0 && cv.z == 0) || isnan(cv.x) || isnan(cv.y) || isnan(cv.z)) continue;
            normalize3(cv);
            cv.w = -(cv.x * A.x + cv.y * A.y + cv.z * A.z);
            int tcnt = 0;
            for (int si = 0; si < valid_count; ++si) {
                const float3 tp = spv3[si];
                const float d = fabsf(cv.x * tp.x + cv.y * tp.y + cv.z * tp.z + cv.w);
                if (d / depth_diff < K.ransac_threshold) tcnt++;
            }
            if (tcnt < 6) continue;
            const float cd = fabsf(cv.x * cpw.x + cv.y * cpw.y + cv.z * cpw.z + cv.w);
            if (tcnt > max_count) { max_count = tcnt; min_cost = cd; best_plane = cv; has_valid = true; ua = a; ub = b; uc = c; }
            else if (tcnt == max_count && cd < min_cost) { min_cost = cd; best_plane = cv; ua = a; ub = b; uc = c; }
        }
    }
    if (!has_valid) { K.reliable[center] = 0; K.weak[center] = APDE_UNKNOWN; return; }
    float weight[32];
    for (int i = 0; i < valid_count; ++i) {
        const float3 tp = spv3[i];
        float d = fabsf(best_plane.x * tp.x + best_plane.y * tp.y + best_plane.z * tp.z + best_plane.w);
        if (d / depth_diff >= K.ransac_threshold) { spv[i] = make_short2(-1, -1); weight[i] = FLT_MAX; continue; }
        if (i == ua || i == ub || i == uc) d -= 1;
        weight[i] = d;
    }
    // sort_small_weighted, APD.cu:25-38
    for (int i = 1; i < valid_count; i++) {
        const short2 tmp = spv[i];
        const float tw = weight[i];
        int j;
        for (j = i; j >= 1 && tw < weight[j - 1]; j--) { spv[j] = spv[j - 1]; weight[j] = weight[j - 1]; }
        spv[j] = tmp; weight[j] = tw;
    }
    for (int i = 1; i < APDE_ANCHOR_NUM; ++i) anchors[i] = spv[i - 1];
    K.reliable[center] = 1;
}

// -------------------------------------------------------------------------------------------- K7 RANSAC fit
// RANSACToGetFitPlane, APD.cu:2486-2598
// fit = current plane for every pixel that is not WEAK (APD.cu:2497-2500): the streaming half of RANSACToGetFitPlane when the
// WEAK pixels are processed from their lists
__global__ void __launch_bounds__(256) k_fit_copy_nonweak(const __grid_constant__ PassK K) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K.W * K.H && K.weak[i] != APDE_WEAK) K.fit[i] = K.planes[i];
}

__global__ void __launch_bounds__(128) k_ransac_fit(const __grid_constant__ PassK K, int iter, int tiles_x) {
    int px, py;
    if (!full_pixel_or_list(K, tiles_x, px, py)) return;
    const int width = K.W;
    const int center = py * width + px;
    if (K.weak[center] != APDE_WEAK) { K.fit[center] = K.planes[center]; return; }
    Rng rng(K.seed, K.stream, (uint32_t)center, SITE_FIT + iter);
    short2 sp[8];
    float3 sp3[8];
    int sc = 0;
    for (int i = 1; i < APDE_ANCHOR_NUM; ++i) {
        const short2 tp = K.anchors[(size_t)center * APDE_ANCHOR_NUM + i];
        if (tp.x == -1 || tp.y == -1) continue;
        sp[sc] = tp;
        const float depth = depth_from_plane(K, K.planes[tp.x + tp.y * width], tp.x, tp.y);
        sp3[sc] = back_project(K, (float)tp.x, (float)tp.y, depth);
        sc++;
    }
    if (sc < 3) { K.fit[center] = K.planes[center]; return; }
    float min_cost = FLT_MAX;
    float4 best = make_float4(0, 0, 0, 0);
    bool has_best = false;
    for (int iteration = 0; iteration < 50; ++iteration) {
        const int a = rng.next() % sc, b = rng.next() % sc, c = rng.next() % sc;
        if (a == b || b == c || a == c) continue;
        if (!point_in_triangle(sp[a], sp[b], sp[c], px, py)) continue;
        const float3 A = sp3[a], B = sp3[b], C = sp3[c];
        const float3 AC = make_float3(A.x - C.x, A.y - C.y, A.z - C.z), BC = make_float3(B.x - C.x, B.y - C.y, B.z - C.z);
        float4 cv = make_float4(AC.y * BC.z - BC.y * AC.z, -(AC.x * BC.z - BC.x * AC.z), AC.x * BC.y - BC.x * AC.y, 0.0f);
        if ((cv.x == 0 && cv.y == 0 && cv.z == 0) || isnan(cv.x) || isnan(cv.y) || isnan(cv.z)) continue;
        normalize3(cv);
        cv.w = -(cv.x * A.x + cv.y * A.y + cv.z * A.z);
        float tc = 0.0f;
        for (int si = 0; si < sc; ++si) {
            if (si == a || si == b || si == c) continue;
            const float3 tp = sp3[si];
            tc += fabsf(cv.x * tp.x + cv.y * tp.y + cv.z * tp.z + cv.w);
        }
        if (tc < min_cost) { min_cost = tc; best = cv; has_best = true; }
        if (min_cost == 0) break;
    }
    if (has_best) {
        const float depth = depth_from_plane(K, K.planes[center], px, py);
        const float3 vd = view_direction(K, px, py, depth);
        if (best.x * vd.x + best.y * vd.y + best.z * vd.z > 0) { best.x = -best.x; best.y = -best.y; best.z = -best.z; best.w = -best.w; }
        K.fit[center] = best;
    } else {
        K.fit[center] = make_float4(0, 0, 0, 0);
    }
}

// -------------------------------------------------------------------------------------------- K8 weak propagation
// Black/RedPixelUpdateWeak -> CheckerboardPropagationWeak -> PlaneHypothesisRefinementWeak,
// APD.cu:1617-1652, 1442-1615, 1008-1096
template <bool U, bool SA>
__device__ __forceinline__ void k_prop_weak_body(const PassK &K, int iter, int color, int tiles_x,
                                                   int ylimit) {
    extern __shared__ float smem[];
    const ViewK *s_vk = stage_views(K, smem);
    int px, py;
    if (!half_pixel_or_list(K, color, tiles_x, ylimit, px, py)) return;
    const int W = K.W, N = K.N;
    const int center = py * W + px;
    if (K.weak[center] != APDE_WEAK) return;
    const int stride = blockDim.x;
    float *sc = smem + views_smem_floats(N) + threadIdx.x;  // [8N][stride]
    float *sp = sc + 8 * N * stride;                         // [N][stride]
    const short2 *anc = K.anchors + (size_t)center * APDE_ANCHOR_NUM;

    RefPatch rp;
    load_ref_patch<U>(K, px, py, rp);
    typename SaTypes<SA>::Info si;
    load_sa<SA>(K, px, py, rp, si);
    typename SaTypes<SA>::Anchors ar;
    load_anchor_ref_x<U, SA>(K, anc, ar, si);
    unsigned n_new = 0, n_geom = 0;

    unsigned flags = 0, anchor_valid = 0;
    int pos[8];
    uint32_t nbr_sel[8];
#pragma unroll 1
    for (int h = 0; h < 8; ++h) {
        const short2 a = anc[h + 1];
        pos[h] = 0;
        nbr_sel[h] = 0;
        bool ok = false;
        if (!(a.x == -1 || a.y == -1)) {
            anchor_valid |= 1u << h;
            const int ap = a.x + a.y * W;
            nbr_sel[h] = K.sel[ap];
            if (K.weak[ap] == APDE_STRONG) { ok = true; pos[h] = ap; flags |= 1u << h; }
        }
        if (ok) {
            const PlaneM m = plane_row(K, K.planes[pos[h]]);
#pragma unroll 1
            for (int v = 0; v < N; ++v) sc[(h * N + v) * stride] = ncc_new_x<U, SA>(K, K.v[v], v, px, py, m, rp, ar, si);
            n_new += N;
        } else {
            for (int v = 0; v < N; ++v) sc[(h * N + v) * stride] = (h == 0 && v == 0) ? 2.0f : 0.0f;
        }
    }
    Rng rng(K.seed, K.stream, (uint32_t)center, SITE_WEAK + iter);
    uint32_t wmask;
    float wnorm;
    const uint4 w = select_views<8>(K, sc, sp, stride, nbr_sel, anchor_valid, iter, rng, &wmask, &wnorm);
    K.vw[center] = w;

    const bool use_geom = K.geom != 0;
    float final_costs[8];
#pragma unroll 1
    for (int h = 0; h < 8; ++h) {
        float acc = 0.0f;
        const bool fl = (flags >> h) & 1u;
        float4 pl = make_float4(0, 0, 0, 0);
        if (fl && use_geom) pl = K.planes[pos[h]];
        for (uint32_t mk = wmask; mk; mk &= mk - 1) {
            const int v = __ffs(mk) - 1;
            const float c = sc[(h * N + v) * stride];
            const float wv = (float)vw_get(w, v);
            if (use_geom) {
                if (fl) { acc += wv * (c + K.geom_factor * geom_cost(K, s_vk[v], v, px, py, pl)); n_geom++; }
                else acc += wv * (c + K.geom_factor * 3.0f);
            } else {
                acc += wv * c;
            }
        }
        final_costs[h] = acc / wnorm;
    }
    int min_idx = 0;
    float fc_min = final_costs[0];
#pragma unroll
    for (int h = 1; h < 8; ++h) if (final_costs[h] <= fc_min) { fc_min = final_costs[h]; min_idx = h; }

    const float dmin = K.depth_min, dmax = K.depth_max;
    const float4 plane_c = K.planes[center];
    float4 plane_now = plane_c;
    float cost_now;
    {
        const PlaneM m = plane_row(K, plane_c);
        float acc = 0.0f;
        for (uint32_t mk = wmask; mk; mk &= mk - 1) {
            const int v = __ffs(mk) - 1;
            float c = ncc_new_x<U, SA>(K, s_vk[v], v, px, py, m, rp, ar, si);
            n_new++;
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
            K.sel[center] = wmask;
        }
    }
    // PlaneHypothesisRefinementWeak: fit plane first; a zero fit plane skips the random refinement too (APD.cu:1028)
    const float4 fitp = K.fit[center];
    if (!(fitp.x == 0 && fitp.y == 0 && fitp.z == 0)) {
        float depth_rand = 0.0f, depth_pert = 0.0f;
        float4 n_rand = make_float4(0, 0, 0, 0), n_pert = n_rand, base_n = n_rand;
        float base_d = 0.0f;
#pragma unroll 1
        for (int i = -1; i < 5; ++i) {
            float4 tp;
            if (i == -1) {
                tp = fitp;
            } else {
                if (i == 0) {
                    depth_rand = rng.uniform() * (dmax - dmin) + dmin;
                    n_rand = random_normal(K, px, py, rng, depth_now);
                    const float lo = (1.0f - 0.02f) * depth_now, hi = (1.0f + 0.02f) * depth_now;
                    depth_pert = rng.uniform() * (hi - lo) + lo;
                    n_pert = perturbed_normal(K, px, py, plane_now, rng, (float)(0.02 * 3.14159265358979323846));
                    base_n = plane_now;
                    base_d = depth_now;
                }
                tp = (i == 0 || i == 4) ? base_n : (i == 3 ? n_pert : n_rand);
                const float d = (i == 0 || i == 2) ? depth_rand : (i == 4 ? depth_pert : base_d);
                tp.w = distance_to_origin(K, px, py, d, tp);
            }
            const PlaneM m = plane_row(K, tp);
            float acc = 0.0f;
            for (uint32_t mk = wmask; mk; mk &= mk - 1) {
                const int v = __ffs(mk) - 1;
                float c = ncc_new_x<U, SA>(K, s_vk[v], v, px, py, m, rp, ar, si);
                n_new++;
                if (use_geom) { c = c + K.geom_factor * geom_cost(K, s_vk[v], v, px, py, tp); n_geom++; }
                acc += (float)vw_get(w, v) * c;
            }
            const float tc = acc / wnorm;
            const float db = depth_from_plane(K, tp, px, py);
            if (db >= dmin && db <= dmax && tc < cost_now) { depth_now = db; plane_now = tp; cost_now = tc; }
        }
    }
    if (K.state == APDE_REFINE_INIT) {
        if ((double)cost_now < (double)cost_written - 0.1) { K.costs[center] = cost_now; K.planes[center] = plane_now; }
        else K.costs[center] = cost_written;
    } else {
        K.costs[center] = cost_now;
        K.planes[center] = plane_now;
    }
    count_evals(K, 0, n_new, n_geom);
}
__global__ void __launch_bounds__(128) k_prop_weak(const __grid_constant__ PassK K, int iter, int color, int tiles_x,
                                                   int ylimit) {
    if (K.tex_unorm > 0.0f) k_prop_weak_body<true, false>(K, iter, color, tiles_x, ylimit);
    else k_prop_weak_body<false, false>(K, iter, color, tiles_x, ylimit);
}
// the same kernel for a problem with a segment-label map (see "segment labels" in apde_device.cuh)
__global__ void __launch_bounds__(128) k_prop_weak_sa(const __grid_constant__ PassK K, int iter, int color, int tiles_x,
                                                      int ylimit) {
    if (K.tex_unorm > 0.0f) k_prop_weak_body<true, true>(K, iter, color, tiles_x, ylimit);
    else k_prop_weak_body<false, true>(K, iter, color, tiles_x, ylimit);
}


cudaError_t launch_stage_apd(const PassK &K, int stage, int iter, int color, cudaStream_t st) {
    const int W = K.W, H = K.H, N = K.N;
    const int tiles8x = (W + 7) / 8;
    const int tiles_full = tiles8x * ((H + 3) / 4);
    // list launches: color carries the host-side length of the longer of the two weak lists; blockIdx.y = colour
    const dim3 list_grid((unsigned)((max(color, 1) + 127) / 128), 2);
    switch (stage) {
        case APDE_STAGE_NEAREST_STRONG:
            k_ns_tiles<<<tiles8x * ((H + 7) / 8), 64, 0, st>>>(K, tiles8x);
            k_nearest_strong<<<(tiles_full + 3) / 4, 128, 0, st>>>(K, tiles8x);
            break;
        case APDE_STAGE_GEN_ANCHORS:
            if (K.list) k_gen_anchors<<<list_grid, 128, 0, st>>>(K, tiles8x);
            else k_gen_anchors<<<(tiles_full + 3) / 4, 128, 0, st>>>(K, tiles8x);
            break;
        case APDE_STAGE_RANSAC_FIT:
            if (K.list) {
                k_fit_copy_nonweak<<<(W * H + 255) / 256, 256, 0, st>>>(K);
                k_ransac_fit<<<list_grid, 128, 0, st>>>(K, iter, tiles8x);
            } else {
                k_ransac_fit<<<(tiles_full + 3) / 4, 128, 0, st>>>(K, iter, tiles8x);
            }
            break;
        case APDE_STAGE_PROP_WEAK: {
            const int ylimit = min(H, half_rows_limit(H));
            const int threads = prop_block_threads(N);
            const size_t smem = sizeof(float) * ((size_t)views_smem_floats(N) + (size_t)9 * N * threads);
            static SmemOptIn configured;
            int dev = 0;
            if (configured.needed(smem, &dev)) {
                cudaError_t e = cudaFuncSetAttribute(k_prop_weak, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e == cudaSuccess) e = cudaFuncSetAttribute(k_prop_weak_sa, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return e;
                configured.done(smem, dev);
            }
            const int tiles = half_tiles(K, ylimit);
            const int tiles8x = half_tiles_x(K);
            const int wpb = threads / 32;
            if (K.sa) k_prop_weak_sa<<<(tiles + wpb - 1) / wpb, threads, smem, st>>>(K, iter, color, tiles8x, ylimit);
            else k_prop_weak<<<(tiles + wpb - 1) / wpb, threads, smem, st>>>(K, iter, color, tiles8x, ylimit);
            break;
        }
        default:
            return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace apde
