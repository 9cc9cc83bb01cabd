// apde_lists.cu -- the four (colour, strong | weak) pixel lists of a problem, in TILE ORDER.  sm_100a.
//
// The red/black kernels of rounds with WEAK pixels walk compacted lists (no lane idles on the other class).  The first
// version appended tiles with one atomicAdd per (tile, class): slots came out in atomic order, so neighbouring CTAs of the
// propagation kernels worked on distant tiles and warps straddled unrelated tiles (measured on B200, r01: the strong kernel
// reached 66.5 % of the gather peak through such a list against 72.8 % on the implicit tile mapping).  Here the list is
// built in three steps -- per-tile class counts, one exclusive scan over [class][tile], fill at the scanned offsets -- so
// that list order == tile order (row-major tiles, pixels of a tile in row-major order) and the result is deterministic.
#include <cub/cub.cuh>

#include "apde_common.cuh"
#include "apde_kernels.h"

namespace apde {

__device__ __forceinline__ int list_class(const PassK &K, int tx, int ty, int shift, int ylimit, int l, int &center) {
    const int px = (tx << shift) + (l & ((1 << shift) - 1)), py = ty * (64 >> shift) + (l >> shift);
    center = py * K.W + px;
    if (px >= K.W || py >= ylimit) return -1;
    return (((px + py) & 1) << 1) | (K.weak[center] == APDE_WEAK ? 1 : 0);
}

// one warp per tile of 64 pixels: tile_counts[class][tile]
__global__ void __launch_bounds__(128) k_list_count(const __grid_constant__ PassK K, int tiles_x, int tiles, int shift, int ylimit,
                                                    int *__restrict__ tile_counts) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (blockDim.x >> 5) + warp;
    if (tile >= tiles) return;
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    int c0, c1, center;
    c0 = list_class(K, tx, ty, shift, ylimit, lane, center);
    c1 = list_class(K, tx, ty, shift, ylimit, lane + 32, center);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int n = __popc(__ballot_sync(0xffffffffu, c0 == c)) + __popc(__ballot_sync(0xffffffffu, c1 == c));
        if (lane == 0) tile_counts[(size_t)c * tiles + tile] = n;
    }
}

__global__ void __launch_bounds__(128) k_list_fill(const __grid_constant__ PassK K, int tiles_x, int tiles, int shift, int ylimit,
                                                   const int *__restrict__ offsets, int *__restrict__ lists, int *__restrict__ counts, int cap) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (blockDim.x >> 5) + warp;
    if (tile >= tiles) return;
    if (tile == 0 && lane < 4) counts[lane] = offsets[(size_t)(lane + 1) * tiles] - offsets[(size_t)lane * tiles];
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    int base[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) base[c] = offsets[(size_t)c * tiles + tile] - offsets[(size_t)c * tiles];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        int center;
        const int cls = list_class(K, tx, ty, shift, ylimit, lane + 32 * half, center);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const unsigned m = __ballot_sync(0xffffffffu, cls == c);
            if (cls == c) lists[(size_t)c * cap + base[c] + __popc(m & ((1u << lane) - 1u))] = center;
            base[c] += __popc(m);
        }
    }
}

// the first version (atomic order), kept behind APDE_LIST_ATOMIC=1 for A/B measurements
__global__ void __launch_bounds__(128) k_build_lists_atomic(const __grid_constant__ PassK K, int tiles_x, int shift, int ylimit, int *lists,
                                                            int *counts, int cap) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (blockDim.x >> 5) + warp;
    const int tx = tile % tiles_x, ty = tile / tiles_x;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        int center;
        const int cls = list_class(K, tx, ty, shift, ylimit, lane + 32 * half, center);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const unsigned m = __ballot_sync(0xffffffffu, cls == c);
            if (m == 0) continue;
            int base = 0;
            if (lane == __ffs(m) - 1) base = atomicAdd(&counts[c], __popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            if (cls == c) lists[(size_t)c * cap + base + __popc(m & ((1u << lane) - 1))] = center;
        }
    }
}

void ListScratch::release() {
    cudaFree(tile_counts); cudaFree(offsets); cudaFree(scan_tmp);
    tile_counts = offsets = nullptr; scan_tmp = nullptr; cap = 0; scan_bytes = 0;
}

// lists: 4 arrays of `cap` ints; counts: 4 ints.  One warp scans a (64 >> shift)-row x (1 << shift)-column tile of 64 pixels;
// shift = 3 is the 8x8 tile of half_pixel().
cudaError_t launch_build_lists(const PassK &K, int *lists, int *counts, int cap, ListScratch &ls, cudaStream_t st) {
    const int ylimit = min(K.H, half_rows_limit(K.H));
    static const int shift = [] { const char *e = getenv("APDE_LIST_TILE_SHIFT"); const int v = e ? atoi(e) : 3; return v < 1 || v > 6 ? 3 : v; }();
    static const bool atomic_order = [] { const char *e = getenv("APDE_LIST_ATOMIC"); return e && e[0] == '1'; }();
    const int tw = 1 << shift, th = 64 >> shift;
    const int tiles_x = (K.W + tw - 1) / tw;
    const int tiles = tiles_x * ((ylimit + th - 1) / th);
    cudaError_t e;
    if (atomic_order) {
        if ((e = cudaMemsetAsync(counts, 0, 4 * sizeof(int), st)) != cudaSuccess) return e;
        k_build_lists_atomic<<<(tiles + 3) / 4, 128, 0, st>>>(K, tiles_x, shift, ylimit, lists, counts, cap);
        return cudaGetLastError();
    }
    const size_t n = (size_t)4 * tiles + 1;
    if (n > ls.cap) {
        ls.release();
        if ((e = cudaMalloc(&ls.tile_counts, n * sizeof(int))) != cudaSuccess) return e;
        if ((e = cudaMalloc(&ls.offsets, n * sizeof(int))) != cudaSuccess) return e;
        cub::DeviceScan::ExclusiveSum(nullptr, ls.scan_bytes, ls.tile_counts, ls.offsets, (int)n);
        if ((e = cudaMalloc(&ls.scan_tmp, ls.scan_bytes)) != cudaSuccess) return e;
        ls.cap = n;
    }
    if ((e = cudaMemsetAsync(ls.tile_counts + (n - 1), 0, sizeof(int), st)) != cudaSuccess) return e;
    k_list_count<<<(tiles + 3) / 4, 128, 0, st>>>(K, tiles_x, tiles, shift, ylimit, ls.tile_counts);
    size_t bytes = ls.scan_bytes;
    if ((e = cub::DeviceScan::ExclusiveSum(ls.scan_tmp, bytes, ls.tile_counts, ls.offsets, (int)n, st)) != cudaSuccess) return e;
    k_list_fill<<<(tiles + 3) / 4, 128, 0, st>>>(K, tiles_x, tiles, shift, ylimit, ls.offsets, lists, counts, cap);
    return cudaGetLastError();
}

}  // namespace apde
