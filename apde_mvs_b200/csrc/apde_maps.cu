// apde_maps.cu -- pyramid / map kernels either side of the pass: OpenCV-compatible INTER_LINEAR image resize,
// INTER_NEAREST map up-sampling, state import / export.  All HBM-bound streaming kernels.
#include <cuda_fp16.h>

#include "apde_kernels.h"

namespace apde {

// cv::resize(CV_32F, INTER_LINEAR) as called at APD.cpp:574, on the float-converted 8-bit image (APD.cpp:150):
//   fx = (float)((dx + 0.5) * scale_x - 0.5); sx = floor(fx); fx -= sx; clamp to [0, W-1] with fx = 0 at the borders;
//   horizontal pass then vertical pass, each  S0 * (1 - f) + S1 * f  in float.
__global__ void __launch_bounds__(256) k_resize_linear_u8(const uint8_t *__restrict__ src, int W, int H,
                                                          float *__restrict__ dst, int w, int h, double scale_x,
                                                          double scale_y) {
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
    if (dx >= w || dy >= h) return;
    if (w == W && h == H) { dst[(size_t)dy * w + dx] = (float)src[(size_t)dy * W + dx]; return; }
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = (int)floorf(fx);
    fx -= sx;
    if (sx < 0) { fx = 0.0f; sx = 0; }
    if (sx >= W - 1) { fx = 0.0f; sx = W - 1; }
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = (int)floorf(fy);
    fy -= sy;
    if (sy < 0) { fy = 0.0f; sy = 0; }
    if (sy >= H - 1) { fy = 0.0f; sy = H - 1; }
    const int sx1 = min(sx + 1, W - 1), sy1 = min(sy + 1, H - 1);
    const float a0 = 1.0f - fx, a1 = fx, b0 = 1.0f - fy, b1 = fy;
    const float r0 = __fadd_rn(__fmul_rn((float)src[(size_t)sy * W + sx], a0), __fmul_rn((float)src[(size_t)sy * W + sx1], a1));
    const float r1 = __fadd_rn(__fmul_rn((float)src[(size_t)sy1 * W + sx], a0), __fmul_rn((float)src[(size_t)sy1 * W + sx1], a1));
    dst[(size_t)dy * w + dx] = __fadd_rn(__fmul_rn(r0, b0), __fmul_rn(r1, b1));
}

cudaError_t launch_resize_linear_u8(const uint8_t *src, int W, int H, float *dst, int w, int h, cudaStream_t st) {
    const double inv_x = (double)w / W, inv_y = (double)h / H;
    dim3 grid((w + 255) / 256, h);
    k_resize_linear_u8<<<grid, 256, 0, st>>>(src, W, H, dst, w, h, 1.0 / inv_x, 1.0 / inv_y);
    return cudaGetLastError();
}

// cv::resize(INTER_NEAREST): sx = min(floor(dx * (1 / inv_scale_x)), W - 1)
template <typename T>
__global__ void __launch_bounds__(256) k_resize_nearest(const T *__restrict__ src, int W, int H, T *__restrict__ dst, int w,
                                                        int h, double ifx, double ify) {
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
    if (dx >= w || dy >= h) return;
    const int sx = min((int)floor(dx * ifx), W - 1), sy = min((int)floor(dy * ify), H - 1);
    dst[(size_t)dy * w + dx] = src[(size_t)sy * W + sx];
}
struct F3 { float x, y, z; };

cudaError_t launch_resize_nearest(const void *src, int W, int H, void *dst, int w, int h, int elem, cudaStream_t st) {
    if (w == W && h == H) return cudaMemcpyAsync(dst, src, (size_t)w * h * elem, cudaMemcpyDeviceToDevice, st);
    const double ifx = 1.0 / ((double)w / W), ify = 1.0 / ((double)h / H);
    dim3 grid((w + 255) / 256, h);
    if (elem == 1) k_resize_nearest<uint8_t><<<grid, 256, 0, st>>>((const uint8_t *)src, W, H, (uint8_t *)dst, w, h, ifx, ify);
    else if (elem == 4) k_resize_nearest<float><<<grid, 256, 0, st>>>((const float *)src, W, H, (float *)dst, w, h, ifx, ify);
    else if (elem == 12) k_resize_nearest<F3><<<grid, 256, 0, st>>>((const F3 *)src, W, H, (F3 *)dst, w, h, ifx, ify);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

// plane_hypotheses_host[center] = (normal, depth), APD.cpp:674-682
__global__ void __launch_bounds__(256) k_planes_from_maps(const float *__restrict__ depth, const float *__restrict__ normal,
                                                          float4 *__restrict__ planes, int P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    planes[i] = make_float4(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2], depth[i]);
}
cudaError_t launch_planes_from_maps(const float *depth, const float *normal, float4 *planes, int P, cudaStream_t st) {
    k_planes_from_maps<<<(P + 255) / 256, 256, 0, st>>>(depth, normal, planes, P);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_float_to_half(const float *__restrict__ src, __half *__restrict__ dst, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __float2half_rn(src[i]);
}
cudaError_t launch_float_to_half(const float *src, void *dst, size_t n, cudaStream_t st) {
    k_float_to_half<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, (__half *)dst, n);
    return cudaGetLastError();
}

// pyramid levels that are exact 2x2 means of 8-bit pixels: 4 x value is an integer <= 1020
__global__ void __launch_bounds__(256) k_float_to_u16x4(const float *__restrict__ src, unsigned short *__restrict__ dst, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (unsigned short)__float2int_rn(src[i] * 4.0f);
}
cudaError_t launch_float_to_u16x4(const float *src, void *dst, size_t n, cudaStream_t st) {
    k_float_to_u16x4<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, (unsigned short *)dst, n);
    return cudaGetLastError();
}

cudaError_t launch_fill_u8(uint8_t *dst, uint8_t v, size_t n, cudaStream_t st) { return cudaMemsetAsync(dst, v, n, st); }

// ProcessProblem tail, main.cpp:168-178
__global__ void __launch_bounds__(256) k_finish(const float4 *__restrict__ planes, const uint8_t *__restrict__ weak,
                                                float *__restrict__ depth, float *__restrict__ normal,
                                                uint8_t *__restrict__ weak_out, int P, float dmin, float dmax) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float4 pl = planes[i];
    float d = pl.w;
    uint8_t wk = weak[i];
    if (d < dmin || d > dmax) { d = 0.0f; wk = APDE_UNKNOWN; }
    depth[i] = d;
    normal[3 * i] = pl.x; normal[3 * i + 1] = pl.y; normal[3 * i + 2] = pl.z;
    weak_out[i] = wk;
}
cudaError_t launch_finish(const float4 *planes, uint8_t *weak, float *depth, float *normal, uint8_t *weak_out, int P,
                          float dmin, float dmax, cudaStream_t st) {
    k_finish<<<(P + 255) / 256, 256, 0, st>>>(planes, weak, depth, normal, weak_out, P, dmin, dmax);
    return cudaGetLastError();
}

}  // namespace apde
