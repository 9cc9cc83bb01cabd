// apde_microbench.cu -- roofline denominators measured on the box, in the same process as the benchmark:
// FP32 FMA pipe rate and texture-unit bilinear gather rate (MEASURED_PEAKS.json only has HBM and bf16 GEMM).
#include <cuda_runtime.h>

#include "apde_kernels.h"

namespace apde {

__global__ void __launch_bounds__(256) k_fma_peak(float *out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// every thread gathers `iters` 6x6 patches of bilinear samples around its own pixel with a mild affine warp: the access
// pattern of one NCC evaluation with good locality (an upper bound for the cost kernels' gather rate)
__global__ void __launch_bounds__(128) k_tex_peak(cudaTextureObject_t tex, int layer, int W, int H, int iters, float *out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int warp = idx >> 5, lane = idx & 31;
    const int tiles_x = W / 8;
    const int tx = warp % tiles_x, ty = (warp / tiles_x) % (H / 4);
    const float px = tx * 8 + (lane & 7), py = ty * 4 + (lane >> 3);
    float acc = 0.0f;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const float ox = 3.3f + 0.37f * it, oy = 1.7f + 0.21f * it;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const float x = px + 1.01f * (2 * i - 5) + 0.02f * (2 * j - 5) + ox;
                const float y = py + 0.99f * (2 * j - 5) - 0.03f * (2 * i - 5) + oy;
                acc += tex2DLayered<float>(tex, x, y, layer);
            }
        }
    }
    out[idx] = acc;
}

cudaError_t microbench_fp32(double *tflops, cudaStream_t st) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, threads = 256, iters = 4096;
    float *out;
    cudaError_t e = cudaMalloc(&out, (size_t)blocks * threads * sizeof(float));
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, st);
        k_fma_peak<<<blocks, threads, 0, st>>>(out, iters, 1.000001f, 1e-7f);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    *tflops = (double)blocks * threads * iters * 64.0 * 2.0 / (best * 1e-3) / 1e12;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    return cudaGetLastError();
}

cudaError_t microbench_tex(cudaTextureObject_t tex, int layer, int W, int H, double *gsamples, cudaStream_t st) {
    const int warps = (W / 8) * (H / 4);
    const int threads = 128, blocks = (warps * 32) / threads, iters = 64;
    float *out;
    cudaError_t e = cudaMalloc(&out, (size_t)blocks * threads * sizeof(float));
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, st);
        k_tex_peak<<<blocks, threads, 0, st>>>(tex, layer, W, H, iters, out);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    *gsamples = (double)blocks * threads * iters * 36.0 / (best * 1e-3) / 1e9;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    return cudaGetLastError();
}

}  // namespace apde

// ------------------------------------------------------------------------------------------------------------------
// access-pattern study for the cost kernels (tools/diag_texpattern.py): how much of the 4 samples/clk/SM filter rate
// survives when the hypotheses of neighbouring pixels scatter their patches over the source image?
//   mode 0: one thread = one evaluation (36 samples), the 32 lanes of a warp are 32 neighbouring pixels
//   mode 1: one QUAD = one evaluation, each lane gathers 9 of the 36 samples (2x2 sample blocks per instruction)
// `spread` = radius in pixels of the per-evaluation random displacement (0 = coherent planes).
namespace apde {
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__global__ void __launch_bounds__(128) k_tex_pattern(cudaTextureObject_t tex, int layer, int W, int H, int iters, int mode,
                                                     float spread, float *out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int warp = idx >> 5, lane = idx & 31;
    const int tiles_x = W / 8;
    const int tx = warp % tiles_x, ty = (warp / tiles_x) % (H / 4);
    float acc = 0.0f;
    if (mode == 0) {
        const float px = tx * 8 + (lane & 7), py = ty * 4 + (lane >> 3);
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
            const uint32_t hsh = hash32((uint32_t)idx * 9781u + it * 6271u);
            const float ox = spread * ((hsh & 0xffff) * (2.0f / 65535.0f) - 1.0f), oy = spread * ((hsh >> 16) * (2.0f / 65535.0f) - 1.0f);
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = 0; j < 6; ++j)
                    acc += tex2DLayered<float>(tex, px + 1.01f * (2 * i - 5) + 0.02f * (2 * j - 5) + ox + 0.37f,
                                               py + 0.99f * (2 * j - 5) - 0.03f * (2 * i - 5) + oy + 0.21f, layer);
        }
    } else {
        // 8 evaluations per warp-iteration: quad e = lanes 4e..4e+3; pixel = warp tile pixel (it*8 + e) % 32
        const int e = lane >> 2, q = lane & 3;
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
            const int pl = (it * 8 + e) & 31;
            const float px = tx * 8 + (pl & 7), py = ty * 4 + (pl >> 3);
            const uint32_t hsh = hash32((uint32_t)(warp * 32 + pl) * 9781u + (it >> 2) * 6271u);
            const float ox = spread * ((hsh & 0xffff) * (2.0f / 65535.0f) - 1.0f), oy = spread * ((hsh >> 16) * (2.0f / 65535.0f) - 1.0f);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int i = 2 * (t / 3) + (q & 1), j = 2 * (t % 3) + (q >> 1);
                acc += tex2DLayered<float>(tex, px + 1.01f * (2 * i - 5) + 0.02f * (2 * j - 5) + ox + 0.37f,
                                           py + 0.99f * (2 * j - 5) - 0.03f * (2 * i - 5) + oy + 0.21f, layer);
            }
        }
    }
    out[idx] = acc;
}

cudaError_t microbench_tex_pattern(cudaTextureObject_t tex, int layer, int W, int H, int mode, float spread, double *gsamples,
                                   cudaStream_t st) {
    const int warps = (W / 8) * (H / 4);
    const int threads = 128, blocks = (warps * 32) / threads;
    const int iters = mode == 0 ? 32 : 128;
    float *out;
    cudaError_t e = cudaMalloc(&out, (size_t)blocks * threads * sizeof(float));
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0, st);
        k_tex_pattern<<<blocks, threads, 0, st>>>(tex, layer, W, H, iters, mode, spread, out);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    const double samples = (double)blocks * threads * iters * (mode == 0 ? 36.0 : 9.0);
    *gsamples = samples / (best * 1e-3) / 1e9;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    return cudaGetLastError();
}
}  // namespace apde
