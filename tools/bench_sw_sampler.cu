// bench_sw_sampler.cu -- round-2 design probe: is a second sampling path beside the texture unit worth it?
//
// Standalone microbenchmark, NOT part of libapde.  It reproduces the access pattern of the DepthToWeak sweep columns -- a CTA of
// 128 threads = one 32x4 tile of reference pixels, every thread walks 61 hypotheses x 36 taps (11x11 patch, stride 2) through a
// plane-induced homography whose translation part grows with the hypothesis index -- and forms every bilinear sample in four ways:
//   TEX     tex2DLayered on an 8-bit UNORM layered texture + the 2-FMA decode of apde_device.cuh (what the product does today)
//   SMEM8   the CTA stages the bounding box of ALL its samples (4 corner taps x 2 extreme hypotheses per thread, block min/max)
//           from a pitch-linear u8 copy into shared memory, then 4 byte loads + the texture unit's integer-weight filter
//           (profiles/r01_texture_filter_model.md: A, B = 1.8 fixed-point fractions, W11 = (A*B + 128) >> 8, ...)
//   SMEM16  the same with (texel, right neighbour) pairs stored as 16-bit words: 2 loads per sample
//   HYBRID  taps with even index through TEX, odd taps through SMEM16 (both pipes busy)
//   SMEM16F / HYBRIDF  (added after the first measurement) the software filter with ~half the instructions (fixed-point
//           coordinate from one F2I per axis, DP4A over packed texels and weights); HYBRIDF sends one tap in four through it
// and reports Gsamples/s per variant plus the number of samples whose value differs from TEX (must be 0: the software filter is
// the pinned model of the texture unit).  The accumulation per sample (sum, sum of squares, product with a reference value) is
// the NCC's, so the instruction mix around the sample is the real one.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o bench_sw_sampler tools/bench_sw_sampler.cu
//   ./bench_sw_sampler [W H layers window_KB]        (default 1920 1080 11 16; the window size bounds the CTAs per SM)
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kHyp = 61, kTaps = 6;       // 61 hypotheses, 6 x 6 taps
enum Mode { TEX = 0, SMEM8 = 1, SMEM16 = 2, HYBRID = 3, SMEM16F = 4, HYBRIDF = 5 };

struct Warp {  // x' = (h0 x + h1 y + h2 + k b0) / (h6 x + h7 y + h8 + k b2), y' alike: a homography whose translation moves with k
    float h[9];
    float b[3];
};

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void project(const Warp &w, float x, float y, float k, float &X, float &Y) {
    const float Z = fmaf(w.h[6], x, fmaf(w.h[7], y, fmaf(w.b[2], k, w.h[8])));
    const float iz = rcp_approx(Z);
    X = fmaf(w.h[0], x, fmaf(w.h[1], y, fmaf(w.b[0], k, w.h[2]))) * iz + 0.5f;
    Y = fmaf(w.h[3], x, fmaf(w.h[4], y, fmaf(w.b[1], k, w.h[5]))) * iz + 0.5f;
}
// texture path: 8-bit UNORM texel -> exact multiple of 1/256 of a grey level (apde_device.cuh: fetch<true>)
__device__ __forceinline__ float sample_tex(cudaTextureObject_t tex, float x, float y, int layer) {
    const float t = tex2DLayered<float>(tex, x, y, layer);
    const float y2 = fmaf(t, 255.0f * 256.0f, 12582912.0f);
    return fmaf(y2, 1.0f / 256.0f, -12582912.0f / 256.0f);
}
// the texture unit's filter in software; s = window (clamping baked in when it was filled), origin (x0, y0), pitch in elements
template <bool PAIR>
__device__ __forceinline__ float sample_smem(const void *win, int x0, int y0, int pitch, float x, float y) {
    const float xB = x - 0.5f, yB = y - 0.5f;
    const float fi = floorf(xB), fj = floorf(yB);
    const int A = (int)floorf(fmaf(xB - fi, 256.0f, 0.5f)), B = (int)floorf(fmaf(yB - fj, 256.0f, 0.5f));
    const int p = ((int)fj - y0) * pitch + ((int)fi - x0);
    int t00, t10, t01, t11;
    if (PAIR) {
        const unsigned short *s = (const unsigned short *)win;
        const unsigned a = s[p], b = s[p + pitch];
        t00 = a & 255; t10 = a >> 8; t01 = b & 255; t11 = b >> 8;
    } else {
        const uint8_t *s = (const uint8_t *)win;
        t00 = s[p]; t10 = s[p + 1]; t01 = s[p + pitch]; t11 = s[p + pitch + 1];
    }
    const int W11 = (A * B + 128) >> 8, W10 = A - W11, W01 = B - W11, W00 = 256 - A - B + W11;
    return (float)(W00 * t00 + W10 * t10 + W01 * t01 + W11 * t11) * (1.0f / 256.0f);
}

// the same filter with fewer instructions (CPU-checked against the pinned model on 60 000 probes incl. bucket edges):
//   one FFMA + one F2I.FLOOR per axis gives the 24.8 fixed-point coordinate fx = floor((x - 0.5) * 256 + 0.5); i = fx >> 8, A = fx & 255
//   (the model's A = 256 case becomes (i + 1, 0): the same four-weight sum);
//   the (texel, right neighbour) pairs of the two rows form one 32-bit word and the four weights another: one DP4A, plus
//   t00 * 256 in the single case W00 == 256 (A == B == 0), whose low byte is 0
__device__ __forceinline__ float sample_smem_fast(const unsigned short *s, int x0, int y0, int pitch, float x, float y) {
    const int fx = __float2int_rd(fmaf(x, 256.0f, -127.5f)), fy = __float2int_rd(fmaf(y, 256.0f, -127.5f));
    const int A = fx & 255, B = fy & 255;
    const int p = ((fy >> 8) - y0) * pitch + ((fx >> 8) - x0);
    const unsigned tex4 = __byte_perm((unsigned)s[p], (unsigned)s[p + pitch], 0x5410);  // t00 | t10 << 8 | t01 << 16 | t11 << 24
    const int W11 = (A * B + 128) >> 8, W10 = A - W11, W01 = B - W11, W00 = 256 - A - B + W11;
    const unsigned w4 = (unsigned)(W00 & 255) | ((unsigned)W10 << 8) | ((unsigned)W01 << 16) | ((unsigned)W11 << 24);
    const unsigned sum = __dp4a(tex4, w4, (unsigned)((W00 >> 8) * ((tex4 & 255u) << 8)));
    return (float)sum * (1.0f / 256.0f);
}

template <int MODE>
__global__ void __launch_bounds__(128) k_sweep(cudaTextureObject_t tex, const uint8_t *__restrict__ lin, int W, int H, int layer, Warp w,
                                               float *__restrict__ out, int *__restrict__ fallback, int win_bytes) {
    extern __shared__ unsigned char smem[];
    __shared__ int box[4];  // min x, min y, max x, max y (texel indices)
    const int tiles_x = W / 32;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int px = tx * 32 + (threadIdx.x & 31), py = ty * 4 + (threadIdx.x >> 5);
    int x0 = 0, y0 = 0, pitch = 0;
    bool staged = false;
    if (MODE != TEX) {
        if (threadIdx.x == 0) { box[0] = box[1] = 1 << 30; box[2] = box[3] = -(1 << 30); }
        __syncthreads();
        // bounding box of all samples of this thread: for a fixed tap the sample moves monotonically along a line with k, for a fixed
        // k the taps map projectively (denominator > 0): extremes at the 4 corner taps x the 2 extreme hypotheses
        int lo_x = 1 << 30, lo_y = 1 << 30, hi_x = -(1 << 30), hi_y = -(1 << 30);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float X, Y;
            project(w, (float)(px + ((c & 1) ? 5 : -5)), (float)(py + ((c & 2) ? 5 : -5)), (c & 4) ? 30.0f : -30.0f, X, Y);
            const int i = (int)floorf(X - 0.5f), j = (int)floorf(Y - 0.5f);
            lo_x = min(lo_x, i); hi_x = max(hi_x, i + 1); lo_y = min(lo_y, j); hi_y = max(hi_y, j + 1);
        }
        atomicMin(&box[0], lo_x); atomicMin(&box[1], lo_y); atomicMax(&box[2], hi_x); atomicMax(&box[3], hi_y);
        __syncthreads();
        x0 = box[0] - 1; y0 = box[1] - 1;  // one texel of slack for rounding at the box edge
        const int ww = box[2] - x0 + 2, wh = box[3] - y0 + 2;
        pitch = (ww + 3) & ~3;
        const int elem = (MODE == SMEM8) ? 1 : 2;  // every other software mode stores pairs
        staged = (long long)pitch * wh * elem <= win_bytes && ww > 0 && wh > 0;
        if (staged) {
            for (int idx = threadIdx.x; idx < pitch * wh; idx += blockDim.x) {
                const int r = idx / pitch, c = idx - r * pitch;
                const int sy = min(max(y0 + r, 0), H - 1), sx = min(max(x0 + c, 0), W - 1), sx1 = min(max(x0 + c + 1, 0), W - 1);
                const uint8_t *row = lin + ((size_t)layer * H + sy) * W;
                if (MODE == SMEM8) smem[idx] = row[sx];
                else ((unsigned short *)smem)[idx] = (unsigned short)(row[sx] | (row[sx1] << 8));
            }
        } else if (threadIdx.x == 0) {
            atomicAdd(fallback, 1);
        }
        __syncthreads();
    }
    float acc = 0.0f;
#pragma unroll 1
    for (int k = -30; k <= 30; ++k) {
        float s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
        for (int i = 0; i < kTaps; ++i) {
#pragma unroll
            for (int j = 0; j < kTaps; ++j) {
                float X, Y;
                project(w, (float)(px + 2 * i - 5), (float)(py + 2 * j - 5), (float)k, X, Y);
                float s;
                // HYBRID: every other tap in software; HYBRIDF: one tap in four (the share at which issue and filter slots balance on paper)
                const bool soft = staged && (MODE == SMEM8 || MODE == SMEM16 || MODE == SMEM16F || (MODE == HYBRID && ((i * kTaps + j) & 1)) ||
                                             (MODE == HYBRIDF && ((i * kTaps + j) & 3) == 3));
                if (MODE == TEX || !soft) s = sample_tex(tex, X, Y, layer);
                else if (MODE == SMEM16F || MODE == HYBRIDF) s = sample_smem_fast((const unsigned short *)smem, x0, y0, pitch, X, Y);
                else s = (MODE == SMEM8) ? sample_smem<false>(smem, x0, y0, pitch, X, Y) : sample_smem<true>(smem, x0, y0, pitch, X, Y);
                s1 += s;
                s2 = fmaf(s, s, s2);
                s3 = fmaf((float)(i * 7 + j), s, s3);
            }
        }
        acc += s1 * (1.0f / 36.0f) + s2 * 1e-6f + s3 * 1e-4f;
    }
    out[(size_t)py * W + px] = acc;
}

// exhaustive value check: one thread per probe, TEX against the software filter on the pitch-linear copy (no staging)
__global__ void k_check(cudaTextureObject_t tex, const uint8_t *__restrict__ lin, int W, int H, int layer, int n, unsigned seed, int *bad) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    unsigned s = seed + 747796405u * (unsigned)t;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (s >> 8) * (1.0f / 16777216.0f); };
    const float x = rnd() * (W + 8) - 4.0f, y = rnd() * (H + 8) - 4.0f;  // includes coordinates outside the image (clamp addressing)
    const float a = sample_tex(tex, x, y, layer);
    const float xB = fminf(fmaxf(x - 0.5f, -1.0f), (float)W), yB = fminf(fmaxf(y - 0.5f, -1.0f), (float)H);
    const float fi = floorf(xB), fj = floorf(yB);
    const int A = (int)floorf(fmaf(xB - fi, 256.0f, 0.5f)), B = (int)floorf(fmaf(yB - fj, 256.0f, 0.5f));
    const int i0 = min(max((int)fi, 0), W - 1), i1 = min(max((int)fi + 1, 0), W - 1), j0 = min(max((int)fj, 0), H - 1), j1 = min(max((int)fj + 1, 0), H - 1);
    const uint8_t *img = lin + (size_t)layer * H * W;
    const int W11 = (A * B + 128) >> 8, W10 = A - W11, W01 = B - W11, W00 = 256 - A - B + W11;
    const float b = (float)(W00 * img[j0 * W + i0] + W10 * img[j0 * W + i1] + W01 * img[j1 * W + i0] + W11 * img[j1 * W + i1]) * (1.0f / 256.0f);
    if (a != b) atomicAdd(bad, 1);
}

template <int MODE>
static float run(cudaTextureObject_t tex, const uint8_t *lin, int W, int H, int layers, const Warp &w, float *out, int *fallback, size_t smem) {
    CK(cudaFuncSetAttribute(k_sweep<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (W / 32) * (H / 4);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_sweep<MODE><<<grid, 128, smem>>>(tex, lin, W, H, 0, w, out, fallback, (int)smem);  // warm-up
    CK(cudaMemset(fallback, 0, sizeof(int)));
    CK(cudaEventRecord(e0));
    for (int l = 0; l < layers; ++l) k_sweep<MODE><<<grid, 128, smem>>>(tex, lin, W, H, l, w, out + (size_t)l * W * H, fallback, (int)smem);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms;
}

int main(int argc, char **argv) {
    const int W = argc > 1 ? atoi(argv[1]) : 1920, H = (argc > 2 ? atoi(argv[2]) : 1080) / 4 * 4, L = argc > 3 ? atoi(argv[3]) : 11;
    const size_t kMaxWin = (size_t)(argc > 4 ? atoi(argv[4]) : 16) * 1024;  // shared memory for the staged window, per CTA
    if (W % 32) { printf("W must be a multiple of 32\n"); return 1; }
    std::vector<uint8_t> img((size_t)L * W * H);
    unsigned s = 12345u;
    for (size_t i = 0; i < img.size(); ++i) { s = s * 1664525u + 1013904223u; img[i] = (uint8_t)(((i % W) * 3 + (i / W) * 5 + (s >> 27)) & 255); }
    uint8_t *lin;
    CK(cudaMalloc(&lin, img.size()));
    CK(cudaMemcpy(lin, img.data(), img.size(), cudaMemcpyHostToDevice));
    cudaArray_t arr;
    cudaChannelFormatDesc cd = cudaCreateChannelDesc(8, 0, 0, 0, cudaChannelFormatKindUnsigned);
    CK(cudaMalloc3DArray(&arr, &cd, make_cudaExtent(W, H, L), cudaArrayLayered));
    cudaMemcpy3DParms cp = {};
    cp.srcPtr = make_cudaPitchedPtr(lin, W, W, H);
    cp.dstArray = arr;
    cp.extent = make_cudaExtent(W, H, L);
    cp.kind = cudaMemcpyDeviceToDevice;
    CK(cudaMemcpy3D(&cp));
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = arr;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 0;
    cudaTextureObject_t tex;
    CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));

    int *d_int;
    CK(cudaMalloc(&d_int, 2 * sizeof(int)));
    CK(cudaMemset(d_int, 0, 2 * sizeof(int)));
    const int nprobe = 1 << 22;
    k_check<<<(nprobe + 255) / 256, 256>>>(tex, lin, W, H, L / 2, nprobe, 99u, d_int + 1);
    int h_int[2];
    CK(cudaMemcpy(h_int, d_int, sizeof(h_int), cudaMemcpyDeviceToHost));
    printf("software filter vs texture unit on %d random probes (incl. out-of-image coordinates): %d mismatches\n", nprobe, h_int[1]);

    // a mild homography: 3 % scale, 1.5 degree shear, weak perspective; hypothesis step ~ 0.9 px in x, 0.05 px in y
    Warp w = {{1.03f, 0.02f, 3.7f, -0.015f, 0.98f, -2.3f, 1.0e-5f, -0.8e-5f, 1.0f}, {0.9f, 0.05f, 1.0e-4f}};
    std::vector<float *> outs(6);
    for (auto &o : outs) CK(cudaMalloc(&o, (size_t)L * W * H * sizeof(float)));
    const double samples = (double)L * W * H * kHyp * kTaps * kTaps;
    const char *names[6] = {"TEX", "SMEM8", "SMEM16", "HYBRID", "SMEM16F", "HYBRIDF"};
    float ms[6];
    int fb[6] = {0, 0, 0, 0, 0, 0};
    ms[0] = run<TEX>(tex, lin, W, H, L, w, outs[0], d_int, 0);
    ms[1] = run<SMEM8>(tex, lin, W, H, L, w, outs[1], d_int, kMaxWin);
    CK(cudaMemcpy(&fb[1], d_int, sizeof(int), cudaMemcpyDeviceToHost));
    ms[2] = run<SMEM16>(tex, lin, W, H, L, w, outs[2], d_int, kMaxWin);
    CK(cudaMemcpy(&fb[2], d_int, sizeof(int), cudaMemcpyDeviceToHost));
    ms[3] = run<HYBRID>(tex, lin, W, H, L, w, outs[3], d_int, kMaxWin);
    CK(cudaMemcpy(&fb[3], d_int, sizeof(int), cudaMemcpyDeviceToHost));
    ms[4] = run<SMEM16F>(tex, lin, W, H, L, w, outs[4], d_int, kMaxWin);
    CK(cudaMemcpy(&fb[4], d_int, sizeof(int), cudaMemcpyDeviceToHost));
    ms[5] = run<HYBRIDF>(tex, lin, W, H, L, w, outs[5], d_int, kMaxWin);
    CK(cudaMemcpy(&fb[5], d_int, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<float> ref((size_t)L * W * H), got(ref.size());
    CK(cudaMemcpy(ref.data(), outs[0], ref.size() * sizeof(float), cudaMemcpyDeviceToHost));
    for (int m = 0; m < 6; ++m) {
        size_t diff = 0;
        if (m) {
            CK(cudaMemcpy(got.data(), outs[m], got.size() * sizeof(float), cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < ref.size(); ++i) diff += got[i] != ref[i];
        }
        printf("%-7s %8.2f ms  %7.1f Gsamples/s  x%.2f vs TEX  pixels differing from TEX: %zu  CTAs that fell back to TEX: %d of %d\n", names[m], ms[m],
               samples / ms[m] * 1e-6, ms[0] / ms[m], diff, fb[m], L * (W / 32) * (H / 4));
    }
    return 0;
}
