// probe_mufu_rcp.cu -- bit patterns of MUFU.RCP(n) (rcp.approx.ftz.f32, what "1.0f / x" compiles to under --use_fast_math)
// for n = 1..36, the tap counts a masked NCC patch can have.  Run on the GPU box; the table goes into oracle/apd_oracle.cpp.
//   nvcc -arch=sm_100a -o probe_mufu_rcp tools/probe_mufu_rcp.cu && ./probe_mufu_rcp
#include <cstdio>
#include <cstring>
__global__ void k(float *out) {
    const int n = threadIdx.x + 1;
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float)n));
    out[threadIdx.x] = r;
}
int main() {
    float *d, h[36];
    cudaMalloc(&d, sizeof(h));
    k<<<1, 36>>>(d);
    if (cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
    for (int i = 0; i < 36; ++i) {
        unsigned u, e;
        const float exact = 1.0f / (float)(i + 1);
        memcpy(&u, &h[i], 4);
        memcpy(&e, &exact, 4);
        printf("%d 0x%08x 0x%08x %d\n", i + 1, u, e, (int)(u - e));
    }
    return 0;
}
