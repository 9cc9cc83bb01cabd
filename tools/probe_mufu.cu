// prints the bit patterns of MUFU.RCP for the patch sample counts (oracle constants)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(const float *in, float *out, int n) {
    int i = threadIdx.x;
    if (i < n) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(in[i])); out[i] = r; }
}
int main() {
    float h[4] = {36.0f, 9.0f, 3.0f, 7.0f}, o[4], *di, *dout;
    cudaMalloc(&di, 16); cudaMalloc(&dout, 16);
    cudaMemcpy(di, h, 16, cudaMemcpyHostToDevice);
    k<<<1, 32>>>(di, dout, 4);
    cudaMemcpy(o, dout, 16, cudaMemcpyDeviceToHost);
    for (int i = 0; i < 4; ++i) { unsigned u, e; float ex = 1.0f / h[i]; memcpy(&u, &o[i], 4); memcpy(&e, &ex, 4); printf("rcp(%g) = %.9g bits 0x%08x ; exact 1/x bits 0x%08x\n", h[i], o[i], u, e); }
    return 0;
}
