// Dependent-issue latencies on B200 of the instructions on the sweep's critical chain.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, long long* cyc, double a, double b)
{
    __shared__ double sh[64];
    __shared__ int chase[64];
    const int lane = threadIdx.x;
    sh[lane] = a + lane; sh[lane + 32] = b;
    chase[lane] = (lane + 1) & 31; chase[lane + 32] = lane;
    __syncwarp();
    double x = a;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) { x = __dadd_rn(x, b); x = __dadd_rn(x, b); x = __dadd_rn(x, b); x = __dadd_rn(x, b); }
    long long t1 = clock64();
    if (lane == 0) cyc[0] = (t1 - t0);
    double y = a;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) { y = __dmul_rn(y, b); y = __dmul_rn(y, b); y = __dmul_rn(y, b); y = __dmul_rn(y, b); }
    t1 = clock64();
    if (lane == 0) cyc[1] = (t1 - t0);
    int p = lane;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) { p = chase[p]; p = chase[p]; p = chase[p]; p = chase[p]; }
    t1 = clock64();
    if (lane == 0) cyc[2] = (t1 - t0);
    // STS -> syncwarp -> LDS round trip through another lane
    double z = a;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; i++)
    {
        ((volatile double*)sh)[lane] = z;
        __syncwarp();
        z = ((volatile double*)sh)[(lane + 1) & 31];
        __syncwarp();
    }
    t1 = clock64();
    if (lane == 0) cyc[3] = (t1 - t0);
    float f = (float)a;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) { f = __fadd_rn(f, 1.5f); f = __fadd_rn(f, 1.5f); f = __fadd_rn(f, 1.5f); f = __fadd_rn(f, 1.5f); }
    t1 = clock64();
    if (lane == 0) cyc[4] = (t1 - t0);
    // DFMA chain
    double w = a;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) { w = __fma_rn(w, b, a); w = __fma_rn(w, b, a); w = __fma_rn(w, b, a); w = __fma_rn(w, b, a); }
    t1 = clock64();
    if (lane == 0) cyc[5] = (t1 - t0);
    out[lane] = x + y + p + z + f + w;
}
int main()
{
    double* out; long long* cyc; cudaMalloc(&out, 8*32); cudaMalloc(&cyc, 8*8);
    for (int warps = 1; warps <= 1; warps++)
    {
        lat<<<1, 32>>>(out, cyc, 1.0, 1.0000001);
        cudaDeviceSynchronize();
        long long c[6]; cudaMemcpy(c, cyc, 48, cudaMemcpyDeviceToHost);
        printf("DADD dependent latency   %.1f cycles\n", c[0]/1024.0);
        printf("DMUL dependent latency   %.1f cycles\n", c[1]/1024.0);
        printf("LDS.32 pointer chase     %.1f cycles\n", c[2]/1024.0);
        printf("STS->syncwarp->LDS->syncwarp %.1f cycles\n", c[3]/1024.0);
        printf("FADD dependent latency   %.1f cycles\n", c[4]/1024.0);
        printf("DFMA dependent latency   %.1f cycles\n", c[5]/1024.0);
    }
    return 0;
}
