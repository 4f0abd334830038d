// Microbenchmark: latency of one producer->consumer hand-off between SMs through global memory, for the
// signalling variants a dataflow (sync-free) triangular sweep could use.  Chain of NHOPS warps, warp k waits for
// slot k-1 then writes slot k; blocks are 1 warp, grid covers the chain round-robin over resident CTAs.
#include <cstdio>
#include <cuda_runtime.h>

#define SENT 0xFFF8B200DEADC0DEULL

template <int V>
__device__ __forceinline__ unsigned long long ld(const unsigned long long* p)
{
    unsigned long long v;
    if (V == 0) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (V == 1) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (V == 2) asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (V == 3) asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <int V>
__device__ __forceinline__ void st(unsigned long long* p, unsigned long long v)
{
    if (V == 0) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
    if (V == 1) asm volatile("st.volatile.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
    if (V == 2) asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
    if (V == 3) { asm volatile("st.global.wt.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
    if (V == 4) atomicExch(p, v);
    if (V == 5) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); __threadfence(); }
}

// one warp per block; block b handles hops b, b+G, b+2G, ...
template <int LV, int SV>
__global__ void chain(unsigned long long* slots, int nHops, int lanes)
{
    const int lane = threadIdx.x;
    for (int k = blockIdx.x; k < nHops; k += gridDim.x)
    {
        unsigned long long v = 1;
        if (k > 0 && lane < lanes)
        {
            const unsigned long long* p = slots + (size_t)(k - 1)*32 + lane;
            do { v = ld<LV>(p); } while (v == SENT);
        }
        __syncwarp();
        if (lane < lanes) st<SV>(slots + (size_t)k*32 + lane, v + 1);
    }
}

__global__ void fill(unsigned long long* p, size_t n)
{
    for (size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x*blockDim.x) p[i] = SENT;
}

template <int LV, int SV>
void run(const char* name, unsigned long long* slots, int nHops, int grid, int lanes)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++)
    {
        fill<<<1024, 256>>>(slots, (size_t)nHops*32);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        chain<LV, SV><<<grid, 32>>>(slots, nHops, lanes);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    printf("%-40s grid %4d lanes %2d : %8.3f us total, %7.1f ns/hop\n", name, grid, lanes, best*1e3, best*1e6/nHops);
}

int main()
{
    const int nHops = 4000;
    unsigned long long* slots;
    cudaMalloc(&slots, (size_t)nHops*32*8);
    for (int grid : {2, 148, 592})
    {
        for (int lanes : {1, 32})
        {
            run<0, 0>("ld.relaxed.gpu / st.relaxed.gpu", slots, nHops, grid, lanes);
            run<1, 1>("ld.volatile / st.volatile", slots, nHops, grid, lanes);
            run<2, 2>("ld.acquire.gpu / st.release.gpu", slots, nHops, grid, lanes);
            run<3, 3>("ld.cv / st.wt", slots, nHops, grid, lanes);
            run<0, 4>("ld.relaxed.gpu / atomicExch", slots, nHops, grid, lanes);
            run<0, 5>("ld.relaxed.gpu / st.relaxed+threadfence", slots, nHops, grid, lanes);
        }
    }
    // same SM: 2 warps of one block ping-pong through global memory vs through shared memory is a separate question;
    // grid 1 = every hop is produced and consumed by the same warp (no wait): the loop overhead floor.
    run<0, 0>("floor: same warp, no wait", slots, nHops, 1, 32);
    return 0;
}
