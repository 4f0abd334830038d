// Microbenchmark: cost of one "round" of a single-warp dependent chain through shared memory, with and
// without a global publish of the result, to find what sits on the round-to-round critical path.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void rounds(double* out, unsigned long long* outbits, int n, double seed)
{
    __shared__ double vs[4096];
    const int lane = threadIdx.x;
    for (int i = lane; i < 4096; i += 32) vs[i] = seed + i;
    __syncwarp();
    // lane 0 carries the chain: vs[k] depends on vs[k-1]
    for (int k = 1; k < n; k++)
    {
        if (lane == 0)
        {
            const int q = k & 4095, p = (k - 1) & 4095;
            double acc = __dsub_rn(vs[q], __dmul_rn(0.5, vs[p]));
            vs[q] = acc;
            if (MODE == 1) out[k] = acc;                                                   // plain st.global
            if (MODE == 2) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(outbits + k), "l"(__double_as_longlong(acc)) : "memory");
            if (MODE == 3) asm volatile("st.volatile.global.u64 [%0], %1;" :: "l"(outbits + k), "l"(__double_as_longlong(acc)) : "memory");
            if (MODE == 4) __stcg(out + k, acc);
            if (MODE == 5) asm volatile("st.weak.global.u64 [%0], %1;" :: "l"(outbits + k), "l"(__double_as_longlong(acc)));
        }
        __syncwarp();
    }
    if (lane == 0) out[0] = vs[(n - 1) & 4095];
}

// MODE 6: the chain warp only writes shared memory + a shared progress counter; a second warp publishes.
__global__ void rounds_split(unsigned long long* outbits, int n, double seed)
{
    __shared__ double vs[4096];
    __shared__ volatile int progress;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) vs[i] = seed + i;
    if (threadIdx.x == 0) progress = 0;
    __syncthreads();
    if (warp == 0)
    {
        for (int k = 1; k < n; k++)
        {
            if (lane == 0)
            {
                const int q = k & 4095, p = (k - 1) & 4095;
                vs[q] = __dsub_rn(vs[q], __dmul_rn(0.5, vs[p]));
            }
            __syncwarp();
            if (lane == 0) progress = k;
        }
    }
    else
    {
        int done = 0;
        while (done < n - 1)
        {
            const int p = progress;
            for (int k = done + 1 + lane; k <= p; k += 32)
                asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(outbits + k), "l"(__double_as_longlong(((volatile double*)vs)[k & 4095])) : "memory");
            done = p;
        }
    }
}

int main()
{
    const int n = 20000;
    double* out; cudaMalloc(&out, sizeof(double)*n);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[] = {"no publish", "plain st.global", "st.relaxed.gpu", "st.volatile", "__stcg", "st.weak (asm, no clobber)"};
    for (int mode = 0; mode < 6; mode++)
    {
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++)
        {
            cudaEventRecord(e0);
            switch (mode)
            {
                case 0: rounds<0><<<1, 32>>>(out, (unsigned long long*)out, n, 1.0); break;
                case 1: rounds<1><<<1, 32>>>(out, (unsigned long long*)out, n, 1.0); break;
                case 2: rounds<2><<<1, 32>>>(out, (unsigned long long*)out, n, 1.0); break;
                case 3: rounds<3><<<1, 32>>>(out, (unsigned long long*)out, n, 1.0); break;
                case 4: rounds<4><<<1, 32>>>(out, (unsigned long long*)out, n, 1.0); break;
                case 5: rounds<5><<<1, 32>>>(out, (unsigned long long*)out, n, 1.0); break;
            }
            cudaEventRecord(e1); cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("%-28s %7.1f ns/round\n", names[mode], best*1e6/n);
    }
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++)
    {
        cudaEventRecord(e0);
        rounds_split<<<1, 64>>>((unsigned long long*)out, n, 1.0);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("%-28s %7.1f ns/round\n", "split: publisher warp", best*1e6/n);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
