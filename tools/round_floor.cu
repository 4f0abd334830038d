// Builds the sweep round up one ingredient at a time (all shared accesses volatile asm, one warp, 32 lanes):
// cycles per round for each stage, to see which ingredient costs what.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double ldv(unsigned a) { double v; asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ double ldw(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void stv(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" :: "r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ int4 ld4(unsigned a) { int4 v; asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ double2 ld2(unsigned a) { double2 v; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory"); return v; }

template <int S>
__global__ void __launch_bounds__(32) k(double* out, long long* cyc, int n)
{
    __shared__ __align__(16) double vals[2048 + 2];
    __shared__ __align__(16) double rowF[4*64];
    __shared__ __align__(16) int rowC[4*64];
    const int lane = threadIdx.x;
    for (int i = lane; i < 2050; i += 32) vals[i] = 1.0 + 1e-3*i;
    for (int i = lane; i < 256; i += 32) { rowF[i] = 0.01*(1 + (i & 3)); }
    __syncwarp();
    const unsigned vb = (unsigned)__cvta_generic_to_shared(vals), fb = (unsigned)__cvta_generic_to_shared(rowF), cb = (unsigned)__cvta_generic_to_shared(rowC);
    const double f0 = 0.11, f1 = 0.07, f2 = 0.05, a0 = 0.5 + lane*1e-3;
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 1; r <= n; r++)
    {
        const unsigned prev = vb + 8u*(((r - 1) & 63)*32), cur = vb + 8u*((r & 63)*32);
        double acc;
        if (S == 0) { acc = __dadd_rn(ldw(prev + 8u*((lane + 1) & 31)), 1e-9); }                                   // LDS, DADD, STS, sync
        if (S == 1) { acc = __dsub_rn(a0, __dmul_rn(f0, ldw(prev + 8u*((lane + 1) & 31)))); }                       // + DMUL
        if (S >= 2)                                                                                                // three deps, chain of three DSUB
        {
            const double v0 = ldw(prev + 8u*lane), v1 = ldw(prev + 8u*((lane + 1) & 31)), v2 = ldw(prev + 8u*((lane + 7) & 31));
            acc = __dsub_rn(__dsub_rn(__dsub_rn(a0, __dmul_rn(f0, v0)), __dmul_rn(f1, v1)), __dmul_rn(f2, v2));
            if (S == 3 && acc != acc) { acc = ldv(prev); }                                                         // + NaN test & (never taken) branch
            if (S == 6 && __any_sync(0xffffffffu, acc != acc)) { acc = ldv(prev); }                               // warp-uniform NaN test
            if (S == 7)                                                                                            // integer input tests + any_sync
            {
                const bool bad = (__double_as_longlong(v0) == 0x7ff8dead0000beefLL) | (__double_as_longlong(v1) == 0x7ff8dead0000beefLL) | (__double_as_longlong(v2) == 0x7ff8dead0000beefLL);
                if (__any_sync(0xffffffffu, bad)) { acc = ldv(prev); }
            }
            if (S == 8)                                                                                            // integer input tests + plain branch (production)
            {
                const bool bad = (__double_as_longlong(v0) == 0x7ff8dead0000beefLL) | (__double_as_longlong(v1) == 0x7ff8dead0000beefLL) | (__double_as_longlong(v2) == 0x7ff8dead0000beefLL);
                if (bad) { acc = ldv(prev); }
            }
            if (S == 9)                                                                                            // high-word test of the RESULT on the integer pipe + any_sync
            {
                if (__any_sync(0xffffffffu, ((unsigned)__double2hiint(acc) & 0x7ff80000u) == 0x7ff80000u)) { acc = ldv(prev); }
            }
            if (S == 10)                                                                                           // ballot of the NaN test, slow path after the store (speculative)
            {
                stv(cur + 8u*lane, acc);
                if (__ballot_sync(0xffffffffu, acc != acc) != 0u) { acc = ldv(prev); stv(cur + 8u*lane, acc); }
            }
            if (S == 5)                                                                                            // + static-data prefetch traffic
            {
                const int4 c = ld4(cb + 16u*(r & 63)); const double2 a = ld2(fb + 32u*(r & 63)), b = ld2(fb + 32u*(r & 63) + 16u);
                if (c.x == 0x7fffffff) acc += a.x + b.y;
            }
        }
        if (S != 10) stv(cur + 8u*lane, acc);
        if (S == 4) asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" :: "l"(out + (r & 63)*32 + lane), "d"(acc) : "memory");
        __syncwarp();
    }
    long long t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
    out[lane] = vals[lane + 32];
}
int main()
{
    double* out; long long* cyc; cudaMalloc(&out, 8*4096); cudaMalloc(&cyc, 8);
    const int n = 4096;
    const char* names[] = {"S0 LDS -> DADD -> STS -> syncwarp", "S1 LDS -> DMUL -> DSUB -> STS -> syncwarp", "S2 3x LDS -> 3 DMUL -> 3 chained DSUB -> STS -> sync",
                           "S3 = S2 + NaN test on the result", "S4 = S2 + st.relaxed.gpu publish", "S5 = S2 + static-data loads (LDS.128 x3)",
                           "S6 = S2 + any_sync(NaN result)", "S7 = S2 + input tests + any_sync", "S8 = S2 + input tests + plain branch", "S9 = S2 + hi-word result test + any_sync", "S10 = S2 + store first, ballot after"};
    void (*ks[])(double*, long long*, int) = {k<0>, k<1>, k<2>, k<3>, k<4>, k<5>, k<6>, k<7>, k<8>, k<9>, k<10>};
    for (int s = 0; s < 11; s++)
    {
        long long best = 1LL << 60;
        for (int rep = 0; rep < 3; rep++) { ks[s]<<<1, 32>>>(out, cyc, n); cudaDeviceSynchronize(); long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); if (c < best) best = c; }
        printf("%-58s %6.1f cycles/round\n", names[s], (double)best/n);
    }
    return 0;
}
