// Microbenchmark of the DIC sweep round loop in isolation: one warp walks NR rounds of 32 rows out of shared memory;
// every row depends on three rows of the previous round (the hex case).  Variants of the loop body are timed with
// clock64 to find the floor of t_round.   nvcc -arch=sm_100a -O3 -fmad=false -o sweep_loop_bench sweep_loop_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

static constexpr unsigned long long SENT = 0xFFF8B10CDEADBEEFULL;
static constexpr int RING = 64;            // rounds held in the tables
static constexpr int T = RING*32;

__device__ __forceinline__ double lds_f64(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ double lds_f64v(unsigned a) { double v; asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_f64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" :: "r"(a), "d"(v)); }
__device__ __forceinline__ int4 lds_v4(unsigned a) { int4 v; asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ double2 lds_d2(unsigned a) { double2 v; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ bool is_sent(double v) { return (unsigned long long)__double_as_longlong(v) == SENT; }

struct Row { int4 c; double f0, f1, f2, f3, acc0; unsigned q8; };

// layout: vals [T+2] | rowF [4T] | rowC [4T] | acc0 [T]
template <int V>
__global__ void __launch_bounds__(32) sweep(double* out, long long* cycles, int laps)
{
    extern __shared__ __align__(16) unsigned char sm[];
    double* vals = (double*)sm;
    double* rowF = vals + (T + 2);
    int* rowC = (int*)(rowF + 4*T);
    double* a0 = (double*)(rowC + 4*T);
    const int lane = threadIdx.x;
    for (int q = lane; q < T; q += 32)
    {
        const int r = q >> 5, l = q & 31, pr = (r + RING - 1) % RING;
        vals[q] = 1.0 + 0.001*q;
        a0[q] = 0.5 + 0.0001*q;
        rowC[4*q + 0] = 8*(pr*32 + l);
        rowC[4*q + 1] = 8*(pr*32 + (l + 1) % 32);
        rowC[4*q + 2] = 8*(pr*32 + (l + 7) % 32);
        rowC[4*q + 3] = 8*T;                     // unused slot -> constant 1.0 with coefficient 0
        rowF[4*q + 0] = 0.11; rowF[4*q + 1] = 0.07; rowF[4*q + 2] = 0.05; rowF[4*q + 3] = 0.0;
    }
    if (lane == 0) { vals[T] = 1.0; vals[T + 1] = 0.0; }
    __syncwarp();
    const unsigned sb = (unsigned)__cvta_generic_to_shared(sm);
    const unsigned vb = sb, fb = sb + 8u*(T + 2), cb = fb + 32u*T, ab = cb + 16u*T;
    const int nItems = laps*RING;
    const long long t0 = clock64();
    if (V == 0)
    {
        // current production structure: prefetch next row's static data, 4 deps, 4 sentinel tests on the inputs
        Row cur;
        { const int q = lane; cur.c = lds_v4(cb + 16u*q); double2 a = lds_d2(fb + 32u*q), b = lds_d2(fb + 32u*q + 16u); cur.f0 = a.x; cur.f1 = a.y; cur.f2 = b.x; cur.f3 = b.y; cur.acc0 = lds_f64(ab + 8u*q); cur.q8 = 8u*q; }
        for (int i = 0; i < nItems; i++)
        {
            Row nxt;
            { const int q = ((i + 1) % RING)*32 + lane; nxt.c = lds_v4(cb + 16u*q); double2 a = lds_d2(fb + 32u*q), b = lds_d2(fb + 32u*q + 16u); nxt.f0 = a.x; nxt.f1 = a.y; nxt.f2 = b.x; nxt.f3 = b.y; nxt.acc0 = lds_f64(ab + 8u*q); nxt.q8 = 8u*q; }
            double v0 = lds_f64(vb + cur.c.x), v1 = lds_f64(vb + cur.c.y), v2 = lds_f64(vb + cur.c.z), v3 = lds_f64(vb + cur.c.w);
            while (is_sent(v0) | is_sent(v1) | is_sent(v2) | is_sent(v3))
            { v0 = lds_f64v(vb + cur.c.x); v1 = lds_f64v(vb + cur.c.y); v2 = lds_f64v(vb + cur.c.z); v3 = lds_f64v(vb + cur.c.w); }
            double acc = cur.acc0;
            acc = __dsub_rn(__dsub_rn(__dsub_rn(__dsub_rn(acc, __dmul_rn(cur.f0, v0)), __dmul_rn(cur.f1, v1)), __dmul_rn(cur.f2, v2)), __dmul_rn(cur.f3, v3));
            sts_f64(vb + cur.q8, acc);
            asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" :: "l"(out + (cur.q8 >> 3)), "d"(acc) : "memory");
            __syncwarp();
            cur = nxt;
        }
    }
    if (V == 1 || V == 2 || V == 3)
    {
        // candidate: unrolled by two (no register rotation), NaN test on the RESULT instead of four tests on the
        // inputs (a sentinel input makes the result NaN; the slow path re-polls), V2: three dependency slots,
        // V3: three slots and no publish (floor)
        constexpr int ND = (V == 1) ? 4 : 3;
        auto load = [&](int item, Row& R)
        {
            const int q = (item % RING)*32 + lane;
            R.c = lds_v4(cb + 16u*q);
            const double2 a = lds_d2(fb + 32u*q), b = lds_d2(fb + 32u*q + 16u);
            R.f0 = a.x; R.f1 = a.y; R.f2 = b.x; R.f3 = b.y; R.acc0 = lds_f64(ab + 8u*q); R.q8 = 8u*q;
        };
        auto step = [&](const Row& R)
        {
            double v0 = lds_f64(vb + R.c.x), v1 = lds_f64(vb + R.c.y), v2 = lds_f64(vb + R.c.z), v3 = (ND == 4) ? lds_f64(vb + R.c.w) : 0.0;
            double acc;
            while (true)
            {
                acc = __dsub_rn(__dsub_rn(__dsub_rn(R.acc0, __dmul_rn(R.f0, v0)), __dmul_rn(R.f1, v1)), __dmul_rn(R.f2, v2));
                if (ND == 4) acc = __dsub_rn(acc, __dmul_rn(R.f3, v3));
                if (acc == acc) break;
                // rare: NaN result -> either a sentinel input (poll) or a genuine NaN (accept)
                if (!(is_sent(v0) | is_sent(v1) | is_sent(v2) | is_sent(v3))) break;
                v0 = lds_f64v(vb + R.c.x); v1 = lds_f64v(vb + R.c.y); v2 = lds_f64v(vb + R.c.z); if (ND == 4) v3 = lds_f64v(vb + R.c.w);
            }
            sts_f64(vb + R.q8, acc);
            if (V != 3) asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" :: "l"(out + (R.q8 >> 3)), "d"(acc) : "memory");
            __syncwarp();
        };
        Row A, B;
        load(0, A);
        for (int i = 0; i < nItems; i += 2)
        {
            load(i + 1, B);
            step(A);
            load(i + 2, A);
            step(B);
        }
    }
    if (V >= 4)
    {
        // V4: dependency loads issued FIRST, static data of the next item behind them (hidden under the LDS latency);
        //     bar.warp.sync directly; three slots; NaN test on the result
        // V5: V4 + publish through a predicated store issued after the shared store
        // V6: V4 without publish
        // V7: V4 but the sentinel test is a 2-instruction test of the result's high word (integer pipe)
        auto load = [&](int item, Row& R)
        {
            const int q = (item % RING)*32 + lane;
            R.c = lds_v4(cb + 16u*q);
            const double2 a = lds_d2(fb + 32u*q), b = lds_d2(fb + 32u*q + 16u);
            R.f0 = a.x; R.f1 = a.y; R.f2 = b.x; R.f3 = b.y; R.acc0 = lds_f64(ab + 8u*q); R.q8 = 8u*q;
        };
        Row A, B;
        load(0, A);
#define STEP(R, N, nextItem)                                                                                          \
        {                                                                                                             \
            double v0 = lds_f64(vb + R.c.x), v1 = lds_f64(vb + R.c.y), v2 = lds_f64(vb + R.c.z);                      \
            load(nextItem, N);                                                                                        \
            double acc;                                                                                               \
            while (true)                                                                                              \
            {                                                                                                         \
                acc = __dsub_rn(__dsub_rn(__dsub_rn(R.acc0, __dmul_rn(R.f0, v0)), __dmul_rn(R.f1, v1)), __dmul_rn(R.f2, v2)); \
                bool bad;                                                                                             \
                if (V == 7) bad = ((unsigned)__double2hiint(acc) & 0x7ff00000u) == 0x7ff00000u; else bad = (acc != acc); \
                if (!bad) break;                                                                                      \
                if (!(is_sent(v0) | is_sent(v1) | is_sent(v2))) break;                                                \
                v0 = lds_f64v(vb + R.c.x); v1 = lds_f64v(vb + R.c.y); v2 = lds_f64v(vb + R.c.z);                      \
            }                                                                                                         \
            sts_f64(vb + R.q8, acc);                                                                                  \
            if (V != 6) asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" :: "l"(out + (R.q8 >> 3)), "d"(acc) : "memory"); \
            asm volatile("bar.warp.sync 0xffffffff;" ::: "memory");                                                   \
        }
        for (int i = 0; i < nItems; i += 2)
        {
            STEP(A, B, i + 1)
            STEP(B, A, i + 2)
        }
    }
    if (V == 8 || V == 9 || V == 10)
    {
        // V8: no sentinel handling at all (floor of the structure);  V9: warp-uniform slow path (__any_sync on the NaN
        // test: the warp never diverges);  V10: V9 + global publish
        auto load = [&](int item, Row& R)
        {
            const int q = (item % RING)*32 + lane;
            R.c = lds_v4(cb + 16u*q);
            const double2 a = lds_d2(fb + 32u*q), b = lds_d2(fb + 32u*q + 16u);
            R.f0 = a.x; R.f1 = a.y; R.f2 = b.x; R.f3 = b.y; R.acc0 = lds_f64(ab + 8u*q); R.q8 = 8u*q;
        };
        Row A, B;
        load(0, A);
#define STEP2(R, N, nextItem)                                                                                         \
        {                                                                                                             \
            double v0 = lds_f64(vb + R.c.x), v1 = lds_f64(vb + R.c.y), v2 = lds_f64(vb + R.c.z);                      \
            load(nextItem, N);                                                                                        \
            double acc = __dsub_rn(__dsub_rn(__dsub_rn(R.acc0, __dmul_rn(R.f0, v0)), __dmul_rn(R.f1, v1)), __dmul_rn(R.f2, v2)); \
            if (V != 8)                                                                                               \
            {                                                                                                         \
                while (__any_sync(0xffffffffu, (acc != acc) && (is_sent(v0) | is_sent(v1) | is_sent(v2))))            \
                {                                                                                                     \
                    v0 = lds_f64v(vb + R.c.x); v1 = lds_f64v(vb + R.c.y); v2 = lds_f64v(vb + R.c.z);                  \
                    acc = __dsub_rn(__dsub_rn(__dsub_rn(R.acc0, __dmul_rn(R.f0, v0)), __dmul_rn(R.f1, v1)), __dmul_rn(R.f2, v2)); \
                }                                                                                                     \
            }                                                                                                         \
            sts_f64(vb + R.q8, acc);                                                                                  \
            if (V == 10) asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" :: "l"(out + (R.q8 >> 3)), "d"(acc) : "memory"); \
            __syncwarp();                                                                                             \
        }
        for (int i = 0; i < nItems; i += 2)
        {
            STEP2(A, B, i + 1)
            STEP2(B, A, i + 2)
        }
    }
    const long long t1 = clock64();
    if (lane == 0) { cycles[0] = t1 - t0; out[T] = vals[5]; }
}

int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double)*(T + 8)); cudaMalloc(&cyc, 8);
    const size_t smem = 8*(T + 2) + 32*T + 16*T + 8*T;
    const int laps = 64;
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    auto run = [&](auto kern, const char* name)
    {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        long long best = 1LL << 60;
        for (int rep = 0; rep < 3; rep++)
        {
            kern<<<1, 32, smem>>>(out, cyc, laps);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            if (c < best) best = c;
        }
        printf("%-58s %6.1f cycles/round  (%.1f ns at %d MHz)\n", name, (double)best/(laps*RING), (double)best/(laps*RING)/(clk*1e-6), clk/1000);
    };
    run(sweep<0>, "V0 production structure (4 deps, input tests, rotation)");
    run(sweep<1>, "V1 unroll x2, NaN test on result, 4 deps");
    run(sweep<2>, "V2 unroll x2, NaN test on result, 3 deps");
    run(sweep<3>, "V3 = V2 without the global publish");
    run(sweep<4>, "V4 deps first, prefetch behind, bar.warp.sync, 3 deps");
    run(sweep<6>, "V6 = V4 without the global publish");
    run(sweep<7>, "V7 = V4 with integer NaN test");
    run(sweep<8>, "V8 no sentinel handling, no publish (floor)");
    run(sweep<9>, "V9 warp-uniform slow path (__any_sync), no publish");
    run(sweep<10>, "V10 = V9 + global publish");
    return 0;
}
