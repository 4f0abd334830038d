// kernels_pencil.cuh -- K9 / K10 on a structured block: the DIC factor and the triangular sweeps
// ([EXT] DICPreconditioner::calcReciprocalD / precondition) as STREAMING PENCIL sweeps.
//
// The sweeps are a dependency chain nx+ny+nz-2 levels long; what a sweep costs is (levels) x (latency of one level).  The
// tile kernels (kernels.cuh / kernels_vec.cuh) pay a shared-memory round trip + __syncwarp per level and a hop through
// L2 every 8 cells.  On a lexicographically numbered block (every synthetic cantilever, every blockMesh block, every
// `simple` sub-domain of one) the dependencies are the three lower (upper) lattice neighbours, so a warp can carry a
// whole PENCIL -- the full length n0 of the longest axis x a w1 x w2 <= 32 cross-section -- in a skewed march: at step
// t lane (p1, p2) holds cell q = t - p1 - p2 of its lattice line, and every dependency of step t was produced in step
// t - 1: by the lane itself (long axis), by lane - 1 (axis a1) or by lane - w1 (axis a2).  One level costs two shuffles
// and the DMUL / DSUB chain of the reference's own arithmetic; nothing inside the pencil goes through memory.  Only the
// two faces of the cross-section that touch other pencils go through L2 (value-is-the-flag polling, as in the tile
// kernels), and a pencil is n0 cells long: a dependency path crosses ny/w1 + nz/w2 pencil faces instead of
// nx/8 + ny/4 + nz/8 brick faces.
//
// One CTA per pencil (handed out in wavefront order by the never-reset ticket, so a CTA only waits on pencils that
// are running or done):
//   SW sweep warps   registers + shuffles, per-row data out of a shared-memory ring, results published with one store
//                    per lane -- that store is also the hand-off to other pencils.  SW = 1: one warp carries all NB
//                    components (256-bit publish); SW = NB: one warp per component, each on its own scheduler -- a sweep
//                    warp is issue-bound (one in-order instruction stream), so the three independent systems of a vector
//                    solve advance at the pace of ONE (profiles/r2_pencil_sweeps.md)
//   1 loader warp    lane 0 keeps the ring full with 1-D bulk copies (cp.async.bulk, one mbarrier per stage)
//   2 fetch warps    poll the neighbour pencils' published rows several steps ahead and drop them into the ring of
//                    external values the sweep warps read
// Rows live in the pencil numbering of plan.h: row = base + 32*step + lane, so a stage of SS steps is one contiguous
// range of every array.  Arithmetic per row = that of the tile kernels = the reference's: products rD*upper first,
// subtractions in the order of the face loop (lower neighbour along z, then y, then x; backward: upper neighbour
// along z, y, x), __d*_rn everywhere.  Rows that do not exist (the skew's lead-in / lead-out, partial pencils) carry
// zero coefficients and compute exact zeros (ones in the factor), which is also what a missing neighbour contributes.
#pragma once
#include "kernels_vec.cuh"

namespace gpupcg
{

constexpr int PEN_TPB = 128;
constexpr int PEN_SS = 4;           // steps per ring stage (the plan pads a pencil to a multiple of it)
constexpr int PEN_R = 32;           // steps held by the ring of external values (power of two, multiple of PEN_SS)
constexpr int PEN_FU = 4;           // polls a fetch lane keeps in flight
#ifndef PEN_FW
#define PEN_FW 2                    // fetch warps per CTA
#endif

struct PencilArgs
{
    const int* meta;                // 8 ints per pencil: base row, actual w1, w2, tile of the a1-, a2-, a1+, a2+ neighbour
    int nPencils, w1, w2, n0, steps;
    int nStages;                    // ring stages that fit (host: shared memory / stage bytes, <= 8)
    int fetchWindow;                // polls in flight per fetch lane: 1 .. PEN_FU fixed, <= 0 adaptive
    int fetchSleep;                 // ns a fetch lane backs off after a round without a delivery
    unsigned int* smCounter;        // [1024] per-SM CTA counter (role rotation), may be null
    const double* coefF;            // [nRows][3] upper of the faces to the lower neighbours along z, y, x (0: none)
    const double* coefB;            // [nRows][3] ... to the upper neighbours
    // forward sweep with the pending x/r update of PCG riding along (FUSE): x += alpha pA; rA -= alpha wA; sum|rA|; wA re-armed
    double* fxX; const double* fxPA; unsigned long long* fxWA; double* fxRA; double* fxHistory; int fxHistStride;
};

template <int NB> struct PenRow { static constexpr int NS = (NB == 3) ? 4 : NB; static constexpr int BYTES = 8*NS; };

// bytes of one ring stage: coefficients + (rD) + start operand (+ rA for the fused dot)
template <int NB, int MODE>
__host__ __device__ constexpr int pen_stage_bytes(bool withDot, bool fuse = false)
{
    return PEN_SS*32*(24 + PenRow<NB>::BYTES*((MODE == 1) ? 1 : (MODE == 2 && withDot) ? 3 : (MODE == 0 && fuse) ? 5 : 2));
}
template <int NB>
__host__ __device__ constexpr int pen_ext_bytes(int w1, int w2) { return PEN_R*(w1 + w2)*PenRow<NB>::BYTES; }
template <int NB, int MODE>
__host__ inline size_t pen_smem_bytes(int nStages, int w1, int w2, bool withDot, bool fuse = false)
{
    return (size_t)nStages*pen_stage_bytes<NB, MODE>(withDot, fuse) + pen_ext_bytes<NB>(w1, w2) + 8*8 + 64 + 8*8;
}

// ---- NB values of a row: shared / global, plain / volatile / published -------------------------------------------
template <int NB>
__device__ __forceinline__ void pen_lds(unsigned a, double (&v)[NB])
{
    if (NB == 1) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v[0]) : "r"(a));
    else if (NB == 2) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[0]), "=d"(v[NB - 1]) : "r"(a));
    else
    {
        double pad;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[0]), "=d"(v[1 % NB]) : "r"(a));
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[NB - 1]), "=d"(pad) : "r"(a + 16u));
    }
}
template <int NB>
__device__ __forceinline__ void pen_lds_volatile(unsigned a, double (&v)[NB])
{
    if (NB == 1) asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v[0]) : "r"(a) : "memory");
    else if (NB == 2) asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[0]), "=d"(v[NB - 1]) : "r"(a) : "memory");
    else
    {
        double pad;
        asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[0]), "=d"(v[1 % NB]) : "r"(a) : "memory");
        asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[NB - 1]), "=d"(pad) : "r"(a + 16u) : "memory");
    }
}
template <int NB>
__device__ __forceinline__ void pen_sts_bits(unsigned a, const unsigned long long (&w)[NB])
{
    if (NB == 1) asm volatile("st.volatile.shared.b64 [%0], %1;" :: "r"(a), "l"(w[0]) : "memory");
    else if (NB == 2) asm volatile("st.volatile.shared.v2.b64 [%0], {%1, %2};" :: "r"(a), "l"(w[0]), "l"(w[NB - 1]) : "memory");
    else
    {
        asm volatile("st.volatile.shared.v2.b64 [%0], {%1, %2};" :: "r"(a), "l"(w[0]), "l"(w[1 % NB]) : "memory");
        asm volatile("st.volatile.shared.b64 [%0], %1;" :: "r"(a + 16u), "l"(w[NB - 1]) : "memory");
    }
}
template <int NB>
__device__ __forceinline__ void pen_publish(unsigned long long* p, const double (&v)[NB])
{
    if (NB == 1) asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" :: "l"(p), "l"(__double_as_longlong(v[0])) : "memory");
    else if (NB == 2)
        asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};"
                     :: "l"(p), "l"(__double_as_longlong(v[0])), "l"(__double_as_longlong(v[NB - 1])) : "memory");
    else
        asm volatile("st.relaxed.gpu.global.v4.b64 [%0], {%1, %2, %3, %4};"
                     :: "l"(p), "l"(__double_as_longlong(v[0])), "l"(__double_as_longlong(v[1 % NB])),
                        "l"(__double_as_longlong(v[NB - 1])), "l"(0LL) : "memory");
}
template <int NB>
__device__ __forceinline__ void pen_peek(const unsigned long long* p, unsigned long long (&w)[NB])
{
    if (NB == 1) asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(w[0]) : "l"(p) : "memory");
    else if (NB == 2) asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(w[0]), "=l"(w[NB - 1]) : "l"(p) : "memory");
    else
    {
        unsigned long long pad;
        asm volatile("ld.relaxed.gpu.global.v4.b64 {%0, %1, %2, %3}, [%4];"
                     : "=l"(w[0]), "=l"(w[1 % NB]), "=l"(w[NB - 1]), "=l"(pad) : "l"(p) : "memory");
    }
}
template <int NB>
__device__ __forceinline__ void pen_store_plain(double* p, const double (&v)[NB])
{
    if (NB == 1) p[0] = v[0];
    else if (NB == 2) *reinterpret_cast<double2*>(p) = make_double2(v[0], v[NB - 1]);
    else
    {
        *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1 % NB]);
        *reinterpret_cast<double2*>(p + 2) = make_double2(v[NB - 1], 0.0);
    }
}
// components [k0, k0 + NBW) of a row (NBW == NB: the whole row with one store; NBW == 1: one component -- the last
// component of a padded row carries the padding lane along so that it never holds the "not yet" pattern afterwards)
template <int NB, int NBW>
__device__ __forceinline__ void pen_publish_part(unsigned long long* p, int k0, const double (&v)[NBW])
{
    if (NBW == NB) { pen_publish<NBW>(p, v); return; }
    if (NB == 3 && k0 == 2)
        asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" :: "l"(p + 2), "l"(__double_as_longlong(v[0])), "l"(0LL) : "memory");
    else
        asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" :: "l"(p + k0), "l"(__double_as_longlong(v[0])) : "memory");
}
template <int NB, int NBW>
__device__ __forceinline__ void pen_store_plain_part(double* p, int k0, const double (&v)[NBW])
{
    if (NBW == NB) { pen_store_plain<NBW>(p, v); return; }
    if (NB == 3 && k0 == 2) *reinterpret_cast<double2*>(p + 2) = make_double2(v[0], 0.0);
    else p[k0] = v[0];
}

__device__ __forceinline__ int pen_ld_progress(unsigned a)
{
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}

// coefficient tables of the pencil sweeps from the U value array (once per matrix)
__global__ void __launch_bounds__(TPB) k_pencil_coeffs(long long n3, const int* __restrict__ idx, const double* __restrict__ Uval,
                                                       double* __restrict__ coef)
{
    for (long long i = (long long)blockIdx.x*TPB + threadIdx.x; i < n3; i += (long long)gridDim.x*TPB)
    {
        const int j = idx[i];
        coef[i] = (j >= 0) ? canon(Uval[j]) : 0.0;
    }
}

// MODE 0 forward sweep (in = rA, out = y), 1 factor (in = diag, out = scratch, rDout = 1/out), 2 backward sweep
// (in = y, out = wA; y re-armed; optional fused dot wA . rA).  LONG = which axis is the long one (0 x, 1 y, 2 z):
// it fixes where the own-lane dependency sits in the reference's z, y, x subtraction order.
template <int NB, int MODE, int LONG, int SW, bool FUSE = false>
__global__ void __launch_bounds__((SW + 1 + PEN_FW + (FUSE ? 1 : 0))*32) k_dic_pencil(PencilArgs A, const double* __restrict__ diagOrRD, const double* __restrict__ rA,
                                                        unsigned long long* y, unsigned long long* out, double* __restrict__ rDout,
                                                        unsigned long long* tileTicket, double* partials, int partStride,
                                                        DevState* st, int gated)
{
    constexpr int NS = PenRow<NB>::NS, RB = PenRow<NB>::BYTES;
    constexpr bool BWD = (MODE == 2), FACTOR = (MODE == 1);
    constexpr int NTH = (SW + 1 + PEN_FW + (FUSE ? 1 : 0))*32;   // SW sweep warps + loader + PEN_FW fetch warps (+ the update warp)
    static_assert(!FUSE || MODE == 0, "the x/r update rides in the forward sweep");
    constexpr int NBW = NB/SW;              // components a sweep warp carries
    static_assert(SW == 1 || SW == NB, "one sweep warp for all components, or one per component");
    if (gated)
    {
        bool d = true;
#pragma unroll
        for (int k = 0; k < NB; k++) d = d && (st[k].done != 0);
        if (d) return;
    }
    const bool withDot = BWD && partials != nullptr;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    unsigned sb;
    asm volatile("mov.u32 %0, %1;" : "=r"(sb) : "r"((unsigned)__cvta_generic_to_shared(smemRaw)));
    const int stageBytes = pen_stage_bytes<NB, MODE>(withDot, FUSE);
    const int NST = A.nStages;
    const int E = A.w1 + A.w2;
    const unsigned sExt = sb + (unsigned)(NST*stageBytes);
    const unsigned sBar = sExt + (unsigned)(PEN_R*E*RB);
    const unsigned sProg = sBar + 64u;
    const unsigned sRdy = sBar + 128u;      // FUSE: per stage, "the update warp is through with it" (what the sweep warps wait for)
    int* slotPtr = (int*)(smemRaw + NST*stageBytes + PEN_R*E*RB + 64 + 16);
    const int lane = threadIdx.x & 31;
    // Which warp sweeps: the CTAs that share an SM take turns (a per-SM counter), so that two sweep warps never sit on
    // the same scheduler -- a sweep warp is issue-bound, two of them on one sub-partition run at half speed each
    // (measured: 0.18 -> 0.45 us per step with two CTAs per SM and the sweep always in warp 0).
    if (threadIdx.x == 0)
    {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        slotPtr[1] = A.smCounter ? (int)(atomicAdd(A.smCounter + (smid & 1023u), 1u) & 3u) : 0;
    }
    __syncthreads();
    // roles: the first four warps (one per scheduler) rotate by the per-SM counter; role < SW: sweep warp of components
    // [role*NBW, ...), role == SW: loader, the two roles above: fetch warps
    const int wid = (int)(threadIdx.x >> 5);
    const int warp = (wid < 4) ? ((wid - slotPtr[1]) & 3) : wid;
    // Persistent CTAs: CTA b sweeps pencils b, b + G, b + 2G, ... of the wavefront order (reversed for the backward sweep).
    // Every CTA of the grid is resident (the host sizes G by the occupancy) and takes its pencils in ascending wavefront
    // order, so a pencil only ever waits on pencils that are done, running, or next in line of a running CTA: no deadlock.
    for (int tileIdx = blockIdx.x; tileIdx < A.nPencils; tileIdx += gridDim.x)
    {
    const int tile = BWD ? A.nPencils - 1 - tileIdx : tileIdx;
    const int* M = A.meta + 8*tile;
    const int base = M[0], w1a = M[1], w2a = M[2];
    const int prodA1 = BWD ? M[5] : M[3], prodA2 = BWD ? M[6] : M[4];
    const int steps = A.steps, nStagesTotal = steps/PEN_SS;
    const double neutral = FACTOR ? 1.0 : 0.0;

    // stage layout: [coef 32*SS*24][dv 32*SS*RB (not FACTOR)][in 32*SS*RB][rA 32*SS*RB (withDot)]
    const unsigned oCoef = 0u, oDv = PEN_SS*32*24, oIn = oDv + (FACTOR ? 0u : (unsigned)(PEN_SS*32*RB));
    const unsigned oRA = oIn + (unsigned)(PEN_SS*32*RB);
    const unsigned oFW = oRA, oFP = oFW + (unsigned)(PEN_SS*32*RB), oFX = oFP + (unsigned)(PEN_SS*32*RB);   // FUSE: wA, pA, x rows

    if (threadIdx.x == 0)
    {
        for (int s = 0; s < NST; s++)
        {
            if (tileIdx != (int)blockIdx.x) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" :: "r"(sBar + 8u*s) : "memory");
            mbar_init(sBar + 8u*s, 1);
            if (FUSE)
            {
                if (tileIdx != (int)blockIdx.x) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" :: "r"(sRdy + 8u*s) : "memory");
                mbar_init(sRdy + 8u*s, 1);
            }
        }
        for (int i = 0; i < SW; i++) asm volatile("st.volatile.shared.s32 [%0], %1;" :: "r"(sProg + 4u*i), "r"(0) : "memory");
    }
    // the ring of external values starts out "not yet"
    for (int i = threadIdx.x; i < PEN_R*E*NS; i += NTH) ((unsigned long long*)(smemRaw + NST*stageBytes))[i] = SENT;
    __syncthreads();

    double dot[NB], vXR[NB];
#pragma unroll
    for (int k = 0; k < NB; k++) { dot[k] = 0.0; vXR[k] = 0.0; }
    const unsigned sStageReady = FUSE ? sRdy : sBar;        // what a sweep warp waits for before it reads a stage

    // slowest sweep warp's progress (steps consumed)
    auto progress = [&]() { int p = pen_ld_progress(sProg); for (int i = 1; i < SW; i++) p = min(p, pen_ld_progress(sProg + 4u*i)); return p; };

    if (warp < SW)
    {
        // ------------------------------------------------------------------------------------------ the sweep
        const int k0 = warp*NBW;                                   // first component of this sweep warp
        const unsigned kOff = 8u*(unsigned)k0;
        // One warp, in-order issue: what a step costs is the number of instructions on it.  So: no index arithmetic in
        // the loop (running addresses, the four steps of a ring stage unrolled with immediate offsets), no divergence
        // (every lane reads "its" external slot -- interior lanes a constant slot -- and selects), and a two-deep software
        // pipeline: the row data of step i + 1 is loaded early in step i and its products are formed right behind the
        // dependent chain of step i, in the shadow of its DMUL / DSUB latencies.
        const int p1 = lane % A.w1, p2 = lane/A.w1;
        const bool laneIn = p2 < A.w2;                             // lanes beyond w1*w2 (tiny cross-sections) idle along
        const bool intA1 = !laneIn || (BWD ? (p1 + 1 < A.w1) : (p1 > 0));      // the a1 dependency comes from a lane of this warp
        const bool intA2 = !laneIn || (BWD ? (p2 + 1 < A.w2) : (p2 > 0));
        const int srcA1 = (BWD ? lane + 1 : lane - 1) & 31, srcA2 = (BWD ? lane + A.w1 : lane - A.w1) & 31;
        // the constant slot interior lanes "poll": the last 32 bytes of the barrier block hold the neutral value
        const unsigned sConst = sBar + 64u + 32u;
        if (lane < NS) asm volatile("st.shared.f64 [%0], %1;" :: "r"(sConst + 8u*lane), "d"(neutral) : "memory");     // every sweep warp writes the same values
        __syncwarp();
        const unsigned stepRing = (unsigned)(E*RB);
        // per-lane ring addresses at ring position 0; interior lanes: the constant slot with a zero stride
        const unsigned ringA1 = intA1 ? sConst : sExt + (unsigned)(RB*p2);
        const unsigned ringA2 = intA2 ? sConst : sExt + (unsigned)(RB*(A.w2 + p1));
        const unsigned strA1 = intA1 ? 0u : stepRing, strA2 = intA2 ? 0u : stepRing;
        const unsigned ringBytes1 = intA1 ? 0u : (unsigned)(PEN_R*E*RB), ringBytes2 = intA2 ? 0u : (unsigned)(PEN_R*E*RB);
        unsigned rOff1 = 0u, rOff2 = 0u;                           // byte offset of the current stage's first step in the ring
        double prev[NBW], dotW[NBW];
#pragma unroll
        for (int k = 0; k < NBW; k++) { prev[k] = neutral; dotW[k] = 0.0; }
        // row data of a step out of the ring stage at sAddr: local step js (compile-time), this lane
        double cf[3], dv[NBW], in[NBW], ra[NBW];
        auto load_static = [&](unsigned sAddr, int js)
        {
            const unsigned r = (unsigned)(js*32) + (unsigned)lane;
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(cf[0]) : "r"(sAddr + oCoef + 24u*r));
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(cf[1]) : "r"(sAddr + oCoef + 24u*r + 8u));
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(cf[2]) : "r"(sAddr + oCoef + 24u*r + 16u));
            if (!FACTOR) pen_lds<NBW>(sAddr + oDv + (unsigned)RB*r + kOff, dv);
            pen_lds<NBW>(sAddr + oIn + (unsigned)RB*r + kOff, in);
            if (withDot) pen_lds<NBW>(sAddr + oRA + (unsigned)RB*r + kOff, ra);
        };
        double F0[NBW], F1[NBW], F2[NBW], acc0[NBW], raC[NBW];
        auto products = [&]()
        {
#pragma unroll
            for (int k = 0; k < NBW; k++)
            {
                if (FACTOR) { F0[k] = mul_here(cf[0], cf[0]); F1[k] = mul_here(cf[1], cf[1]); F2[k] = mul_here(cf[2], cf[2]); acc0[k] = canon(in[k]); }
                else
                {
                    F0[k] = mul_here(dv[k], cf[0]); F1[k] = mul_here(dv[k], cf[1]); F2[k] = mul_here(dv[k], cf[2]);
                    acc0[k] = (MODE == 0) ? mul_here(dv[k], in[k]) : canon(in[k]);
                }
                raC[k] = withDot ? ra[k] : 0.0;
            }
        };
        int slot = 0;
        unsigned phase = 0u;
        unsigned sAddr = sb;                                       // ring stage being swept
        mbar_wait(sStageReady, 0);
        load_static(sAddr, BWD ? PEN_SS - 1 : 0);
        products();
        // global row of the current step: rows ascend with the step forward, descend backward
        size_t rowG = (size_t)base + (size_t)32*(BWD ? steps - 1 : 0) + lane;
        const long long rowStep = BWD ? -32 : 32;
        for (int g = 0; g < nStagesTotal; g++)
        {
            // next stage (for the look-ahead load of its first step)
            int slotN = slot + 1; unsigned phaseN = phase;
            if (slotN == NST) { slotN = 0; phaseN ^= 1u; }
            const unsigned sAddrN = sb + (unsigned)(slotN*stageBytes);
#pragma unroll
            for (int j = 0; j < PEN_SS; j++)
            {
                // a. this step's external values (interior lanes: the constant slot)
                const unsigned a1 = ringA1 + rOff1 + (unsigned)j*strA1 + kOff, a2 = ringA2 + rOff2 + (unsigned)j*strA2 + kOff;
                double xA1[NBW], xA2[NBW];
                pen_lds_volatile<NBW>(a1, xA1);
                pen_lds_volatile<NBW>(a2, xA2);
                // b. the previous step's results of the neighbouring lanes
                double sA1[NBW], sA2[NBW];
#pragma unroll
                for (int k = 0; k < NBW; k++)
                {
                    sA1[k] = __shfl_sync(0xffffffffu, prev[k], srcA1);
                    sA2[k] = __shfl_sync(0xffffffffu, prev[k], srcA2);
                }
                // c. row data of the next step
                const bool more = (j + 1 < PEN_SS) || (g + 1 < nStagesTotal);
                if (j + 1 < PEN_SS) load_static(sAddr, BWD ? PEN_SS - 2 - j : j + 1);
                else if (g + 1 < nStagesTotal)
                {
                    mbar_wait(sStageReady + 8u*slotN, phaseN);
                    load_static(sAddrN, BWD ? PEN_SS - 1 : 0);
                }
                // d. wait for the external values (warp-uniform loop: the warp stays converged)
                {
                    bool ny = false;
#pragma unroll
                    for (int k = 0; k < NBW; k++) ny = ny | hi_is_sent(xA1[k]) | hi_is_sent(xA2[k]);
                    while (__any_sync(0xffffffffu, ny))
                    {
                        pen_lds_volatile<NBW>(a1, xA1);
                        pen_lds_volatile<NBW>(a2, xA2);
                        ny = false;
#pragma unroll
                        for (int k = 0; k < NBW; k++) ny = ny | hi_is_sent(xA1[k]) | hi_is_sent(xA2[k]);
                    }
                }
                // e. the chain, in the reference's order z, y, x; LONG tells which of them is the own-lane dependency:
                //    (z, y, x) = (A2, A1, own) for LONG 0, (A2, own, A1) for LONG 1, (own, A2, A1) for LONG 2
                double res[NBW];
#pragma unroll
                for (int k = 0; k < NBW; k++)
                {
                    const double vA1 = intA1 ? sA1[k] : xA1[k], vA2 = intA2 ? sA2[k] : xA2[k];
                    const double vz = (LONG == 2) ? prev[k] : vA2;
                    const double vy = (LONG == 0) ? vA1 : (LONG == 1) ? prev[k] : vA2;
                    const double vx = (LONG == 0) ? prev[k] : vA1;
                    double a = acc0[k];
                    if (FACTOR)
                    {
                        a = __dsub_rn(a, __ddiv_rn(F0[k], vz));
                        a = __dsub_rn(a, __ddiv_rn(F1[k], vy));
                        a = __dsub_rn(a, __ddiv_rn(F2[k], vx));
                    }
                    else
                    {
                        a = __dsub_rn(a, __dmul_rn(F0[k], vz));
                        a = __dsub_rn(a, __dmul_rn(F1[k], vy));
                        a = __dsub_rn(a, __dmul_rn(F2[k], vx));
                    }
                    res[k] = a;
                    if (withDot) dotW[k] += a*raC[k];
                }
                // f. the next step's products, behind the chain
                if (more) products();
#pragma unroll
                for (int k = 0; k < NBW; k++) prev[k] = res[k];
                // g. publish: the result of the sweep AND the hand-off to the neighbouring pencils; re-arm the slots
                pen_publish_part<NB, NBW>(out + (size_t)NS*rowG, k0, prev);
                if (FACTOR)
                {
                    double r[NBW];
#pragma unroll
                    for (int k = 0; k < NBW; k++) r[k] = __ddiv_rn(1.0, prev[k]);
                    pen_store_plain_part<NB, NBW>(rDout + (size_t)NS*rowG, k0, r);
                }
                {
                    unsigned long long w[NBW];
#pragma unroll
                    for (int k = 0; k < NBW; k++) w[k] = SENT;
                    if (!intA1) pen_sts_bits<NBW>(a1, w);
                    if (!intA2) pen_sts_bits<NBW>(a2, w);
                }
                rowG = (size_t)((long long)rowG + rowStep);
            }
            // the stage is consumed: y "not yet" again for the next forward sweep (backward), hand the stage back
            if (BWD)
            {
                unsigned long long* q = y + (size_t)NS*(rowG + 32);          // first (lowest) row of the stage, this lane
#pragma unroll
                for (int j = 0; j < PEN_SS; j++)
                {
                    if (NBW == NB)
                    {
#pragma unroll
                        for (int k = 0; k < NS; k++) q[(size_t)NS*32*j + k] = SENT;
                    }
                    else
                    {
                        q[(size_t)NS*32*j + k0] = SENT;
                        if (k0 + 1 == NB && NS > NB) q[(size_t)NS*32*j + NB] = SENT;       // the padding lane travels with the last component
                    }
                }
            }
            rOff1 += PEN_SS*strA1; if (rOff1 >= ringBytes1) rOff1 = 0u;
            rOff2 += PEN_SS*strA2; if (rOff2 >= ringBytes2) rOff2 = 0u;
            slot = slotN; phase = phaseN; sAddr = sAddrN;
            // all lanes' re-arming stores precede the progress store: shared-memory accesses of one warp are performed
            // in order (no membar: it would wait for the global stores in flight)
            __syncwarp();
            if (lane == 0) asm volatile("st.volatile.shared.s32 [%0], %1;" :: "r"(sProg + 4u*(unsigned)warp), "r"((g + 1)*PEN_SS) : "memory");
        }
        // this warp's share of the fused dot (static indexing: dot[] stays in registers)
#pragma unroll
        for (int k = 0; k < NB; k++) dot[k] = (NBW == NB) ? dotW[k % NBW] : ((k == k0) ? dotW[0] : 0.0);
    }
    else if (warp == SW)
    {
        // ------------------------------------------------------------------------------------------ the loader
        if (lane == 0)
        {
            const double* coefTab = BWD ? A.coefB : A.coefF;
            const double* inArr = FACTOR ? diagOrRD : (BWD ? (const double*)y : rA);
            for (int g = 0; g < nStagesTotal; g++)
            {
                const int s = g % NST;
                if (g >= NST)
                    while (progress() < (g - NST + 1)*PEN_SS) { __nanosleep(200); }
                const int tFirst = BWD ? steps - PEN_SS*(g + 1) : PEN_SS*g;
                const size_t r0 = (size_t)base + (size_t)32*tFirst;
                const unsigned dst = sb + (unsigned)(s*stageBytes), bar = sBar + 8u*s;
                mbar_expect_tx(bar, (unsigned)stageBytes);
                bulk_g2s(dst + oCoef, coefTab + 3*r0, PEN_SS*32*24, bar);
                if (!FACTOR) bulk_g2s(dst + oDv, diagOrRD + (size_t)NS*r0, PEN_SS*32*RB, bar);
                bulk_g2s(dst + oIn, inArr + (size_t)NS*r0, PEN_SS*32*RB, bar);
                if (withDot) bulk_g2s(dst + oRA, rA + (size_t)NS*r0, PEN_SS*32*RB, bar);
                if (FUSE)
                {
                    bulk_g2s(dst + oFW, (const double*)A.fxWA + (size_t)NS*r0, PEN_SS*32*RB, bar);
                    bulk_g2s(dst + oFP, A.fxPA + (size_t)NS*r0, PEN_SS*32*RB, bar);
                    bulk_g2s(dst + oFX, A.fxX + (size_t)NS*r0, PEN_SS*32*RB, bar);
                }
            }
        }
    }
    else if (FUSE && warp == SW + PEN_FW + 1)
    {
        // ------------------------------------------------------------------------------------------ the update warp
        // x += alpha pA;  rA -= alpha wA;  sum|rA|;  wA re-armed -- the BLAS-1 tail of the previous PCG iteration (k_xr_update_v),
        // applied stage by stage to the rows the loader has just brought in: the new rA goes back to global memory AND into the
        // stage's `in` slot, where the sweep warps pick it up.  The sweep is a dependency chain with the SM's bandwidth idle;
        // this warp uses it.
        bool pend[NB], any = false;
        double alpha[NB];
#pragma unroll
        for (int k = 0; k < NB; k++)
        {
            pend[k] = st[k].pending != 0;
            any = any || pend[k];
            alpha[k] = pend[k] ? __ddiv_rn(st[k].wArA, st[k].wApA) : 0.0;
        }
        unsigned long long sentBits[NS];
#pragma unroll
        for (int k = 0; k < NS; k++) sentBits[k] = SENT;
        for (int g = 0; g < nStagesTotal; g++)
        {
            const int s = g % NST;
            const unsigned ph = (unsigned)((g/NST) & 1);
            const unsigned stg = sb + (unsigned)(s*stageBytes);
            mbar_wait(sBar + 8u*s, ph);
            if (any)
            {
                const size_t r0 = (size_t)base + (size_t)32*(PEN_SS*g);
#pragma unroll
                for (int i = 0; i < PEN_SS; i++)
                {
                    const unsigned r = (unsigned)(i*32 + lane);
                    double w[NB], pv[NB], xx[NB], rr[NB];
                    pen_lds<NB>(stg + oFW + (unsigned)RB*r, w);
                    pen_lds<NB>(stg + oFP + (unsigned)RB*r, pv);
                    pen_lds<NB>(stg + oFX + (unsigned)RB*r, xx);
                    pen_lds<NB>(stg + oIn + (unsigned)RB*r, rr);
#pragma unroll
                    for (int k = 0; k < NB; k++)
                    {
                        if (pend[k])
                        {
                            xx[k] = __dadd_rn(xx[k], __dmul_rn(alpha[k], pv[k]));
                            rr[k] = __dsub_rn(rr[k], __dmul_rn(alpha[k], w[k]));
                            vXR[k] += fabs(rr[k]);
                        }
                    }
                    pen_store_plain<NB>(A.fxX + (size_t)NS*(r0 + r), xx);
                    pen_store_plain<NB>(A.fxRA + (size_t)NS*(r0 + r), rr);
                    {
                        unsigned long long* q = A.fxWA + (size_t)NS*(r0 + r);
                        if (NS == 4) { *reinterpret_cast<ulonglong2*>(q) = make_ulonglong2(SENT, SENT); *reinterpret_cast<ulonglong2*>(q + 2) = make_ulonglong2(SENT, SENT); }
                        else { for (int k = 0; k < NS; k++) q[k] = sentBits[k]; }
                    }
                    // the sweep warps read the updated rA from the stage
                    unsigned long long bits[NB];
#pragma unroll
                    for (int k = 0; k < NB; k++) bits[k] = (unsigned long long)__double_as_longlong(rr[k]);
                    pen_sts_bits<NB>(stg + oIn + (unsigned)RB*r, bits);
                }
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(sRdy + 8u*s) : "memory");
        }
    }
    else if (warp > SW && warp <= SW + PEN_FW)
    {
        // ------------------------------------------------------------------------------------------ the fetch lanes
        const int fl = (warp - SW - 1)*32 + lane;
        const int SPR = (PEN_FW*32)/E;                          // steps covered by one round of the fetch lanes
        if (SPR > 0 && fl < SPR*E)
        {
            const int e = fl % E, off = fl/E;
            const bool isA1 = e < A.w2;
            const int pe = isA1 ? e : e - A.w2;                 // p2 of an a1-face value, p1 of an a2-face value
            const int prod = isA1 ? prodA1 : prodA2;
            const bool laneActive = isA1 ? (pe < w2a) : (pe < w1a);
            // producer row of consumer step t:  base' + 32*(t + dt) + lane'
            const int dt = BWD ? -(isA1 ? A.w1 - 1 : A.w2 - 1) : (isA1 ? A.w1 - 1 : A.w2 - 1);
            const int lanep = BWD ? (isA1 ? A.w1*pe : pe) : (isA1 ? (A.w1 - 1) + A.w1*pe : pe + A.w1*(A.w2 - 1));
            // consumer cell index along the pencil at step t: q = t - qoff
            const int qoff = BWD ? (isA1 ? (A.w1 - 1) + pe : pe + (A.w2 - 1)) : pe;
            const size_t pbase = (prod >= 0) ? (size_t)A.meta[8*prod] : 0;
            const unsigned dstBase = sExt + (unsigned)(RB*e);
            unsigned long long nb[NB];
#pragma unroll
            for (int k = 0; k < NB; k++) nb[k] = (unsigned long long)__double_as_longlong(neutral);
            // Up to PEN_FU polls in flight per lane, for its next steps (next, next + SPR, ...).  Values become ready in step
            // order, so the lane consumes the ready PREFIX of its window and the window adapts: a poll that comes back
            // "not yet" is wasted L2 traffic on exactly the lines the producer is about to write, so a lane that runs
            // close behind its producer polls one step ahead only, a lane that is far behind (started late) catches up
            // with the full window.
            int next = off, W = A.fetchWindow > 0 ? A.fetchWindow : 1;
            const bool adaptive = A.fetchWindow <= 0;
            SpinWatch watch;
            bool bail = false;
            // One poll in flight (the default): the components of a row are published by different sweep warps (SW = NB) that
            // drift against each other by a few steps, so a component is handed over the moment IT arrives -- waiting for the
            // whole row would tie every consumer warp to the slowest producer warp at every pencil face.
            const bool perCmpt = (W == 1 && !adaptive && NB > 1);
            unsigned got = 0u;
            constexpr unsigned ALLC = (1u << NB) - 1u;
            while (next < steps)
            {
                const int prog = progress();
                unsigned long long w[PEN_FU][NB];
                bool can[PEN_FU];
                int issued = 0;
#pragma unroll
                for (int u = 0; u < PEN_FU; u++)
                {
                    const int i = next + u*SPR;
                    can[u] = (u < W) && (i < steps) && (i - PEN_R < prog);      // ring slot (i mod R) free: step i - R consumed
                    if (can[u])
                    {
                        issued++;
                        const int t = BWD ? steps - 1 - i : i;
                        const int q = t - qoff;
                        const bool valid = prod >= 0 && laneActive && q >= 0 && q < A.n0 && !bail;
                        if (valid) pen_peek<NB>(out + (size_t)NS*(pbase + (size_t)32*(t + dt) + lanep), w[u]);
                        else { for (int k = 0; k < NB; k++) w[u][k] = nb[k]; }
                    }
                }
                int m = 0;
                if (perCmpt)
                {
                    if (can[0])
                    {
                        const unsigned dst = dstBase + (unsigned)((next & (PEN_R - 1))*E*RB);
#pragma unroll
                        for (int k = 0; k < NB; k++)
                        {
                            const bool fresh = !(got & (1u << k)) && (w[0][k] != SENT);
                            if (fresh) asm volatile("st.volatile.shared.b64 [%0], %1;" :: "r"(dst + 8u*k), "l"(w[0][k]) : "memory");
                            got |= fresh ? (1u << k) : 0u;
                        }
                        if (got == ALLC) { m = 1; got = 0u; }
                    }
                }
                else
                {
#pragma unroll
                for (int u = 0; u < PEN_FU; u++)
                {
                    bool ok = can[u] && (m == u);
#pragma unroll
                    for (int k = 0; k < NB; k++) ok = ok && (w[u][k] != SENT);
                    if (ok)
                    {
                        pen_sts_bits<NB>(dstBase + (unsigned)(((next + u*SPR) & (PEN_R - 1))*E*RB), w[u]);
                        m++;
                    }
                }
                }
                next += m*SPR;
                if (adaptive) W = (issued > 0 && m == issued) ? min(PEN_FU, W + 1) : max(1, m);
                if (m > 0) watch.progress();
                else if (issued > 0 && watch.expired(SPIN_PENCIL_FETCH)) bail = true;   // (issued == 0: waiting for this CTA's own sweep)
                if (m == 0) __nanosleep(A.fetchSleep);
            }
        }
    }
    // ---------------------------------------------------------------------------------------------- epilogue
    if (FUSE)       // every CTA takes part (S_XR is a no-op when nothing was pending)
        grid_reduce_finish_v<1, NB, NTH>(vXR, partials, partStride, tile, A.nPencils, st, S_XR, A.fxHistory, A.fxHistStride);
    if (withDot)
    {
        if (NB == 1)
        {
            double d1[1] = {dot[0]};
            grid_reduce_finish<1, NTH>(d1, partials, partStride, tile, A.nPencils, st, S_SWEEP, nullptr);
        }
        else
            grid_reduce_finish_v<1, NB, NTH>(dot, partials, partStride, tile, A.nPencils, st, S_SWEEP, nullptr, 0);
    }
    __syncthreads();        // every role is done with this pencil's ring, barriers and progress words
    }
}

} // namespace gpupcg
