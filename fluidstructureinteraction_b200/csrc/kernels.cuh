// kernels.cuh -- hand-written sm_100a kernels of the gpuPCG solve.
//
// Data layout (see plan.h): rows are renumbered into DIC wavefront order and stored SELL-32:
// a slice = 32 consecutive rows = one warp, all of one dependency level; entry j of lane k of
// slice s lives at off + j*32 + k, so every coefficient / index load of a warp is one fully
// coalesced 128 B / 256 B request.  The symmetric coefficient upper[f] is stored ONCE (owner-side
// block Uval); the neighbour-side list carries an index into Uval (gather, L2-resident because
// the owning row sits one wavefront earlier).
//
// Arithmetic contract: every product and sum that exists in the reference face loops is issued
// with __dmul_rn/__dadd_rn/__dsub_rn/__ddiv_rn in the reference's operand order, so no FMA
// contraction can occur and Amul / DIC results are bit-identical to the CPU loops.  Only the
// global reductions (gSumProd, gSumMag, gSum) are re-associated (fixed tree, deterministic).
//
// All kernels are HBM-bound FP64 streaming / gather kernels: no tensor cores by design
// (0.19 flop/B).  Grids are persistent: gridDim = numSMs x k, slices dealt round-robin.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "plan.h"
#include "comm.h"

namespace gpupcg
{

// "not yet written" marker of the dataflow sweeps.  A quiet NaN with a payload that arithmetic
// cannot produce: every published value is canonicalised (publish()) if it is a NaN.
static constexpr unsigned long long SENT = 0xFFF8B200DEADC0DEULL;
static constexpr unsigned long long QNAN = 0x7FF8000000000000ULL;

static constexpr int TPB = 256;             // threads per block of every kernel
static constexpr int WPB = TPB/32;          // warps per block
static constexpr int MAXPART = 4096;        // max CTAs of any streaming (grid-stride) reducing kernel

static constexpr double K_VSMALL = 1.0e-300;    // Foam::VSMALL
static constexpr double K_SMALL  = 1.0e-15;     // Foam::SMALL
static constexpr double K_MSMALL = 1.0e-20;     // lduMatrix::small_
static constexpr double K_MGREAT = 1.0e+20;     // lduMatrix::great_

// Device-resident state of one PCG::solve.  The host only reads it (pinned copy).
struct DevState
{
    double wArA, wArAold, wApA;
    double normFactor, initialResidual, finalResidual;
    double xRef;
    double tol, relTol;
    double red[4];              // local reduction results awaiting a cross-rank all-reduce
    long long nGlobalCells;
    int minIter, maxIter;
    int nIter, done, converged, singular;
    int historyCap, multiRank;
    unsigned int ticket;
    int pending;                // an x/r update (alpha known, wA = A pA) has not been applied yet
    P2PDev* p2p;                // multi-rank: peer-memory mailboxes for the scalar all-reduces (null: ncclAllReduce + k_scalar_step)
    int ordered;                // GPUPCG_ORDERED_SUMS: the fused tree reductions stand down, k_ordered_step sums in the caller's cell order
};

enum ScalarStep { S_INIT = 0, S_NORM = 1, S_SWEEP = 2, S_AMUL = 3, S_XR = 4 };

// ------------------------------------------------------------------------------------------
// scalar recurrences of PCG::solve [EXT PCG.C], evaluated by exactly one thread
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool dev_stop(DevState* st)
{
    // lduMatrix::solver::stop + solverPerformance::checkConvergence [EXT]
    if (st->nIter < st->minIter) return false;
    bool conv = (st->finalResidual < st->tol)
             || (st->relTol > K_SMALL && st->finalResidual <= __dmul_rn(st->relTol, st->initialResidual));
    st->converged = conv ? 1 : 0;
    return (st->nIter >= st->maxIter) || conv;
}

__device__ __forceinline__ void dev_scalar_step(DevState* st, int step, const double* v, double* history)
{
    switch (step)
    {
        case S_INIT:    // gAverage(x) = gSum(x)/global size
            st->xRef = __ddiv_rn(v[0], (double)st->nGlobalCells);
            break;
        case S_NORM:    // normFactor; initial residual; first stop() test
            st->normFactor = __dadd_rn(v[0], K_MSMALL);
            st->initialResidual = __ddiv_rn(v[1], st->normFactor);
            st->finalResidual = st->initialResidual;
            st->nIter = 0;
            st->wArA = K_MGREAT;
            st->wArAold = K_MGREAT;
            st->done = dev_stop(st) ? 1 : 0;
            break;
        case S_SWEEP:   // wArAold = wArA; wArA = gSumProd(wA, rA)
            st->wArAold = st->wArA;
            st->wArA = v[0];
            break;
        case S_AMUL:    // wApA = gSumProd(wA, pA); checkSingularity(mag(wApA)/normFactor)
            st->wApA = v[0];
            if (!(__ddiv_rn(fabs(v[0]), st->normFactor) > K_VSMALL))
            {
                st->singular = 1;
                st->done = 1;
            }
            else
            {
                st->singular = 0;
                st->pending = 1;
            }
            break;
        case S_XR:      // finalResidual = gSumMag(rA)/normFactor; nIterations++; stop()
            if (!st->pending) break;
            st->pending = 0;
            st->finalResidual = __ddiv_rn(v[0], st->normFactor);
            if (history && st->nIter < st->historyCap) history[st->nIter] = st->finalResidual;
            st->nIter++;
            if (dev_stop(st)) st->done = 1;
            break;
    }
}

// ------------------------------------------------------------------------------------------
// Watchdog of the device-side waits.  Every wait of the library is rooted in a poll of GLOBAL memory (a fetch lane waiting
// for another tile's / pencil's row, a rank waiting for a peer's mailbox flag); everything else waits on those.  A root
// wait that makes no progress for g_spinTimeoutNs (a dead peer, a schedule bug) records its code in g_spinFault and lets
// go: it delivers zeros, the kernel runs to its end on garbage, and the host returns GPUPCG_ERR_STATE instead of hanging
// the GPU.  Cost: a counter in the retry path of the poll; %globaltimer is read once every (mask + 1) failed polls.
// ------------------------------------------------------------------------------------------
__device__ unsigned int g_spinFault = 0;                       // 0, or the code of the first wait that timed out
__device__ unsigned long long g_spinTimeoutNs = 20000000000ULL;
__device__ unsigned int g_spinCheckMask = 1023u;
enum SpinCode { SPIN_P2P = 1, SPIN_TILE_FETCH = 2, SPIN_TILE_FETCH_V = 3, SPIN_PENCIL_FETCH = 4, SPIN_GLOBAL = 5 };

struct SpinWatch
{
    unsigned int n = 0;
    unsigned long long t0 = 0;
    // call after every failed poll; true: give up
    __device__ __forceinline__ bool expired(int code)
    {
        if ((++n & g_spinCheckMask) != 0u) return false;
        if (*(volatile unsigned int*)&g_spinFault) return true;         // somebody else already gave up: follow at once
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t0 == 0) { t0 = t; return false; }
        // a peer rank may legitimately be late by what its HOST is late (plan build, set_matrix): six times the budget of a
        // wait whose producer is a kernel already running on this device
        const unsigned long long lim = (code == SPIN_P2P) ? 6ULL*g_spinTimeoutNs : g_spinTimeoutNs;
        if (t - t0 > lim) { atomicCAS(&g_spinFault, 0u, (unsigned int)code); return true; }
        return false;
    }
    __device__ __forceinline__ void progress() { t0 = 0; }
};

// One-shot all-reduce of st->red[0..count) through the peers' mailboxes (comm.h), by ONE thread: write my values and
// then my flag (= sequence number, release at system scope) into every rank's mailbox, wait for every rank's flag in my
// own mailbox, sum in rank order.  Every rank gets the same bits.
__device__ __forceinline__ void p2p_allreduce(P2PDev* c, const double* mine, int count, double* out)
{
    const unsigned long long seq = ++c->seq;
    const int par = (int)(seq & 1ULL), me = c->rank;
    for (int r = 0; r < c->nRanks; r++)
    {
        P2PMailbox* m = c->peer[r];
        for (int k = 0; k < count; k++)
            asm volatile("st.relaxed.sys.global.f64 [%0], %1;" :: "l"(&m->val[par][me][k]), "d"(mine[k]) : "memory");
    }
    for (int r = 0; r < c->nRanks; r++)
        asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(&c->peer[r]->flag[par][me]), "l"(seq) : "memory");
    P2PMailbox* own = c->peer[me];
    for (int k = 0; k < count; k++) out[k] = 0.0;
    for (int r = 0; r < c->nRanks; r++)
    {
        unsigned long long f;
        SpinWatch watch;
        do { asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(&own->flag[par][r]) : "memory"); }
        while (f != seq && !watch.expired(SPIN_P2P));
        for (int k = 0; k < count; k++)
        {
            double v;
            asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(&own->val[par][r][k]) : "memory");
            out[k] += v;
        }
    }
}

// ------------------------------------------------------------------------------------------
// deterministic grid reduction: block tree -> partials[block] -> the last block to finish sums the
// partials in a fixed order and either advances the scalar recurrence (single rank) or parks the
// local sums in st->red for the all-reduce (multi rank).
// ------------------------------------------------------------------------------------------
template <int NV, int NT>
__device__ __forceinline__ void grid_reduce_finish(double (&v)[NV], double* partials, int stride, int slot, int nSlots,
                                                   DevState* st, int step, double* history, int redSlot = 0,
                                                   bool lastOfStep = true)
{
    constexpr int NW = NT/32;
    __shared__ double sm[NV][NW];
    __shared__ bool amLast;
    if (st->ordered) return;    // verification mode: k_ordered_step (kernels_ordered.cuh) forms the sums after this kernel
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++)
    {
        double a = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) sm[k][warp] = a;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
#pragma unroll
        for (int k = 0; k < NV; k++)
        {
            double a = sm[k][0];
#pragma unroll
            for (int w = 1; w < NW; w++) a += sm[k][w];
            __stcg(&partials[(size_t)k*stride + slot], a);
        }
        __threadfence();
        const unsigned int t = atomicAdd(&st->ticket, 1u);
        amLast = (t == (unsigned int)nSlots - 1);
    }
    __syncthreads();
    if (!amLast) return;
    __threadfence();
    // fixed-shape final sum: thread t takes partials t, t+NT, ... then the same block tree
    double f[NV];
#pragma unroll
    for (int k = 0; k < NV; k++)
    {
        double a = 0.0;
        for (int i = threadIdx.x; i < nSlots; i += NT) a += __ldcg(&partials[(size_t)k*stride + i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) sm[k][warp] = a;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
#pragma unroll
        for (int k = 0; k < NV; k++)
        {
            double a = sm[k][0];
#pragma unroll
            for (int w = 1; w < NW; w++) a += sm[k][w];
            f[k] = a;
        }
        st->ticket = 0;
        if (st->multiRank)
        {
            // park the local sums for the all-reduce; the Amul dot arrives in two parts (rows without / with
            // coupled faces): slot 0 from k_amul_tile (which also clears slot 1), slot 1 from k_iface_update
#pragma unroll
            for (int k = 0; k < NV; k++) st->red[redSlot + k] = f[k];
            if (step == S_AMUL && redSlot == 0) st->red[1] = 0.0;
            if (st->p2p && lastOfStep)
            {
                // the cross-rank sum and the scalar step happen right here (no ncclAllReduce, no k_scalar_step launch)
                const int count = (step == S_AMUL || step == S_NORM) ? 2 : 1;
                double loc[2] = {st->red[0], st->red[1]}, g[4] = {0.0, 0.0, 0.0, 0.0};
                p2p_allreduce(st->p2p, loc, count, g);
                if (step == S_AMUL) g[0] = g[0] + g[1];
                if (!(st->done && step >= S_SWEEP)) dev_scalar_step(st, step, g, history);
            }
        }
        else
        {
            dev_scalar_step(st, step, f, history);
        }
    }
}

// multi-rank: after the all-reduce of st->red
__global__ void k_scalar_step(DevState* st, int step, double* history)
{
    if (threadIdx.x == 0 && blockIdx.x == 0)
    {
        if (st->done && step >= S_SWEEP) return;
        double v[4] = {st->red[0], st->red[1], st->red[2], st->red[3]};
        if (step == S_AMUL) v[0] = v[0] + v[1];
        dev_scalar_step(st, step, v, history);
    }
}

// ------------------------------------------------------------------------------------------
// layout kernels: original cell/face order <-> wavefront SELL order
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_rows_in(const double* __restrict__ srcOld, double* __restrict__ dstNew,
                                                 const int* __restrict__ rowOld, int nRows, double padValue)
{
    for (int r = blockIdx.x*TPB + threadIdx.x; r < nRows; r += gridDim.x*TPB)
    {
        const int c = rowOld[r];
        dstNew[r] = (c >= 0) ? srcOld[c] : padValue;
    }
}

__global__ void __launch_bounds__(TPB) k_rows_out(const double* __restrict__ srcNew, double* __restrict__ dstOld,
                                                  const int* __restrict__ rowOld, int nRows)
{
    for (int r = blockIdx.x*TPB + threadIdx.x; r < nRows; r += gridDim.x*TPB)
    {
        const int c = rowOld[r];
        if (c >= 0) dstOld[c] = srcNew[r];
    }
}

__global__ void __launch_bounds__(TPB) k_coeffs_in(const double* __restrict__ upperOld, double* __restrict__ Uval,
                                                   const int* __restrict__ UfaceOld, long long nEntries)
{
    for (long long e = (long long)blockIdx.x*TPB + threadIdx.x; e < nEntries; e += (long long)gridDim.x*TPB)
    {
        const int f = UfaceOld[e];
        Uval[e] = (f >= 0) ? upperOld[f] : 0.0;
    }
}

__global__ void __launch_bounds__(TPB) k_fill64(unsigned long long* __restrict__ p, unsigned long long v, long long n)
{
    for (long long i = (long long)blockIdx.x*TPB + threadIdx.x; i < n; i += (long long)gridDim.x*TPB) p[i] = v;
}

// gather x[faceCells] of every coupled patch into the send buffer (initInterfaceMatrixUpdate)
__global__ void __launch_bounds__(TPB) k_iface_gather(const double* __restrict__ x, const int* __restrict__ ifFaceRow,
                                                      double* __restrict__ sendBuf, int n)
{
    for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) sendBuf[i] = x[ifFaceRow[i]];
}

// updateInterfaceMatrix: result[faceCells[i]] -= coeffs[i]*xNbr[i], patches in index order, faces in
// patch order.  One thread per row that owns interface faces, entries in (patch, i) order.
// constX != 0: the neighbour field is the constant st->xRef (normFactor's Amul).
// dotMode: add this row's  Ax*x  to the reduction (rows masked out of the main kernel's dot).
template <int DOT>
__global__ void __launch_bounds__(TPB) k_iface_update(const int* __restrict__ ifRow, const int* __restrict__ ifRowStart,
                                                      const int* __restrict__ ifEntry, int nIfRows,
                                                      const double* __restrict__ bou, const double* __restrict__ xNbr,
                                                      int constX, const double* __restrict__ x,
                                                      double* __restrict__ Ax, double* partials, DevState* st, int gated)
{
    if (gated && st->done) return;
    double dot[1] = {0.0};
    const double xc = constX ? st->xRef : 0.0;
    for (int k = blockIdx.x*TPB + threadIdx.x; k < nIfRows; k += gridDim.x*TPB)
    {
        const int r = ifRow[k];
        double acc = Ax[r];
        for (int q = ifRowStart[k]; q < ifRowStart[k + 1]; q++)
        {
            const int e = ifEntry[q];
            const double xn = constX ? xc : xNbr[e];
            acc = __dsub_rn(acc, __dmul_rn(bou[e], xn));
        }
        Ax[r] = acc;
        if (DOT) dot[0] += acc*x[r];
    }
    // multi-rank only (a sub-domain with coupled faces): second part of wApA, summed with the first after the all-reduce
    if (DOT) grid_reduce_finish<1, TPB>(dot, partials, MAXPART, blockIdx.x, gridDim.x, st, S_AMUL, nullptr, 1);
}

// ------------------------------------------------------------------------------------------
// dataflow primitives of the tiled sweeps: across tiles the value IS the flag (SENT = not yet).
// ------------------------------------------------------------------------------------------
__device__ int g_spinNs = 0;     // back-off of the global polls (0 = tight), set from GPUPCG_SPIN_NS
__device__ unsigned long long* g_tileTimes = nullptr;   // debug: 3 x globaltimer per tile (GPUPCG_TILE_TRACE)

__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ unsigned long long peek64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long wait_global(const unsigned long long* p)
{
    unsigned long long v = peek64(p);
    const int ns = g_spinNs;
    SpinWatch watch;
    while (v == SENT)
    {
        if (watch.expired(SPIN_GLOBAL)) return 0ULL;
        if (ns > 0) __nanosleep(ns);
        v = peek64(p);
    }
    return v;
}

__device__ __forceinline__ void publish(unsigned long long* p, double v)
{
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    if (v != v) b = QNAN;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(b) : "memory");
}

__device__ __forceinline__ double wait_shared(const unsigned long long* p)
{
    unsigned long long v;
    do { v = *(const volatile unsigned long long*)p; } while (v == SENT);
    return __longlong_as_double((long long)v);
}

// Static description of the tiling handed to the tile kernels (Amul, DIC factor, DIC sweeps).
// Everything a tile needs lives in CONTIGUOUS global ranges (rows are numbered tile-major, the SELL blocks of a
// tile are adjacent), so a tile is staged into shared memory with a handful of 1-D bulk copies (cp.async.bulk,
// SASS UBLKCP) completing on one mbarrier -- no per-thread dependent index->value load chains.
struct AmulTileMeta { int r0, nr, eU0, nEU, eL0, nEL, nX, nC; };     // one tile's contiguous ranges (rows, U / L entries)

struct TileArgs
{
    const SliceInfo* slices;
    const int* tileStart;           // [nTiles+1]
    const int* tileItemPtr;         // [nTiles+1]
    const int2* items;              // {first row, count | first<<16 | last<<17}
    const unsigned short* colL16;   // neighbour-side column codes   (0xFFFF pad, 0x8000|k cross-tile, else local row)
    const unsigned short* idxL16;   // neighbour-side coefficient: index into the tile's U block (local owners)
    const unsigned short* colU16;   // owner-side column codes
    const int* extFPtr;             // [nTiles+1] cross-tile entries of the neighbour side ...
    const int* extFCol;             //   their global rows
    const int* extFIdx;             //   and the U-array index of their coefficient
    const int* extBPtr;             // owner side
    const int* extBCol;
    int nTiles;
    int tileRowsMax, maxEU, maxEL, maxExtF, maxExtB, maxItems;     // shared-memory carve-up
    // ---- Amul view (built once per mesh in gpupcg.cu: build_amul_view)
    const AmulTileMeta* amulMeta;   // [nTiles]
    const unsigned short* amulU16;  // owner-side operand: index into xv, 0x8000 = pad
    const unsigned int* amulL32;    // neighbour-side operand: xv index | cv index << 16, bit 31 = pad
    const int* amulExtX;            // [nTiles][amulXStride] global rows of the cross-tile x values, -1 = none
    const int* amulExtC;            // [nTiles][amulCStride] U-array indices of the cross-tile coefficients, -1 = none
    int amulXStride, amulCStride;   // multiples of AMUL_TPB
    int nAmulTiles, amulT, amulEU, amulEL;  // Amul tiles (pieces of the sweep tiles) and their shared-memory maxima
};

// ---- mbarrier + 1-D bulk copy (TMA) primitives
__device__ __forceinline__ void mbar_init(unsigned bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar)
{
    if (bytes)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned phase)
{
    unsigned ok = 0;
    while (!ok)
    {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
    }
}

// Tiles are handed out in launch order through a never-reset ticket: CTA k of a launch gets tile
// k mod nTiles (forward) or its mirror (backward).  A running CTA therefore only ever waits on tiles whose
// CTAs are already running or finished: no deadlock, whatever the hardware's block scheduling order.
__device__ __forceinline__ int take_tile(unsigned long long* tileTicket, int nTiles, int* slot, bool backward)
{
    if (threadIdx.x == 0)
    {
        const unsigned long long t = atomicAdd(tileTicket, 1ULL);
        const int k = (int)(t % (unsigned long long)nTiles);
        *slot = backward ? nTiles - 1 - k : k;
    }
    __syncthreads();
    return *slot;
}

// ------------------------------------------------------------------------------------------
// K8  lduMatrix::Amul  [EXT lduMatrixATmul.C], one CTA per tile, row-gather form out of shared memory.
// MODE 0: Ax = A x                      (test entry, first residual)
// MODE 1: Ax = A x and  sum(Ax*x)       (PCG: wA = A pA, wApA = gSumProd(wA, pA)); gated on st->done
// MODE 2: Ax = A (xRef field)           (normFactor)
// Bulk-staged per tile: U coefficient block, 16-bit column/coefficient codes, diag, x.  Cross-tile x values and
// cross-tile neighbour-side coefficients are gathered (L2 hits: the owning tile streams them too).
// Per-row operand order = reference face order: diag*x, neighbour-side faces ascending, owner-side ascending.
// ifMask (may be null): rows that own coupled faces are left out of the MODE-1 dot here (k_iface_update adds them).
// ------------------------------------------------------------------------------------------
static constexpr int AMUL_TPB = 128;
#ifndef AMUL_MINB
#define AMUL_MINB 1
#endif
#ifndef AMUL_V_MINB
#define AMUL_V_MINB 1
#endif
static constexpr int AMUL_RX = 4;       // cross-tile x gathers / coefficient gathers whose indices a thread carries in
static constexpr int AMUL_RC = 2;       // registers one tile ahead (more than that: blocking index loads)

// Shared-memory carve-up of one stage (byte offsets):
//   cv [EU + XC] doubles: the tile's U block, then its cross-tile neighbour-side coefficients
//   dg [T]       doubles
//   xv [T + XX]  doubles: the tile's x, then its cross-tile x values (neighbour side first, then owner side)
//   U16 [EU] u16, L32 [EL] u32: operand codes (see AmulCodes in gpupcg.cu), sl [T/32] SliceInfo, mbarrier
struct AmulLayout
{
    unsigned cv, dg, xv, U16, L32, sl, bar, bytes;
};

__host__ __device__ inline AmulLayout amul_layout(int T, int EU, int EL, int XX, int XC)
{
    AmulLayout L;
    L.cv = 0;
    L.dg = L.cv + 8u*(EU + XC);
    L.xv = L.dg + 8u*T;
    L.L32 = L.xv + 8u*(T + XX);
    L.U16 = L.L32 + 4u*EL;
    L.sl = (L.U16 + 2u*EU + 15u) & ~15u;
    L.bar = L.sl + 16u*(T/32 + 1);
    L.bytes = L.bar + 16;
    return L;
}

__device__ __forceinline__ void cp_async8(unsigned dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Persistent, double-buffered: a CTA walks tiles blockIdx.x, +gridDim.x, ...; while tile k is computed out of
// buffer k&1 the bulk copies of tile k+1 and its cross-tile gathers (8-byte cp.async, indices prefetched into
// registers while tile k-1 was computed) are in flight into the other buffer.  One __syncthreads per tile.
// Operand codes are direct indices into xv / cv (built once per mesh), so the row loop is branch-free:
// all codes of a row are loaded first, then all operands, then the products are added in reference order.
template <int MODE>
__global__ void __launch_bounds__(AMUL_TPB, AMUL_MINB) k_amul_tile(TileArgs A, const double* __restrict__ diag,
                                                        const double* __restrict__ Uval, const int* __restrict__ rowOld,
                                                        const unsigned int* __restrict__ ifMask,
                                                        const double* __restrict__ x, double* __restrict__ Ax,
                                                        double* partials, int partStride, DevState* st)
{
    if (MODE == 1 && st->done) return;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    const AmulLayout L = amul_layout(A.amulT, A.amulEU, A.amulEL, A.amulXStride, A.amulCStride);
    const unsigned sb0 = (unsigned)__cvta_generic_to_shared(smemRaw);
    const unsigned stageBytes = (L.bytes + 127u) & ~127u;
    const double xc = (MODE == 2) ? st->xRef : 0.0;
    const int tid = threadIdx.x;
    const int T = A.amulT, EU = A.amulEU;
    const int nkX = A.amulXStride/AMUL_TPB, nkC = A.amulCStride/AMUL_TPB;

    auto issue_bulk = [&](const AmulTileMeta& M, int stage)
    {
        const unsigned sb = sb0 + stage*stageBytes, bar = sb + L.bar;
        const unsigned bytes = 8u*M.nEU + 2u*M.nEU + 4u*M.nEL + 8u*M.nr + (MODE == 2 ? 0u : 8u*M.nr) + 16u*(M.nr >> 5);
        mbar_expect_tx(bar, bytes);
        bulk_g2s(sb + L.cv, Uval + M.eU0, 8u*M.nEU, bar);
        bulk_g2s(sb + L.L32, A.amulL32 + M.eL0, 4u*M.nEL, bar);
        bulk_g2s(sb + L.U16, A.amulU16 + M.eU0, 2u*M.nEU, bar);
        bulk_g2s(sb + L.dg, diag + M.r0, 8u*M.nr, bar);
        if (MODE != 2) bulk_g2s(sb + L.xv, x + M.r0, 8u*M.nr, bar);
        bulk_g2s(sb + L.sl, A.slices + (M.r0 >> 5), 16u*(M.nr >> 5), bar);
    };
    int iX[AMUL_RX], iC[AMUL_RC];
    auto load_indices = [&](int tile)
    {
        const int* lx = A.amulExtX + (size_t)tile*A.amulXStride + tid;
        const int* lc = A.amulExtC + (size_t)tile*A.amulCStride + tid;
#pragma unroll
        for (int k = 0; k < AMUL_RX; k++) iX[k] = (k < nkX) ? lx[k*AMUL_TPB] : -1;
#pragma unroll
        for (int k = 0; k < AMUL_RC; k++) iC[k] = (k < nkC) ? lc[k*AMUL_TPB] : -1;
    };
    auto issue_gathers = [&](int tile, int stage)
    {
        const unsigned sb = sb0 + stage*stageBytes;
        const unsigned dx = sb + L.xv + 8u*(T + tid), dc = sb + L.cv + 8u*(EU + tid);
        if (MODE != 2)
        {
#pragma unroll
            for (int k = 0; k < AMUL_RX; k++)
                if (iX[k] >= 0) cp_async8(dx + 8u*k*AMUL_TPB, x + iX[k]);
            for (int k = AMUL_RX; k < nkX; k++)
            {
                const int i = A.amulExtX[(size_t)tile*A.amulXStride + tid + k*AMUL_TPB];
                if (i >= 0) cp_async8(dx + 8u*k*AMUL_TPB, x + i);
            }
        }
#pragma unroll
        for (int k = 0; k < AMUL_RC; k++)
            if (iC[k] >= 0) cp_async8(dc + 8u*k*AMUL_TPB, Uval + iC[k]);
        for (int k = AMUL_RC; k < nkC; k++)
        {
            const int i = A.amulExtC[(size_t)tile*A.amulCStride + tid + k*AMUL_TPB];
            if (i >= 0) cp_async8(dc + 8u*k*AMUL_TPB, Uval + i);
        }
        cp_async_commit();
    };

    const int t0 = blockIdx.x, G = gridDim.x;
    if (tid == 0)
    {
        mbar_init(sb0 + L.bar, 1);
        mbar_init(sb0 + stageBytes + L.bar, 1);
    }
    __syncthreads();
    AmulTileMeta Mcur = {0, 0, 0, 0, 0, 0, 0, 0}, Mnext = Mcur;
    if (t0 < A.nAmulTiles)
    {
        Mcur = A.amulMeta[t0];
        if (tid == 0) issue_bulk(Mcur, 0);
        load_indices(t0);
        issue_gathers(t0, 0);
        if (t0 + G < A.nAmulTiles)
        {
            Mnext = A.amulMeta[t0 + G];
            load_indices(t0 + G);
        }
    }
    double dot[1] = {0.0};
    unsigned phase[2] = {0u, 0u};
    int stage = 0;
    for (int tile = t0; tile < A.nAmulTiles; tile += G, stage ^= 1)
    {
        const int nextTile = tile + G;
        const bool haveNext = nextTile < A.nAmulTiles;
        cp_async_wait_all();                // my gathers of this tile have landed
        __syncthreads();                    // ... and everyone's; everyone is done reading buffer stage^1 (tile - G)
        AmulTileMeta M2 = Mnext;
        if (haveNext)
        {
            if (tid == 0) issue_bulk(Mnext, stage ^ 1);
            issue_gathers(nextTile, stage ^ 1);
            if (nextTile + G < A.nAmulTiles)
            {
                M2 = A.amulMeta[nextTile + G];
                load_indices(nextTile + G);
            }
        }
        mbar_wait(sb0 + stage*stageBytes + L.bar, phase[stage]);
        phase[stage] ^= 1u;

        // ---- compute tile out of buffer `stage`
        unsigned char* base = smemRaw + stage*stageBytes;
        const double* cv = (const double*)(base + L.cv);
        const double* dg = (const double*)(base + L.dg);
        double* xv = (double*)(base + L.xv);
        const unsigned short* U16 = (const unsigned short*)(base + L.U16);
        const unsigned int* L32 = (const unsigned int*)(base + L.L32);
        const SliceInfo* sl = (const SliceInfo*)(base + L.sl);
        const int r0 = Mcur.r0, nr = Mcur.nr;
        if (MODE == 2)
        {
            for (int q = tid; q < nr; q += AMUL_TPB) xv[q] = (rowOld[r0 + q] >= 0) ? xc : 0.0;
            for (int q = tid; q < A.amulXStride; q += AMUL_TPB) xv[T + q] = xc;
            __syncthreads();
        }
        for (int q = tid; q < nr; q += AMUL_TPB)
        {
            const SliceInfo si = sl[q >> 5];
            const int ln = q & 31;
            const unsigned int* pl = L32 + (si.offL - Mcur.eL0 + ln);
            const unsigned short* pu = U16 + (si.offU - Mcur.eU0 + ln);
            const double* cu = cv + (si.offU - Mcur.eU0 + ln);
            unsigned cl[4], cuv[4];
            double fL[4], xL[4], fU[4], xU[4];
#pragma unroll
            for (int j = 0; j < 4; j++) cl[j] = (j < si.widthL) ? pl[32*j] : 0x80000000u;
#pragma unroll
            for (int j = 0; j < 4; j++) cuv[j] = (j < si.widthU) ? (unsigned)pu[32*j] : 0x8000u;
            const double xr = xv[q];
            double acc = __dmul_rn(dg[q], xr);
#pragma unroll
            for (int j = 0; j < 4; j++)
            {
                fL[j] = cv[(cl[j] >> 16) & 0x7FFFu];
                xL[j] = xv[cl[j] & 0xFFFFu];
            }
#pragma unroll
            for (int j = 0; j < 4; j++)
            {
                fU[j] = (j < si.widthU) ? cu[32*j] : 0.0;
                xU[j] = xv[cuv[j] & 0x7FFFu];
            }
            // neighbour-side faces (u == row), ascending f
#pragma unroll
            for (int j = 0; j < 4; j++)
                if ((int)cl[j] >= 0) acc = __dadd_rn(acc, __dmul_rn(fL[j], xL[j]));
            for (int j = 4; j < si.widthL; j++)
            {
                const unsigned c = pl[32*j];
                if ((int)c >= 0) acc = __dadd_rn(acc, __dmul_rn(cv[(c >> 16) & 0x7FFFu], xv[c & 0xFFFFu]));
            }
            // owner-side faces (l == row), ascending f
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (!(cuv[j] & 0x8000u)) acc = __dadd_rn(acc, __dmul_rn(fU[j], xU[j]));
            for (int j = 4; j < si.widthU; j++)
            {
                const unsigned c = pu[32*j];
                if (!(c & 0x8000u)) acc = __dadd_rn(acc, __dmul_rn(cu[32*j], xv[c & 0x7FFFu]));
            }
            Ax[r0 + q] = acc;
            if (MODE == 1)
            {
                const int r = r0 + q;
                const bool masked = ifMask && ((ifMask[r >> 5] >> (r & 31)) & 1u);
                if (!masked) dot[0] += acc*xr;
            }
        }
        Mcur = Mnext;
        Mnext = M2;
    }
    // with coupled faces k_iface_update<1> adds their rows' part of the dot and closes the step
    if (MODE == 1) grid_reduce_finish<1, AMUL_TPB>(dot, partials, partStride, blockIdx.x, gridDim.x, st, S_AMUL, nullptr, 0, ifMask == nullptr);
}

#ifndef GPUPCG_SWEEP_TPB
#define GPUPCG_SWEEP_TPB 128
#endif
static constexpr int SWEEP_TPB = GPUPCG_SWEEP_TPB;       // one warp sweeps, the others fetch cross-tile values
static constexpr int FETCHERS = SWEEP_TPB - 32;

// Shared-memory carve-up of one sweep CTA (byte offsets, all 16 B aligned).
//  working set of the round loop:
//   vals  [NV]  doubles: [0,T) the tile's rows, [T,T+X) cross-tile inputs (SENT until a fetch warp delivers
//               them), [T+X] a constant 1.0 that unused entry slots point at with a zero coefficient, so the
//               round loop is branch-free (acc - 0*1 == acc and acc - 0/1 == acc exactly)
//   rowF  [4T]  doubles: the first four coefficients of every row
//   rowC  [4T]  ints   : their value indices
//   items [I+2] int2   : work items
//   ovf   [T]   bytes  : entries beyond the first four (walked from the staged SELL block; polyhedral cells only)
//  staging area (bulk copies land here; rowF/rowC are built from it):
//   Ub [E] doubles, col [E] u16, idx [EL] u16, dv [T] (rD or diag), av [T] (rA or y), cX [X] cross-tile coefficients
struct SweepLayout
{
    int T, X, NV, one;
    unsigned vals, rowF, rowC, items, ovf, slot, Ub, col, idx, dv, av, cX, sl, bar, bytes;
};

__host__ __device__ inline SweepLayout sweep_layout(int T, int X, int I, int E, int EL)
{
    SweepLayout L;
    L.T = T; L.X = X;
    L.NV = (T + X + 3) & ~1;
    L.one = T + X;
    L.vals = 0;
    L.rowF = 8u*L.NV;
    L.rowC = L.rowF + 32u*(T + 1);                  // +1: the dummy row idle lanes work on (coefficients 0)
    L.items = L.rowC + 16u*(T + 1);
    L.ovf = L.items + 8u*(I + 2);
    L.slot = L.ovf + ((T + 1 + 15) & ~15u);
    L.Ub = (L.slot + 16 + 15) & ~15u;
    L.dv = L.Ub + 8u*E;
    L.av = L.dv + 8u*T;
    L.cX = L.av + 8u*T;
    L.col = L.cX + 8u*((X + 1) & ~1);
    L.idx = L.col + 2u*E;
    L.sl = L.idx + 2u*EL;
    L.bar = L.sl + 16u*(T/32 + 1);
    L.bytes = L.bar + 16;
    return L;
}

// Table variant (meshes with <= 4 faces per row and side): the row tables {4 coefficients, 4 value offsets} are global
// arrays (indices static per mesh, coefficients pre-multiplied once per matrix), so staging is bulk copies only.
__host__ __device__ inline SweepLayout tab_layout(int T, int X, int I)
{
    SweepLayout L;
    L.T = T; L.X = X;
    L.NV = (T + X + 3) & ~1;
    L.one = T + X;
    L.vals = 0;
    L.rowF = 8u*L.NV;
    L.rowC = L.rowF + 32u*(T + 1);
    L.dv = L.rowC + 16u*(T + 1);
    L.av = L.dv + 8u*(T + 2);
    L.cX = L.av + 8u*(T + 2);               // forward sweep with the fused x/r update: the tile of wA = A pA
    L.items = L.cX + 8u*(T + 2);
    L.slot = L.items + 8u*(I + 2);
    L.bar = (L.slot + 16 + 15) & ~15u;
    L.bytes = L.bar + 16;
    L.ovf = L.Ub = L.col = L.idx = L.sl = 0;
    return L;
}

// shared-memory accessors on 32-bit shared addresses.  All volatile asm: the sweep loop is hand-scheduled and the
// compiler must keep the order (prefetches of the NEXT item first, then the dependency loads of the current one).
__device__ __forceinline__ double lds_f64(unsigned a)
{
    // ld.VOLATILE: this is the load the sweep warp polls cross-tile slots with.  A plain (weak) ld.shared lets
    // ptxas assume no other warp writes the slot and delete the re-poll loop (seen in SASS, r1).
    double v;
    asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
// weak load of a value slot: ordered against the other lanes' stores by the __syncwarp between rounds; for a
// cross-tile slot it returns either SENT or the final value (slots change exactly once), so it is a valid
// first probe.  (ld.volatile serialises: four of them cost ~120 cycles per round instead of ~30.)
__device__ __forceinline__ double lds_f64_weak(unsigned a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f64(unsigned a, double v)
{
    asm volatile("st.shared.f64 [%0], %1;" :: "r"(a), "d"(v));
}
__device__ __forceinline__ void sts_u64(unsigned a, unsigned long long v)
{
    asm volatile("st.volatile.shared.u64 [%0], %1;" :: "r"(a), "l"(v) : "memory");
}
__device__ __forceinline__ int4 lds_v4(unsigned a)
{
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ double2 lds_d2(unsigned a)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ int2 lds_i2(unsigned a)
{
    int2 v;
    asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ int lds_u8(unsigned a)
{
    int v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

// value index of a 16-bit tile column code: local row, cross-tile slot, or the constant slot for a pad
__device__ __forceinline__ int value_index16(unsigned code, int T, int one)
{
    return code == 0xFFFFu ? one : ((code & 0x8000u) ? T + (int)(code & 0x7FFFu) : (int)code);
}

__device__ __forceinline__ bool is_sent(double v)
{
    return (unsigned long long)__double_as_longlong(v) == SENT;
}

// Static data of one row, fetched into registers one work item ahead of its use so that the only loads left on
// the round-to-round dependency chain are the four value loads.  rowC holds BYTE offsets into the value array.
// Lanes without a row in a pass work on the dummy row T (zero coefficients, constant-slot dependencies) and store
// into the scratch slot one+1: no predication in the loop except the global publish.
struct RowRegs
{
    unsigned q8;        // byte offset of the row's value slot (scratch slot for idle lanes)
    int qg;             // tile-local row for the global publish, -1 = idle lane
    int4 c;             // byte offsets of the four dependency values
    double f0, f1, f2, f3;
    double acc0;
    int ovf;
};

template <bool HAS_OVF>
__device__ __forceinline__ RowRegs load_row(unsigned sb, const SweepLayout& L, int2 item, int lane)
{
    RowRegs R;
    const bool active = lane < (item.y & 0xffff);
    const int q = active ? item.x + lane : L.T;
    R.qg = active ? q : -1;
    R.q8 = active ? 8u*q : 8u*(L.one + 1);
    R.c = lds_v4(sb + L.rowC + 16u*q);
    const double2 fa = lds_d2(sb + L.rowF + 32u*q);
    const double2 fb = lds_d2(sb + L.rowF + 32u*q + 16u);
    R.acc0 = lds_f64_weak(sb + L.vals + R.q8);
    R.ovf = HAS_OVF ? lds_u8(sb + L.ovf + q) : 0;
    R.f0 = fa.x; R.f1 = fa.y; R.f2 = fb.x; R.f3 = fb.y;
    return R;
}

// table variant: the start value is formed here, one item ahead of its use (MODE 0: rD*rA, 1: diag, 2: y)
template <int MODE>
__device__ __forceinline__ RowRegs load_row_tab(unsigned sb, const SweepLayout& L, int2 item, int lane)
{
    RowRegs R;
    const bool active = lane < (item.y & 0xffff);
    const int q = active ? item.x + lane : L.T;
    R.qg = active ? q : -1;
    R.q8 = active ? 8u*q : 8u*(L.one + 1);
    R.c = lds_v4(sb + L.rowC + 16u*q);
    const double2 fa = lds_d2(sb + L.rowF + 32u*q);
    const double2 fb = lds_d2(sb + L.rowF + 32u*q + 16u);
    if (MODE == 0) R.acc0 = __dmul_rn(lds_f64_weak(sb + L.dv + 8u*q), lds_f64_weak(sb + L.av + 8u*q));
    else if (MODE == 1) R.acc0 = lds_f64_weak(sb + L.dv + 8u*q);
    else R.acc0 = lds_f64_weak(sb + L.av + 8u*q);
    R.acc0 = (R.acc0 != R.acc0) ? __longlong_as_double((long long)QNAN) : R.acc0;
    R.ovf = 0;
    R.f0 = fa.x; R.f1 = fa.y; R.f2 = fb.x; R.f3 = fb.y;
    return R;
}

// the four register-resident dependencies of a row: four loads issued together, one combined readiness test.
// Rows of the tile are final by construction of the rounds; cross-tile slots may still hold SENT.
__device__ __forceinline__ void dep_values4(unsigned vb, const int4 c, double& v0, double& v1, double& v2, double& v3)
{
    v0 = lds_f64_weak(vb + c.x); v1 = lds_f64_weak(vb + c.y);
    v2 = lds_f64_weak(vb + c.z); v3 = lds_f64_weak(vb + c.w);
    while (is_sent(v0) | is_sent(v1) | is_sent(v2) | is_sent(v3))
    {
        v0 = lds_f64(vb + c.x); v1 = lds_f64(vb + c.y); v2 = lds_f64(vb + c.z); v3 = lds_f64(vb + c.w);
    }
}

__device__ __forceinline__ double dep_value(unsigned vb, int idx)
{
    double v = lds_f64(vb + 8u*idx);
    while (is_sent(v)) v = lds_f64(vb + 8u*idx);
    return v;
}

// ------------------------------------------------------------------------------------------
// K9 / K10.  One CTA sweeps one tile.
//   MODE 0: forward sweep of DICPreconditioner::precondition  [EXT DICPreconditioner.C]
//           y[row] = rD[row]*rA[row] - sum_{(l,row) asc f} (rD[row]*upper[f])*y[l]        out = y (SENT on entry)
//   MODE 1: DICPreconditioner::calcReciprocalD
//           d[row] = diag[row] - sum_{(l,row) asc f} upper[f]*upper[f]/d[l];  rD = 1/d     out = pivots (scratch)
//   MODE 2: backward sweep, tiles and rounds in reverse order
//           z[row] = y[row] - sum_{(row,u) desc f} (rD[row]*upper[f])*z[u]                 out = z (= wA)
//           y is complete (previous kernel) and is re-armed with SENT here for the next forward sweep; z is
//           re-armed by k_xr_update.  wArA = sum(wA*rA) is fused: one partial per tile, summed in a fixed order.
// Staging: thread 0 issues the bulk copies of the tile's blocks (coefficients, 16-bit codes, rD/diag, rA/y) onto
// one mbarrier while all threads gather the cross-tile coefficients; then per-row {4 coefficients, 4 value
// indices} and the start values are built in shared memory.  Then warp 0 walks the work items (a round = one
// global dependency level present in the tile, cut into passes of 32 rows) with __syncwarp between rounds, while
// warps 1-3 poll the cross-tile inputs from global memory, in consumption order, into the value array.
// ------------------------------------------------------------------------------------------
// a NaN that reaches the value array must not look like SENT: inputs are canonicalised once at staging (arithmetic
// on canonical NaNs yields canonical NaNs), so the loop publishes values as they are
__device__ __forceinline__ double canon(double v)
{
    return (v != v) ? __longlong_as_double((long long)QNAN) : v;
}

// The x/r update of PCG::solve folded into the staging of the NEXT forward sweep (table variant only): the sweep is
// latency-bound, its CTAs have bandwidth to spare, and the update's five streams ride along for free.
//   x += alpha*pA;  rA -= alpha*wA;  sum|rA| -> S_XR (residual, stop test, history);  wA re-armed with SENT
struct FuseXR
{
    double* x;                  // null: plain sweep
    const double* pA;
    unsigned long long* wA;
    double* rA;
    double* history;
};

// TAB: table variant (tabC / tabF = the row tables of this sweep direction; Uval unused)
template <int MODE, bool HAS_OVF, bool TAB = false>
__global__ void __launch_bounds__(SWEEP_TPB) k_dic_tile(TileArgs A, const double* __restrict__ diagOrRD,
                                                        const double* __restrict__ Uval,
                                                        const double* __restrict__ rA, unsigned long long* y,
                                                        unsigned long long* out, double* __restrict__ rDout,
                                                        unsigned long long* tileTicket, double* partials, int partStride,
                                                        DevState* st, int gated,
                                                        const int4* __restrict__ tabC = nullptr, const double* __restrict__ tabF = nullptr,
                                                        FuseXR fuse = FuseXR{nullptr, nullptr, nullptr, nullptr, nullptr})
{
    if (gated && st->done) return;
    const bool fused = TAB && MODE == 0 && fuse.x != nullptr;
    const bool doXR = fused && st->pending;             // uniform over the grid: S_XR clears it after the last CTA
    const double alphaXR = doXR ? __ddiv_rn(st->wArA, st->wApA) : 0.0;
    double sumXR[1] = {0.0};
    static_assert(!(TAB && HAS_OVF), "the table variant carries four faces per row and side");
    constexpr bool BWD = (MODE == 2), FACTOR = (MODE == 1);
    extern __shared__ __align__(128) unsigned char smemRaw[];
    const int X = BWD ? A.maxExtB : A.maxExtF;
    const SweepLayout L = TAB ? tab_layout(A.tileRowsMax, X, A.maxItems)
                              : sweep_layout(A.tileRowsMax, X, A.maxItems, BWD ? A.maxEU : max(A.maxEU, A.maxEL), A.maxEL);
    const unsigned sb = (unsigned)__cvta_generic_to_shared(smemRaw);
    const unsigned vb = sb + L.vals;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long tBegin = (g_tileTimes && threadIdx.x == 0) ? gtime() : 0ULL;
    const int tile = take_tile(tileTicket, A.nTiles, (int*)(smemRaw + L.slot), BWD);
    const int r0 = A.tileStart[tile], nr = A.tileStart[tile + 1] - r0;
    const int s0 = r0 >> 5, nsl = nr >> 5;
    const int ip = A.tileItemPtr[tile], nItems = A.tileItemPtr[tile + 1] - ip;
    const int xBase = BWD ? A.extBPtr[tile] : A.extFPtr[tile];
    const int nExt = (BWD ? A.extBPtr[tile + 1] : A.extFPtr[tile + 1]) - xBase;
    const int* extCol = BWD ? A.extBCol : A.extFCol;
    const int T = L.T;
    int eU0 = 0, nEU = 0, eL0 = 0, nEL = 0;
    if (!TAB)
    {
        const SliceInfo first = A.slices[s0], last = A.slices[s0 + nsl - 1];
        eU0 = first.offU; nEU = last.offU + 32*last.widthU - eU0;
        eL0 = first.offL; nEL = last.offL + 32*last.widthL - eL0;
    }
    double* vals = (double*)(smemRaw + L.vals);
    double* cX = (double*)(smemRaw + L.cX);
    const double* Ub = (const double*)(smemRaw + L.Ub);
    const double* dv = (const double*)(smemRaw + L.dv);
    const unsigned short* colS = (const unsigned short*)(smemRaw + L.col);
    const unsigned short* idxS = (const unsigned short*)(smemRaw + L.idx);
    const SliceInfo* sl = (const SliceInfo*)(smemRaw + L.sl);

    if (TAB)
    {
        // ---- staging, table variant: five bulk copies, nothing to build
        if (threadIdx.x == 0)
        {
            mbar_init(sb + L.bar, 1);
            const unsigned bytes = 32u*nr + 16u*nr + (BWD ? 0u : 8u*nr) + (FACTOR ? 0u : 8u*nr) + (doXR ? 8u*nr : 0u);
            mbar_expect_tx(sb + L.bar, bytes);
            bulk_g2s(sb + L.rowF, tabF + 4*(size_t)r0, 32u*nr, sb + L.bar);
            bulk_g2s(sb + L.rowC, tabC + r0, 16u*nr, sb + L.bar);
            if (!BWD) bulk_g2s(sb + L.dv, diagOrRD + r0, 8u*nr, sb + L.bar);
            if (!FACTOR) bulk_g2s(sb + L.av, BWD ? (const void*)(y + r0) : (const void*)(rA + r0), 8u*nr, sb + L.bar);
            if (doXR) bulk_g2s(sb + L.cX, fuse.wA + r0, 8u*nr, sb + L.bar);
        }
        int2* items = (int2*)(smemRaw + L.items);
        for (int i = threadIdx.x; i < nItems + 2; i += SWEEP_TPB)
            items[i] = (i < nItems) ? A.items[ip + (BWD ? nItems - 1 - i : i)] : make_int2(0, 0);
        for (int i = threadIdx.x; i < nExt; i += SWEEP_TPB) ((unsigned long long*)vals)[T + i] = SENT;
        if (threadIdx.x < 4)            // the dummy row of idle lanes (never touched by the bulk copies: nr <= T)
        {
            ((int*)(smemRaw + L.rowC))[4*T + threadIdx.x] = 8*L.one;
            ((double*)(smemRaw + L.rowF))[4*T + threadIdx.x] = 0.0;
            if (threadIdx.x == 0)
            {
                vals[L.one] = 1.0; vals[L.one + 1] = 0.0;
                ((double*)(smemRaw + L.dv))[T] = 0.0; ((double*)(smemRaw + L.av))[T] = 0.0;
            }
        }
        __syncthreads();
        mbar_wait(sb + L.bar, 0);
        if (BWD && warp != 0)           // y has been read: re-arm it for the next forward sweep
            for (int q = threadIdx.x - 32; q < nr; q += FETCHERS) y[r0 + q] = SENT;
        if (doXR)
        {
            double* avw = (double*)(smemRaw + L.av);
            const double* bv = (const double*)(smemRaw + L.cX);
            for (int q = threadIdx.x; q < nr; q += SWEEP_TPB)
            {
                const int r = r0 + q;
                fuse.x[r] = __dadd_rn(fuse.x[r], __dmul_rn(alphaXR, fuse.pA[r]));
                const double rr = __dsub_rn(avw[q], __dmul_rn(alphaXR, bv[q]));
                avw[q] = rr;
                fuse.rA[r] = rr;
                sumXR[0] += fabs(rr);
                fuse.wA[r] = SENT;
            }
            __syncthreads();            // the sweep warp forms its start values from the updated residual
        }
    }
    else
    {
    // ---- staging: bulk copies (one mbarrier) + cross-tile coefficient gathers
    if (threadIdx.x == 0)
    {
        mbar_init(sb + L.bar, 1);
        unsigned bytes = 8u*nEU + 8u*nr + 16u*nsl + (FACTOR ? 0u : 8u*nr) + (BWD ? 2u*nEU : 4u*nEL);
        mbar_expect_tx(sb + L.bar, bytes);
        bulk_g2s(sb + L.Ub, Uval + eU0, 8u*nEU, sb + L.bar);
        bulk_g2s(sb + L.dv, diagOrRD + r0, 8u*nr, sb + L.bar);
        if (!FACTOR) bulk_g2s(sb + L.av, BWD ? (const void*)(y + r0) : (const void*)(rA + r0), 8u*nr, sb + L.bar);
        bulk_g2s(sb + L.sl, A.slices + s0, 16u*nsl, sb + L.bar);
        if (BWD) bulk_g2s(sb + L.col, A.colU16 + eU0, 2u*nEU, sb + L.bar);
        else
        {
            bulk_g2s(sb + L.col, A.colL16 + eL0, 2u*nEL, sb + L.bar);
            bulk_g2s(sb + L.idx, A.idxL16 + eL0, 2u*nEL, sb + L.bar);
        }
    }
    if (!BWD)
    {
        for (int k = threadIdx.x; k < nExt; k += SWEEP_TPB) cX[k] = Uval[A.extFIdx[xBase + k]];
    }
    {
        int2* items = (int2*)(smemRaw + L.items);
        for (int i = threadIdx.x; i < nItems + 2; i += SWEEP_TPB)
            items[i] = (i < nItems) ? A.items[ip + (BWD ? nItems - 1 - i : i)] : make_int2(0, 0);
        for (int i = threadIdx.x; i < nExt; i += SWEEP_TPB) ((unsigned long long*)vals)[T + i] = SENT;
        if (threadIdx.x == 0) vals[L.one] = 1.0;
    }
    __syncthreads();
    mbar_wait(sb + L.bar, 0);

    // ---- build the round loop's working set from the staged blocks
    const double* av = (const double*)(smemRaw + L.av);
    double* rowF = (double*)(smemRaw + L.rowF);
    int* rowC = (int*)(smemRaw + L.rowC);
    unsigned char* ovf = smemRaw + L.ovf;
    for (int q = threadIdx.x; q < nr; q += SWEEP_TPB)
    {
        const SliceInfo si = sl[q >> 5];
        const int wid = BWD ? si.widthU : si.widthL;
        const int off = (BWD ? si.offU - eU0 : si.offL - eL0) + (q & 31);
        const double d = dv[q];
        if (BWD) { vals[q] = canon(av[q]); y[r0 + q] = SENT; }
        else vals[q] = canon(FACTOR ? d : __dmul_rn(d, av[q]));
        int nOver = 0;
#pragma unroll
        for (int j = 0; j < 4; j++)
        {
            unsigned code = 0xFFFFu;
            double cf = 0.0;
            if (j < wid)
            {
                const int e = off + j*32;
                code = colS[e];
                if (code != 0xFFFFu)
                {
                    const double c = BWD ? Ub[e] : ((code & 0x8000u) ? cX[code & 0x7FFFu] : Ub[idxS[e]]);
                    cf = canon(FACTOR ? __dmul_rn(c, c) : __dmul_rn(d, c));
                }
            }
            rowC[4*q + j] = 8*value_index16(code, T, L.one);
            rowF[4*q + j] = cf;
        }
        for (int j = 4; j < wid; j++) nOver += (colS[off + j*32] != 0xFFFFu);
        ovf[q] = (unsigned char)nOver;
    }
    if (threadIdx.x < 4)        // the dummy row of idle lanes
    {
        rowC[4*T + threadIdx.x] = 8*L.one;
        rowF[4*T + threadIdx.x] = 0.0;
        if (threadIdx.x == 0) { ovf[T] = 0; vals[L.one + 1] = 0.0; }
    }
    __syncthreads();
    }
    if (g_tileTimes && threadIdx.x == 0 && MODE == 0) { g_tileTimes[3*tile] = tBegin; g_tileTimes[3*tile + 1] = gtime(); }

    // (tried, no effect on B200: rotating the sweep-warp role over the SM's schedulers, sleeping between shared
    //  polls, gating / doubling / tripling the global polls of the fetch lanes -- profiles/r1c_dic_latency_study.md)
    if (warp == 0)
    {
        // Software pipeline over the work items: in iteration i the descriptor of item i+2 and the static row data
        // of item i+1 are requested first (consumed in later iterations); then the only wait of the iteration:
        // the four dependency values of item i.  The item table was staged in sweep order (reversed for MODE 2).
        const unsigned ib = sb + L.items;
        const int syncBit = BWD ? (1 << 16) : (1 << 17);       // end of a round in sweep order
        int2 d0 = lds_i2(ib), d1 = lds_i2(ib + 8u);
        RowRegs cur = TAB ? load_row_tab<MODE>(sb, L, d0, lane) : load_row<HAS_OVF>(sb, L, d0, lane);
        for (int i = 0; i < nItems; i++)
        {
            const int2 d2 = lds_i2(ib + 8u*(i + 2));
            const RowRegs nxt = TAB ? load_row_tab<MODE>(sb, L, d1, lane) : load_row<HAS_OVF>(sb, L, d1, lane);
            double v0, v1, v2, v3;
            dep_values4(vb, cur.c, v0, v1, v2, v3);
            double acc = cur.acc0;
            if (BWD)
            {
                if (HAS_OVF && cur.ovf)        // more than four owner-side faces: the tail (descending) comes first
                {
                    const SliceInfo si = sl[cur.qg >> 5];
                    const int off = si.offU - eU0 + (cur.qg & 31);
                    const double d = dv[cur.qg];
                    for (int j = si.widthU - 1; j >= 4; j--)
                    {
                        const unsigned code = colS[off + j*32];
                        if (code == 0xFFFFu) continue;
                        acc = __dsub_rn(acc, __dmul_rn(__dmul_rn(d, Ub[off + j*32]), dep_value(vb, value_index16(code, T, L.one))));
                    }
                }
                const double t0 = __dmul_rn(cur.f0, v0), t1 = __dmul_rn(cur.f1, v1);
                const double t2 = __dmul_rn(cur.f2, v2), t3 = __dmul_rn(cur.f3, v3);
                acc = __dsub_rn(__dsub_rn(__dsub_rn(__dsub_rn(acc, t3), t2), t1), t0);
            }
            else
            {
                double t0, t1, t2, t3;
                if (FACTOR)
                {
                    // unused slots (coefficient 0, value 1) are exact no-ops: skip their ~250-cycle IEEE divisions
                    t0 = (cur.c.x != 8*L.one) ? __ddiv_rn(cur.f0, v0) : 0.0;
                    t1 = (cur.c.y != 8*L.one) ? __ddiv_rn(cur.f1, v1) : 0.0;
                    t2 = (cur.c.z != 8*L.one) ? __ddiv_rn(cur.f2, v2) : 0.0;
                    t3 = (cur.c.w != 8*L.one) ? __ddiv_rn(cur.f3, v3) : 0.0;
                }
                else
                {
                    t0 = __dmul_rn(cur.f0, v0); t1 = __dmul_rn(cur.f1, v1);
                    t2 = __dmul_rn(cur.f2, v2); t3 = __dmul_rn(cur.f3, v3);
                }
                acc = __dsub_rn(__dsub_rn(__dsub_rn(__dsub_rn(acc, t0), t1), t2), t3);
                if (HAS_OVF && cur.ovf)        // more than four neighbour-side faces: tail from the staged SELL block
                {
                    const SliceInfo si = sl[cur.qg >> 5];
                    const int off = si.offL - eL0 + (cur.qg & 31);
                    const double d = dv[cur.qg];
                    for (int j = 4; j < si.widthL; j++)
                    {
                        const unsigned code = colS[off + j*32];
                        if (code == 0xFFFFu) break;
                        const double c = (code & 0x8000u) ? cX[code & 0x7FFFu] : Ub[idxS[off + j*32]];
                        const double vv = dep_value(vb, value_index16(code, T, L.one));
                        acc = FACTOR ? __dsub_rn(acc, __ddiv_rn(__dmul_rn(c, c), vv))
                                     : __dsub_rn(acc, __dmul_rn(__dmul_rn(d, c), vv));
                    }
                }
            }
            sts_f64(vb + cur.q8, acc);
            if (cur.qg >= 0)
                asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" :: "l"(out + r0 + cur.qg), "d"(acc) : "memory");
#ifdef GPUPCG_DEEPTRACE
            if (MODE == 0 && g_tileTimes && cur.qg >= 0) g_tileTimes[3*A.nTiles + r0 + cur.qg] = gtime();
#endif
            if (d0.y & syncBit) __syncwarp();
            cur = nxt; d0 = d1; d1 = d2;
        }
        if (g_tileTimes && lane == 0 && MODE == 0) g_tileTimes[3*tile + 2] = gtime();
    }
    else
    {
        // Fetch lanes are independent state machines: each polls ITS current entry and delivers it the moment it
        // arrives.  (A plain "wait, then store" loop reconverges the warp before the store, so an early entry was
        // held back until the latest of its 32 warp-mates had arrived: 1.7 us per tile hop instead of ~1.)
        constexpr int stride = FETCHERS;
        int i = (warp - 1)*32 + lane;
        const int ns = g_spinNs;
        const unsigned long long* p = (i < nExt) ? out + extCol[xBase + i] : nullptr;
        // the address of the lane's NEXT entry is fetched while it waits for the current one: advancing must not put
        // a global-load latency into the poll loop of the whole warp
        int colNext = (i + stride < nExt) ? extCol[xBase + i + stride] : 0;
        SpinWatch watch;
        bool bail = false;
        while (i < nExt)
        {
            const unsigned long long v = bail ? 0ULL : peek64(p);
            if (v != SENT)
            {
                watch.progress();
                sts_u64(vb + 8u*(T + i), v);
#ifdef GPUPCG_DEEPTRACE
                if (MODE == 0 && g_tileTimes) g_tileTimes[3*A.nTiles + A.tileStart[A.nTiles] + xBase + i] = gtime();
#endif
                i += stride;
                p = out + colNext;
                if (i + stride < nExt) colNext = extCol[xBase + i + stride];
            }
            else if (watch.expired(SPIN_TILE_FETCH)) bail = true;
            else if (ns > 0) __nanosleep(ns);
        }
    }
    if (fused)      // every CTA of a fused launch takes part (S_XR is a no-op when nothing was pending)
        grid_reduce_finish<1, SWEEP_TPB>(sumXR, partials, partStride, tile, A.nTiles, st, S_XR, fuse.history);
    if (FACTOR)
    {
        __syncthreads();
        for (int q = threadIdx.x; q < nr; q += SWEEP_TPB) rDout[r0 + q] = __ddiv_rn(1.0, vals[q]);
    }
    if (BWD && partials)
    {
        __syncthreads();
        double dot = 0.0;
        for (int q = threadIdx.x; q < nr; q += SWEEP_TPB) dot += vals[q]*rA[r0 + q];
        double v[1] = {dot};
        grid_reduce_finish<1, SWEEP_TPB>(v, partials, partStride, tile, A.nTiles, st, S_SWEEP, nullptr);
    }
}

// Row-table coefficients of the table variant.  WHICH 0 (before the factor): c*c per neighbour-side face;
// WHICH 1 (after it): rD[row]*c -- the products the reference forms first in  rD[u]*upper[f]*wA[l]  (left to right).
template <int WHICH>
__global__ void __launch_bounds__(TPB) k_tab_coeffs(int nRows, const int4* __restrict__ idx, const double* __restrict__ Uval,
                                                    const double* __restrict__ rD, double* __restrict__ F,
                                                    const DevState* __restrict__ st, int gated)
{
    if (gated && st->done) return;
    for (int r = blockIdx.x*TPB + threadIdx.x; r < nRows; r += gridDim.x*TPB)
    {
        const int4 i4 = idx[r];
        const int ii[4] = {i4.x, i4.y, i4.z, i4.w};
        const double d = (WHICH == 1) ? rD[r] : 0.0;
        double f[4];
#pragma unroll
        for (int j = 0; j < 4; j++)
        {
            f[j] = 0.0;
            if (ii[j] >= 0)
            {
                const double c = Uval[ii[j]];
                f[j] = canon((WHICH == 0) ? __dmul_rn(c, c) : __dmul_rn(d, c));
            }
        }
        double2* o = (double2*)(F + 4*(size_t)r);
        o[0] = make_double2(f[0], f[1]);
        o[1] = make_double2(f[2], f[3]);
    }
}

// ------------------------------------------------------------------------------------------
// K11 BLAS-1 of PCG::solve
// ------------------------------------------------------------------------------------------
// rA = b - wA;  sum(x)  (for gAverage)
__global__ void __launch_bounds__(TPB) k_init_residual(const double* __restrict__ b, const double* __restrict__ wA,
                                                       const double* __restrict__ x, double* __restrict__ rA,
                                                       int nRows, double* partials, DevState* st)
{
    double v[1] = {0.0};
    for (int r = blockIdx.x*TPB + threadIdx.x; r < nRows; r += gridDim.x*TPB)
    {
        rA[r] = __dsub_rn(b[r], wA[r]);
        v[0] += x[r];
    }
    grid_reduce_finish<1, TPB>(v, partials, MAXPART, blockIdx.x, gridDim.x, st, S_INIT, nullptr);
}

// normFactor = sum(|Ax - t| + |b - t|) + small_ ;  sum|rA|
__global__ void __launch_bounds__(TPB) k_norm(const double* __restrict__ wA, const double* __restrict__ t,
                                              const double* __restrict__ b, const double* __restrict__ rA,
                                              int nRows, double* partials, DevState* st)
{
    double v[2] = {0.0, 0.0};
    for (int r = blockIdx.x*TPB + threadIdx.x; r < nRows; r += gridDim.x*TPB)
    {
        const double tr = t[r];
        v[0] += __dadd_rn(fabs(__dsub_rn(wA[r], tr)), fabs(__dsub_rn(b[r], tr)));
        v[1] += fabs(rA[r]);
    }
    grid_reduce_finish<2, TPB>(v, partials, MAXPART, blockIdx.x, gridDim.x, st, S_NORM, nullptr);
}

// pA = wA (first iteration) | wA + beta*pA,  beta = wArA/wArAold
__global__ void __launch_bounds__(TPB) k_p_update(const double* __restrict__ wA, double* __restrict__ pA,
                                                  int nRows, const DevState* __restrict__ st)
{
    if (st->done) return;
    const bool firstIter = (st->nIter == 0);
    const double beta = firstIter ? 0.0 : __ddiv_rn(st->wArA, st->wArAold);
    for (int r = blockIdx.x*TPB + threadIdx.x; r < nRows; r += gridDim.x*TPB)
    {
        const double w = wA[r];
        pA[r] = firstIter ? w : __dadd_rn(w, __dmul_rn(beta, pA[r]));
    }
}

// x += alpha*pA;  rA -= alpha*wA;  sum|rA|;  wA re-armed with SENT for the next backward sweep
__global__ void __launch_bounds__(TPB) k_xr_update(const double* __restrict__ pA, unsigned long long* wAbits,
                                                   double* __restrict__ x, double* __restrict__ rA,
                                                   int nRows, double* partials, DevState* st, double* history)
{
    if (!st->pending) return;       // nothing to apply: converged / singular before this update, or already applied (fused)
    const double alpha = __ddiv_rn(st->wArA, st->wApA);
    double v[1] = {0.0};
    for (int r = blockIdx.x*TPB + threadIdx.x; r < nRows; r += gridDim.x*TPB)
    {
        const double w = __longlong_as_double((long long)wAbits[r]);
        x[r] = __dadd_rn(x[r], __dmul_rn(alpha, pA[r]));
        const double rr = __dsub_rn(rA[r], __dmul_rn(alpha, w));
        rA[r] = rr;
        v[0] += fabs(rr);
        wAbits[r] = SENT;
    }
    grid_reduce_finish<1, TPB>(v, partials, MAXPART, blockIdx.x, gridDim.x, st, S_XR, history);
}

// L2 flush helper for the roofline timings
__global__ void __launch_bounds__(TPB) k_flush(float* __restrict__ p, long long n, float v)
{
    for (long long i = (long long)blockIdx.x*TPB + threadIdx.x; i < n; i += (long long)gridDim.x*TPB) p[i] = v;
}

} // namespace gpupcg
