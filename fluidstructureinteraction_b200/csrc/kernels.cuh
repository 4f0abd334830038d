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

namespace gpupcg
{

// "not yet written" marker of the dataflow sweeps.  A quiet NaN with a payload that arithmetic
// cannot produce: every published value is canonicalised (publish()) if it is a NaN.
static constexpr unsigned long long SENT = 0xFFF8B200DEADC0DEULL;
static constexpr unsigned long long QNAN = 0x7FF8000000000000ULL;

static constexpr int TPB = 256;             // threads per block of every kernel
static constexpr int WPB = TPB/32;          // warps per block
static constexpr int MAXPART = 4096;        // max CTAs of any reducing kernel

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
    unsigned int pad_;
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
            }
            break;
        case S_XR:      // finalResidual = gSumMag(rA)/normFactor; nIterations++; stop()
            st->finalResidual = __ddiv_rn(v[0], st->normFactor);
            if (history && st->nIter < st->historyCap) history[st->nIter] = st->finalResidual;
            st->nIter++;
            if (dev_stop(st)) st->done = 1;
            break;
    }
}

// ------------------------------------------------------------------------------------------
// deterministic grid reduction: block tree -> partials[block] -> the last block to finish sums the
// partials in a fixed order and either advances the scalar recurrence (single rank) or parks the
// local sums in st->red for the all-reduce (multi rank).
// ------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void grid_reduce_finish(double (&v)[NV], double* partials, DevState* st,
                                                   int step, double* history)
{
    __shared__ double sm[NV][WPB];
    __shared__ bool amLast;
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
            for (int w = 1; w < WPB; w++) a += sm[k][w];
            __stcg(&partials[k*MAXPART + blockIdx.x], a);
        }
        __threadfence();
        const unsigned int t = atomicAdd(&st->ticket, 1u);
        amLast = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!amLast) return;
    __threadfence();
    // fixed-shape final sum: thread t takes partials t, t+TPB, ... then the same block tree
    double f[NV];
#pragma unroll
    for (int k = 0; k < NV; k++)
    {
        double a = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += TPB) a += __ldcg(&partials[k*MAXPART + i]);
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
            for (int w = 1; w < WPB; w++) a += sm[k][w];
            f[k] = a;
        }
        st->ticket = 0;
        if (st->multiRank)
        {
#pragma unroll
            for (int k = 0; k < NV; k++) st->red[k] = f[k];
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
    if (DOT) grid_reduce_finish<1>(dot, partials + 2*MAXPART, st, S_AMUL, nullptr);
}

// ------------------------------------------------------------------------------------------
// K8  lduMatrix::Amul  [EXT lduMatrixATmul.C], row-gather form.
// MODE 0: Ax = A x                      (test entry, first residual)
// MODE 1: Ax = A x and  sum(Ax*x)       (PCG: wA = A pA, wApA = gSumProd(wA, pA)); gated on st->done
// MODE 2: Ax = A (xRef field)           (normFactor)
// ifMask (may be null): bit k of ifMask[s] set = row owns coupled faces; such rows are left out of
// the MODE-1 dot here and added by k_iface_update once the halo has arrived.
// ------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(TPB) k_amul(const SliceInfo* __restrict__ slices, int nSlices,
                                              const double* __restrict__ diag, const double* __restrict__ Uval,
                                              const int* __restrict__ Ucol, const int* __restrict__ Lidx,
                                              const int* __restrict__ Lcol, const int* __restrict__ rowOld,
                                              const unsigned int* __restrict__ ifMask,
                                              const double* __restrict__ x, double* __restrict__ Ax,
                                              double* partials, DevState* st)
{
    if (MODE == 1 && st->done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double xc = (MODE == 2) ? st->xRef : 0.0;
    double dot[1] = {0.0};
    for (int s = blockIdx.x + gridDim.x*warp; s < nSlices; s += gridDim.x*WPB)
    {
        const SliceInfo si = slices[s];
        const int r = s*32 + lane;
        double xr;
        if (MODE == 2) xr = (rowOld[r] >= 0) ? xc : 0.0; else xr = x[r];
        double acc = __dmul_rn(diag[r], xr);
        // neighbour-side faces (u == row), ascending f
#pragma unroll 4
        for (int j = 0; j < si.widthL; j++)
        {
            const int e = si.offL + j*32 + lane;
            const int col = Lcol[e];
            if (col >= 0)
            {
                const double c = Uval[Lidx[e]];
                const double xv = (MODE == 2) ? xc : x[col];
                acc = __dadd_rn(acc, __dmul_rn(c, xv));
            }
        }
        // owner-side faces (l == row), ascending f
#pragma unroll 4
        for (int j = 0; j < si.widthU; j++)
        {
            const int e = si.offU + j*32 + lane;
            const int col = Ucol[e];
            if (col >= 0)
            {
                const double c = Uval[e];
                const double xv = (MODE == 2) ? xc : x[col];
                acc = __dadd_rn(acc, __dmul_rn(c, xv));
            }
        }
        Ax[r] = acc;
        if (MODE == 1)
        {
            const bool masked = ifMask && ((ifMask[s] >> lane) & 1u);
            if (!masked) dot[0] += acc*xr;
        }
    }
    if (MODE == 1) grid_reduce_finish<1>(dot, partials, st, S_AMUL, nullptr);
}

// ------------------------------------------------------------------------------------------
// dataflow primitives of the wavefront sweeps: the value IS the flag.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long peek64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ double wait_value(const unsigned long long* p, unsigned long long first)
{
    unsigned long long v = first;
    while (v == SENT)
    {
        __nanosleep(20);
        v = peek64(p);
    }
    return __longlong_as_double((long long)v);
}

__device__ __forceinline__ void publish(unsigned long long* p, double v)
{
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    if (v != v) b = QNAN;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(b) : "memory");
}

// ------------------------------------------------------------------------------------------
// K9  DICPreconditioner::calcReciprocalD  [EXT DICPreconditioner.C], gather + dataflow form:
//   d[row] = diag[row] - sum_{faces (l,row), ascending f} upper[f]*upper[f]/d[l];   rD = 1/d
// dscratch must hold SENT on entry; it receives the un-inverted pivots.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_dic_factor(const SliceInfo* __restrict__ slices, int nSlices,
                                                    const double* __restrict__ diag, const double* __restrict__ Uval,
                                                    const int* __restrict__ Lidx, const int* __restrict__ Lcol,
                                                    unsigned long long* dscratch, double* __restrict__ rD, DevState* st,
                                                    int gated)
{
    if (gated && st->done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int s = blockIdx.x + gridDim.x*warp; s < nSlices; s += gridDim.x*WPB)
    {
        const SliceInfo si = slices[s];
        const int r = s*32 + lane;
        double d = diag[r];
        for (int j0 = 0; j0 < si.widthL; j0 += 4)
        {
            int col[4]; double c[4]; unsigned long long dep[4];
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
                col[k] = -1;
                if (j0 + k < si.widthL)
                {
                    const int e = si.offL + (j0 + k)*32 + lane;
                    col[k] = Lcol[e];
                    if (col[k] >= 0) { c[k] = Uval[Lidx[e]]; dep[k] = peek64(dscratch + col[k]); }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
                if (col[k] >= 0)
                {
                    const double dl = wait_value(dscratch + col[k], dep[k]);
                    d = __dsub_rn(d, __ddiv_rn(__dmul_rn(c[k], c[k]), dl));
                }
            }
        }
        publish(dscratch + r, d);
        rD[r] = __ddiv_rn(1.0, d);
    }
}

// ------------------------------------------------------------------------------------------
// K10 DICPreconditioner::precondition(wA, rA)  [EXT DICPreconditioner.C] as ONE persistent dataflow
// kernel: forward sweep into y, backward sweep into z (= wA), wArA = sum(wA*rA) fused in the tail.
//   forward : y[row] = rD[row]*rA[row] - sum_{(l,row) asc f} (rD[row]*upper[f])*y[l]
//   backward: z[row] = y[row]          - sum_{(row,u) desc f} (rD[row]*upper[f])*z[u]
// y and z must hold SENT on entry.  y is re-armed (SENT) here; z is re-armed by k_xr_update.
// Deadlock freedom: slices are in topological order, dealt round-robin to resident warps that walk
// them in increasing (forward) / decreasing (backward) order, so the smallest unfinished slice
// always has its inputs; requires all CTAs co-resident (grid <= occupancy x numSMs).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_dic_sweeps(const SliceInfo* __restrict__ slices, int nSlices,
                                                    const double* __restrict__ rD, const double* __restrict__ Uval,
                                                    const int* __restrict__ Ucol, const int* __restrict__ Lidx,
                                                    const int* __restrict__ Lcol, const double* __restrict__ rA,
                                                    unsigned long long* y, unsigned long long* z,
                                                    double* partials, DevState* st, int gated)
{
    if (gated && st->done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s0 = blockIdx.x + gridDim.x*warp;
    const int stride = gridDim.x*WPB;

    // ---- forward
    int s = s0;
    for (; s < nSlices; s += stride)
    {
        const SliceInfo si = slices[s];
        const int r = s*32 + lane;
        const double rd = rD[r];
        double acc = __dmul_rn(rd, rA[r]);
        for (int j0 = 0; j0 < si.widthL; j0 += 4)
        {
            int col[4]; double c[4]; unsigned long long dep[4];
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
                col[k] = -1;
                if (j0 + k < si.widthL)
                {
                    const int e = si.offL + (j0 + k)*32 + lane;
                    col[k] = Lcol[e];
                    if (col[k] >= 0) { c[k] = Uval[Lidx[e]]; dep[k] = peek64(y + col[k]); }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
                if (col[k] >= 0)
                {
                    const double yl = wait_value(y + col[k], dep[k]);
                    acc = __dsub_rn(acc, __dmul_rn(__dmul_rn(rd, c[k]), yl));
                }
            }
        }
        publish(y + r, acc);
    }

    // ---- backward, same slices in reverse
    double dot[1] = {0.0};
    for (s -= stride; s >= s0; s -= stride)
    {
        const SliceInfo si = slices[s];
        const int r = s*32 + lane;
        const double rd = rD[r];
        const unsigned long long own = peek64(y + r);
        double acc = 0.0;
        bool first = true;
        for (int j1 = si.widthU; j1 > 0; j1 -= 4)
        {
            int col[4]; double c[4]; unsigned long long dep[4];
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
                col[k] = -1;
                const int j = j1 - 1 - k;
                if (j >= 0)
                {
                    const int e = si.offU + j*32 + lane;
                    col[k] = Ucol[e];
                    if (col[k] >= 0) { c[k] = Uval[e]; dep[k] = peek64(z + col[k]); }
                }
            }
            if (first) { acc = wait_value(y + r, own); first = false; }
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
                if (col[k] >= 0)
                {
                    const double zu = wait_value(z + col[k], dep[k]);
                    acc = __dsub_rn(acc, __dmul_rn(__dmul_rn(rd, c[k]), zu));
                }
            }
        }
        if (first) acc = wait_value(y + r, own);
        publish(z + r, acc);
        y[r] = SENT;                 // every forward reader of y[r] has finished (they precede z[u])
        dot[0] += acc*rA[r];
    }
    if (partials) grid_reduce_finish<1>(dot, partials, st, S_SWEEP, nullptr);
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
    grid_reduce_finish<1>(v, partials, st, S_INIT, nullptr);
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
    grid_reduce_finish<2>(v, partials, st, S_NORM, nullptr);
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
    if (st->done) return;
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
    grid_reduce_finish<1>(v, partials, st, S_XR, history);
}

// L2 flush helper for the roofline timings
__global__ void __launch_bounds__(TPB) k_flush(float* __restrict__ p, long long n, float v)
{
    for (long long i = (long long)blockIdx.x*TPB + threadIdx.x; i < n; i += (long long)gridDim.x*TPB) p[i] = v;
}

} // namespace gpupcg
