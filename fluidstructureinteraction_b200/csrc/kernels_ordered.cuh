// kernels_ordered.cuh -- VERIFICATION MODE (GPUPCG_ORDERED_SUMS=1): the five global sums of PCG::solve formed exactly as
// the reference forms them -- [EXT FieldFunctions.C] sumProd / sumMag / sum: one sequential loop over the cells in the
// caller's cell order, then (multi-rank) the rank-ordered sum of the local results -- instead of the fused tree reductions
// of the production kernels.  Everything else (Amul, sweeps, updates, halo swaps, mailboxes, scalar recurrence) is the
// production path untouched, so in this mode the residual history is BIT-IDENTICAL to the oracle's at any rank count:
// what separates the production path from the reference is then the association of these sums and nothing else
// (tools/multigpu_parity.py states that sensitivity separately).  One CTA; serial in N by design -- a checker, not a product
// path: nothing selects it unless the environment variable is set when the handle is created.
#pragma once
#include "kernels.cuh"
#include "kernels_vec.cuh"

namespace gpupcg
{

constexpr int ORD_TPB = 256;

// per-cell term of sum `j` of `step`, row r, component k of NS-strided vectors
//   S_INIT : sum(x)                                            gAverage(x)        [EXT lduMatrixSolver.C normFactor]
//   S_NORM : sum(mag(wA - pA) + mag(b - pA)), sum(mag(rA))     normFactor, gSumMag(rA)
//   S_SWEEP: sum(wA*rA)                                        gSumProd(wA, rA)   [EXT PCG.C]
//   S_AMUL : sum(wA*pA)                                        gSumProd(wA, pA)
//   S_XR   : sum(mag(rA))                                      gSumMag(rA)
struct OrderedVecs { const double *x, *b, *rA, *wA, *pA; };

__device__ __forceinline__ double ordered_term(int step, int j, const OrderedVecs& V, size_t i)
{
    switch (step)
    {
        case S_INIT:  return V.x[i];
        case S_NORM:
            if (j == 0) { const double t = V.pA[i]; return __dadd_rn(fabs(__dsub_rn(V.wA[i], t)), fabs(__dsub_rn(V.b[i], t))); }
            return fabs(V.rA[i]);
        case S_SWEEP: return __dmul_rn(V.wA[i], V.rA[i]);
        case S_AMUL:  return __dmul_rn(V.wA[i], V.pA[i]);
        default:      return fabs(V.rA[i]);
    }
}

// NB components, vectors laid out [row][NS]; pos[cell] = row.  <<<1, ORD_TPB>>>
template <int NB, int NS>
__global__ void __launch_bounds__(ORD_TPB) k_ordered_step(DevState* st, int step, OrderedVecs V, const int* __restrict__ pos,
                                                          int nCells, double* history, int histStride)
{
    __shared__ double term[2][NB][ORD_TPB];
    const int nv = (step == S_NORM) ? 2 : 1;
    double acc[2][NB];
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
        for (int k = 0; k < NB; k++) acc[j][k] = 0.0;
    for (int c0 = 0; c0 < nCells; c0 += ORD_TPB)
    {
        const int c = c0 + threadIdx.x;
        if (c < nCells)
        {
            const size_t r = (size_t)pos[c];
#pragma unroll
            for (int k = 0; k < NB; k++)
                for (int j = 0; j < nv; j++) term[j][k][threadIdx.x] = ordered_term(step, j, V, r*NS + k);
        }
        __syncthreads();
        if (threadIdx.x == 0)
        {
            const int m = min(ORD_TPB, nCells - c0);
            for (int i = 0; i < m; i++)
#pragma unroll
                for (int k = 0; k < NB; k++)
                    for (int j = 0; j < nv; j++) acc[j][k] = __dadd_rn(acc[j][k], term[j][k][i]);
        }
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    // the tail of grid_reduce_finish(_v): park / all-reduce / advance the scalar recurrence of every component
    if (st[0].multiRank)
    {
#pragma unroll
        for (int k = 0; k < NB; k++)
        {
            st[k].red[0] = acc[0][k];
            st[k].red[1] = (nv == 2) ? acc[1][k] : 0.0;
        }
        if (!st[0].p2p) return;         // ncclAllReduce + k_scalar_step follow on the stream (scalar path only)
        const int cnt = (step == S_AMUL || step == S_NORM) ? 2 : 1;
        double loc[2*NB], g[2*NB];
#pragma unroll
        for (int k = 0; k < NB; k++)
            for (int j = 0; j < cnt; j++) loc[k*cnt + j] = st[k].red[j];
        p2p_allreduce(st[0].p2p, loc, cnt*NB, g);
#pragma unroll
        for (int k = 0; k < NB; k++)
        {
            double gv[2] = {g[k*cnt], (cnt == 2) ? g[k*cnt + 1] : 0.0};
            if (step == S_AMUL) gv[0] = gv[0] + gv[1];
            if (!(st[k].done && step >= S_SWEEP))
                dev_scalar_step(&st[k], step, gv, history ? history + (size_t)k*histStride : nullptr);
        }
    }
    else
    {
#pragma unroll
        for (int k = 0; k < NB; k++)
        {
            double gv[2] = {acc[0][k], acc[1][k]};
            if (!(st[k].done && step >= S_SWEEP))
                dev_scalar_step(&st[k], step, gv, history ? history + (size_t)k*histStride : nullptr);
        }
    }
}

} // namespace gpupcg
