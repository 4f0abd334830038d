// kernels_vec.cuh -- the component-batched ("vector") variant of the gpuPCG kernels.
//
// fvMatrix<vector>::solve [EXT fvMatrixSolve.C] (DEqn.solve(), unsTotalLagrangianStress.C:908) hands the solver boundary
// NB = 2 or 3 scalar systems that share upper()/lduAddr() and differ only in diag (addBoundaryDiag per component) and
// source.  The solves are independent, and the DIC sweeps of one system are bound by the LATENCY of the dependency
// chain (nx+ny+nz-2 levels), not by bandwidth: a sweep warp spends its rounds waiting for LDS / DADD / hand-offs.  The
// kernels here advance all NB systems in the same round: NB independent chains per row in flight, one set of index
// tables, coefficients, tile hand-offs and launches for all of them.
//
// Layout: every vector is interleaved [row][NS], NS = 4 for NB = 3 (one padding lane) and 2 for NB = 2: a row is one
// aligned 32- or 16-byte unit.  A tile is still one contiguous range per array (one bulk copy), and -- what decides
// the sweep time -- the NB values of a cross-tile dependency are published with ONE 256-bit store and polled with ONE
// 256-bit load of one L2 sector.  (Three 8-byte polls per hand-off instead of one make the sweeps 1.7x slower: the
// fetch lanes' polling traffic on the lines being handed over is what sets the hop latency,
// profiles/r1c_dic_latency_study.md section 4 and profiles/r1d_vector_solve.md.)  Arithmetic per component is EXACTLY that of the scalar kernels (same operand order, no
// FMA), so each component reproduces the reference's PCG::solve for that component; each component carries its own
// scalar recurrence (DevState[k]) and stops on its own; the launch sequence ends when all have stopped.
// Table variant only (<= 4 faces per row and side: hex / tet / prism meshes); other meshes take the scalar path.
#pragma once
#include "kernels.cuh"

namespace gpupcg
{

// row stride of the interleaved vectors
template <int NB> struct VecStride { static constexpr int v = (NB == 3) ? 4 : NB; };

// NB values of one row (p 16-byte aligned: rows are 16 or 32 bytes) out of SHARED memory, for lanes that read consecutive
// rows (32-byte stride): half the lanes fetch the upper 16 bytes first, so that the eight lanes of a quarter-warp cover
// all 32 banks in each of the two accesses (two 2-way conflicts otherwise: every access touches only half of each row).
template <int NB>
__device__ __forceinline__ void row_load_sw(const double* p, double (&v)[NB], bool hiFirst)
{
    if (NB == 3)
    {
        const double2 a = *reinterpret_cast<const double2*>(p + (hiFirst ? 2 : 0));
        const double2 b = *reinterpret_cast<const double2*>(p + (hiFirst ? 0 : 2));
        v[0] = hiFirst ? b.x : a.x; v[1] = hiFirst ? b.y : a.y; v[2] = hiFirst ? a.x : b.x;
    }
    else
    {
        const double2 a = *reinterpret_cast<const double2*>(p);
        v[0] = a.x; v[1] = a.y;
    }
}
template <int NB>
__device__ __forceinline__ void row_store(double* p, const double (&v)[NB])
{
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
    if (NB == 3) *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], 0.0);
}

template <int NB>
__device__ __forceinline__ bool all_done(const DevState* st)
{
    bool d = true;
#pragma unroll
    for (int k = 0; k < NB; k++) d = d && (st[k].done != 0);
    return d;
}

__device__ __forceinline__ bool hi_is_sent(double v)
{
    // inputs are canonicalised and arithmetic only produces the canonical NaN: the high word identifies SENT
    return __double2hiint(v) == (int)(SENT >> 32);
}

// ------------------------------------------------------------------------------------------
// grid reduction of NV sums for each of NB components (v[k*NV + j]); the last block advances every component's
// scalar recurrence (components that have stopped are skipped).  Multi-rank: peer-memory mailboxes only.
// ------------------------------------------------------------------------------------------
template <int NV, int NB, int NT>
__device__ __forceinline__ void grid_reduce_finish_v(double (&v)[NV*NB], double* partials, int stride, int slot, int nSlots,
                                                     DevState* st, int step, double* history, int histStride,
                                                     int redSlot = 0, bool lastOfStep = true)
{
    constexpr int NW = NT/32, NVB = NV*NB;
    __shared__ double sm[NVB][NW];
    __shared__ bool amLast;
    if (st[0].ordered) return;  // verification mode: k_ordered_step_v forms the sums after this kernel
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NVB; k++)
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
        for (int k = 0; k < NVB; k++)
        {
            double a = sm[k][0];
#pragma unroll
            for (int w = 1; w < NW; w++) a += sm[k][w];
            __stcg(&partials[(size_t)k*stride + slot], a);
        }
        __threadfence();
        const unsigned int t = atomicAdd(&st[0].ticket, 1u);
        amLast = (t == (unsigned int)nSlots - 1);
    }
    __syncthreads();
    if (!amLast) return;
    __threadfence();
    double f[NVB];
#pragma unroll
    for (int k = 0; k < NVB; k++)
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
        for (int k = 0; k < NVB; k++)
        {
            double a = sm[k][0];
#pragma unroll
            for (int w = 1; w < NW; w++) a += sm[k][w];
            f[k] = a;
        }
        st[0].ticket = 0;
        if (st[0].multiRank)
        {
#pragma unroll
            for (int k = 0; k < NB; k++)
            {
#pragma unroll
                for (int j = 0; j < NV; j++) st[k].red[redSlot + j] = f[k*NV + j];
                if (step == S_AMUL && redSlot == 0) st[k].red[1] = 0.0;
            }
            if (st[0].p2p && lastOfStep)
            {
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
        }
        else
        {
#pragma unroll
            for (int k = 0; k < NB; k++)
                if (!(st[k].done && step >= S_SWEEP))
                    dev_scalar_step(&st[k], step, &f[k*NV], history ? history + (size_t)k*histStride : nullptr);
        }
    }
}


// ------------------------------------------------------------------------------------------
// layout kernels: foam's Field<vector> ([cell][NB], the caller's cell order) <-> [row][NS] in tile order
// ------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(TPB) k_rows_in_v(const double* __restrict__ srcOld, double* __restrict__ dstNew,
                                                   const int* __restrict__ rowOld, long long nFlat, double padValue)
{
    constexpr int NS = VecStride<NB>::v;
    for (long long i = (long long)blockIdx.x*TPB + threadIdx.x; i < nFlat; i += (long long)gridDim.x*TPB)
    {
        const int r = (int)(i/NS), k = (int)(i - (long long)r*NS);
        const int c = rowOld[r];
        dstNew[i] = (k >= NB) ? 0.0 : ((c >= 0) ? srcOld[(size_t)c*NB + k] : padValue);
    }
}

template <int NB>
__global__ void __launch_bounds__(TPB) k_rows_out_v(const double* __restrict__ srcNew, double* __restrict__ dstOld,
                                                    const int* __restrict__ rowOld, long long nFlat)
{
    constexpr int NS = VecStride<NB>::v;
    for (long long i = (long long)blockIdx.x*TPB + threadIdx.x; i < nFlat; i += (long long)gridDim.x*TPB)
    {
        const int r = (int)(i/NS), k = (int)(i - (long long)r*NS);
        const int c = rowOld[r];
        if (c >= 0 && k < NB) dstOld[(size_t)c*NB + k] = srcNew[i];
    }
}

// one component of an interleaved field <-> a scalar field (the sequential fall-back of the vector entry points)
__global__ void __launch_bounds__(TPB) k_cmpt_get(const double* __restrict__ v, int nb, int k, double* __restrict__ s, int n)
{
    for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) s[i] = v[(size_t)i*nb + k];
}
__global__ void __launch_bounds__(TPB) k_cmpt_put(const double* __restrict__ s, int nb, int k, double* __restrict__ v, int n)
{
    for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) v[(size_t)i*nb + k] = s[i];
}

// send buffer / neighbour values / interface coefficients stay compact: [face][NB]
template <int NB>
__global__ void __launch_bounds__(TPB) k_iface_gather_v(const double* __restrict__ x, const int* __restrict__ ifFaceRow,
                                                        double* __restrict__ sendBuf, int nFlat)
{
    constexpr int NS = VecStride<NB>::v;
    for (int i = blockIdx.x*TPB + threadIdx.x; i < nFlat; i += gridDim.x*TPB)
    {
        const int e = i/NB, k = i - e*NB;
        sendBuf[i] = x[(size_t)ifFaceRow[e]*NS + k];
    }
}

// updateInterfaceMatrix for NB components: bou / xNbr are [face][NB]
template <int NB, int DOT>
__global__ void __launch_bounds__(TPB) k_iface_update_v(const int* __restrict__ ifRow, const int* __restrict__ ifRowStart,
                                                        const int* __restrict__ ifEntry, int nIfRows,
                                                        const double* __restrict__ bou, const double* __restrict__ xNbr,
                                                        int constX, const double* __restrict__ x,
                                                        double* __restrict__ Ax, double* partials, DevState* st, int gated)
{
    constexpr int NS = VecStride<NB>::v;
    if (gated && all_done<NB>(st)) return;
    double dot[NB];
    double xc[NB];
#pragma unroll
    for (int k = 0; k < NB; k++) { dot[k] = 0.0; xc[k] = constX ? st[k].xRef : 0.0; }
    for (int i = blockIdx.x*TPB + threadIdx.x; i < nIfRows; i += gridDim.x*TPB)
    {
        const int r = ifRow[i];
        double acc[NB];
#pragma unroll
        for (int k = 0; k < NB; k++) acc[k] = Ax[(size_t)r*NS + k];
        for (int q = ifRowStart[i]; q < ifRowStart[i + 1]; q++)
        {
            const int e = ifEntry[q];
#pragma unroll
            for (int k = 0; k < NB; k++)
            {
                const double xn = constX ? xc[k] : xNbr[(size_t)e*NB + k];
                acc[k] = __dsub_rn(acc[k], __dmul_rn(bou[(size_t)e*NB + k], xn));
            }
        }
#pragma unroll
        for (int k = 0; k < NB; k++)
        {
            Ax[(size_t)r*NS + k] = acc[k];
            if (DOT) dot[k] += acc[k]*x[(size_t)r*NS + k];
        }
    }
    if (DOT) grid_reduce_finish_v<1, NB, TPB>(dot, partials, MAXPART, blockIdx.x, gridDim.x, st, S_AMUL, nullptr, 0, 1);
}

// ------------------------------------------------------------------------------------------
// K8, NB components: the tile's coefficient block, operand codes and cross-tile coefficient gathers are staged once
// and applied to NB x-vectors.
// ------------------------------------------------------------------------------------------
template <int NB>
__host__ __device__ inline AmulLayout amul_layout_v(int T, int EU, int EL, int XX, int XC)
{
    constexpr int NS = VecStride<NB>::v;
    AmulLayout L;
    L.cv = 0;
    L.dg = (L.cv + 8u*(EU + XC) + 15u) & ~15u;
    L.xv = (L.dg + 8u*NS*T + 15u) & ~15u;
    L.L32 = (L.xv + 8u*NS*(T + XX) + 15u) & ~15u;
    L.U16 = L.L32 + 4u*EL;
    L.sl = (L.U16 + 2u*EU + 15u) & ~15u;
    L.bar = L.sl + 16u*(T/32 + 1);
    L.bytes = L.bar + 16;
    return L;
}

__device__ __forceinline__ void cp_async16(unsigned dst, const void* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}

template <int NB, int MODE>
__global__ void __launch_bounds__(AMUL_TPB, AMUL_V_MINB) k_amul_tile_v(TileArgs A, const double* __restrict__ diag,
                                                          const double* __restrict__ Uval, const int* __restrict__ rowOld,
                                                          const unsigned int* __restrict__ ifMask,
                                                          const double* __restrict__ x, double* __restrict__ Ax,
                                                          double* partials, int partStride, DevState* st)
{
    constexpr int NS = VecStride<NB>::v;
    if (MODE == 1 && all_done<NB>(st)) return;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    const AmulLayout L = amul_layout_v<NB>(A.amulT, A.amulEU, A.amulEL, A.amulXStride, A.amulCStride);
    const unsigned sb0 = (unsigned)__cvta_generic_to_shared(smemRaw);
    const unsigned stageBytes = (L.bytes + 127u) & ~127u;
    double xc[NB];
#pragma unroll
    for (int k = 0; k < NB; k++) xc[k] = (MODE == 2) ? st[k].xRef : 0.0;
    const int tid = threadIdx.x;
    const int T = A.amulT, EU = A.amulEU;
    const int nkX = A.amulXStride/AMUL_TPB, nkC = A.amulCStride/AMUL_TPB;

    auto issue_bulk = [&](const AmulTileMeta& M, int stage)
    {
        const unsigned sb = sb0 + stage*stageBytes, bar = sb + L.bar;
        const unsigned bytes = 8u*M.nEU + 2u*M.nEU + 4u*M.nEL + 8u*NS*M.nr + (MODE == 2 ? 0u : 8u*NS*M.nr) + 16u*(M.nr >> 5);
        mbar_expect_tx(bar, bytes);
        bulk_g2s(sb + L.cv, Uval + M.eU0, 8u*M.nEU, bar);
        bulk_g2s(sb + L.L32, A.amulL32 + M.eL0, 4u*M.nEL, bar);
        bulk_g2s(sb + L.U16, A.amulU16 + M.eU0, 2u*M.nEU, bar);
        bulk_g2s(sb + L.dg, diag + (size_t)NS*M.r0, 8u*NS*M.nr, bar);
        if (MODE != 2) bulk_g2s(sb + L.xv, x + (size_t)NS*M.r0, 8u*NS*M.nr, bar);
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
    auto gather_x = [&](unsigned dst, int row)      // one row of x: 16 or 32 aligned bytes
    {
        cp_async16(dst, x + (size_t)NS*row);
        if (NS == 4) cp_async16(dst + 16u, x + (size_t)NS*row + 2);
    };
    auto issue_gathers = [&](int tile, int stage)
    {
        const unsigned sb = sb0 + stage*stageBytes;
        const unsigned dx = sb + L.xv + 8u*NS*(T + tid), dc = sb + L.cv + 8u*(EU + tid);
        if (MODE != 2)
        {
#pragma unroll
            for (int k = 0; k < AMUL_RX; k++)
                if (iX[k] >= 0) gather_x(dx + 8u*NS*k*AMUL_TPB, iX[k]);
            for (int k = AMUL_RX; k < nkX; k++)
            {
                const int i = A.amulExtX[(size_t)tile*A.amulXStride + tid + k*AMUL_TPB];
                if (i >= 0) gather_x(dx + 8u*NS*k*AMUL_TPB, i);
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
    double dot[NB];
#pragma unroll
    for (int k = 0; k < NB; k++) dot[k] = 0.0;
    unsigned phase[2] = {0u, 0u};
    int stage = 0;
    for (int tile = t0; tile < A.nAmulTiles; tile += G, stage ^= 1)
    {
        const int nextTile = tile + G;
        const bool haveNext = nextTile < A.nAmulTiles;
        cp_async_wait_all();
        __syncthreads();
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
            for (int q = tid; q < nr; q += AMUL_TPB)
            {
                const bool real = rowOld[r0 + q] >= 0;
#pragma unroll
                for (int k = 0; k < NB; k++) xv[q*NS + k] = real ? xc[k] : 0.0;
            }
            for (int q = tid; q < A.amulXStride; q += AMUL_TPB)
            {
#pragma unroll
                for (int k = 0; k < NB; k++) xv[(T + q)*NS + k] = xc[k];
            }
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
            double fL[4], fU[4];
#pragma unroll
            for (int j = 0; j < 4; j++) cl[j] = (j < si.widthL) ? pl[32*j] : 0x80000000u;
#pragma unroll
            for (int j = 0; j < 4; j++) cuv[j] = (j < si.widthU) ? (unsigned)pu[32*j] : 0x8000u;
#pragma unroll
            for (int j = 0; j < 4; j++)
            {
                fL[j] = cv[(cl[j] >> 16) & 0x7FFFu];
                fU[j] = (j < si.widthU) ? cu[32*j] : 0.0;
            }
            double xr[NB], dr[NB], acc[NB], xL[4][NB], xU[4][NB];
            const bool hiFirst = (tid & 4) != 0;
            row_load_sw<NB>(xv + (size_t)q*NS, xr, hiFirst);
            row_load_sw<NB>(dg + (size_t)q*NS, dr, hiFirst);
#pragma unroll
            for (int j = 0; j < 4; j++)
            {
                row_load_sw<NB>(xv + (size_t)(cl[j] & 0xFFFFu)*NS, xL[j], hiFirst);
                row_load_sw<NB>(xv + (size_t)(cuv[j] & 0x7FFFu)*NS, xU[j], hiFirst);
            }
#pragma unroll
            for (int k = 0; k < NB; k++) acc[k] = __dmul_rn(dr[k], xr[k]);
            // neighbour-side faces (u == row), ascending f
#pragma unroll
            for (int j = 0; j < 4; j++)
            {
                if ((int)cl[j] >= 0)
                {
#pragma unroll
                    for (int k = 0; k < NB; k++) acc[k] = __dadd_rn(acc[k], __dmul_rn(fL[j], xL[j][k]));
                }
            }
            for (int j = 4; j < si.widthL; j++)
            {
                const unsigned c = pl[32*j];
                if ((int)c >= 0)
                {
                    const double cf = cv[(c >> 16) & 0x7FFFu];
                    const double* xs = xv + (size_t)(c & 0xFFFFu)*NS;
#pragma unroll
                    for (int k = 0; k < NB; k++) acc[k] = __dadd_rn(acc[k], __dmul_rn(cf, xs[k]));
                }
            }
            // owner-side faces (l == row), ascending f
#pragma unroll
            for (int j = 0; j < 4; j++)
            {
                if (!(cuv[j] & 0x8000u))
                {
#pragma unroll
                    for (int k = 0; k < NB; k++) acc[k] = __dadd_rn(acc[k], __dmul_rn(fU[j], xU[j][k]));
                }
            }
            for (int j = 4; j < si.widthU; j++)
            {
                const unsigned c = pu[32*j];
                if (!(c & 0x8000u))
                {
                    const double cf = cu[32*j];
                    const double* xs = xv + (size_t)(c & 0x7FFFu)*NS;
#pragma unroll
                    for (int k = 0; k < NB; k++) acc[k] = __dadd_rn(acc[k], __dmul_rn(cf, xs[k]));
                }
            }
            const int r = r0 + q;
            row_store<NB>(Ax + (size_t)r*NS, acc);
            if (MODE == 1)
            {
                const bool masked = ifMask && ((ifMask[r >> 5] >> (r & 31)) & 1u);
                if (!masked)
                {
#pragma unroll
                    for (int k = 0; k < NB; k++) dot[k] += acc[k]*xr[k];
                }
            }
        }
        Mcur = Mnext;
        Mnext = M2;
    }
    if (MODE == 1)
        grid_reduce_finish_v<1, NB, AMUL_TPB>(dot, partials, partStride, blockIdx.x, gridDim.x, st, S_AMUL, nullptr, 0, 0, ifMask == nullptr);
}

// ------------------------------------------------------------------------------------------
// K9 / K10, NB components, table variant.  Shared memory of one sweep CTA:
//   vals [(T+X+2)][NS]  tile rows (staged with the sweep's start operand: rA | diag | y), cross-tile inputs, the
//                       constant slot (1.0) and the scratch slot of idle lanes
//   rowF [T+1][4]       raw face coefficients (forward/backward) or their squares (factor): shared by the components
//   rowC [T+1] int4     value offsets (bytes, for one component; scaled by NS at use)
//   dv   [T+1][NS]      rD of the tile's rows (forward / backward)
//   cX   [T+1][NS]      forward sweep carrying the x/r update: the tile of wA = A pA
// ------------------------------------------------------------------------------------------
template <int NB>
__host__ __device__ inline SweepLayout tabv_layout(int T, int X, int I, bool withXR)
{
    constexpr int NS = VecStride<NB>::v;
    SweepLayout L;
    L.T = T; L.X = X;
    L.NV = T + X + 2;
    L.one = T + X;
    L.vals = 0;
    L.rowF = (8u*NS*L.NV + 15u) & ~15u;
    L.rowC = L.rowF + 32u*(T + 1);
    L.dv = L.rowC + 16u*(T + 1);
    L.cX = (L.dv + 8u*NS*(T + 1) + 15u) & ~15u;
    L.items = withXR ? ((L.cX + 8u*NS*(T + 1) + 15u) & ~15u) : L.cX;
    L.slot = L.items + 8u*(I + 2);
    L.bar = (L.slot + 16 + 15) & ~15u;
    L.bytes = L.bar + 16;
    L.av = L.ovf = L.Ub = L.col = L.idx = L.sl = 0;
    return L;
}

// a product pinned in program order (volatile asm keeps its place among the volatile shared-memory accesses): ptxas
// otherwise sinks the coefficient products below the readiness branch, onto the dependency chain
__device__ __forceinline__ double mul_here(double a, double b)
{
    double r;
    asm volatile("mul.rn.f64 %0, %1, %2;" : "=d"(r) : "d"(a), "d"(b));
    return r;
}

// NB values of one row's slot in shared memory (16-byte aligned): weak / volatile (the polled form) / store.
// NB = 3: two 16-byte accesses per row; lanes with (lane & 4) take the upper half first, so that the eight lanes of a
// quarter-warp working on consecutive rows (32-byte stride) cover all 32 banks in each access.
template <int NB, bool VOL>
__device__ __forceinline__ void lds_row(unsigned a, double (&v)[NB], bool hiFirst)
{
    if (NB == 3)
    {
        double p0, p1, q0, q1;
        const unsigned a0 = a + (hiFirst ? 16u : 0u), a1 = a + (hiFirst ? 0u : 16u);
        if (VOL)
        {
            asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(p0), "=d"(p1) : "r"(a0) : "memory");
            asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(q0), "=d"(q1) : "r"(a1) : "memory");
        }
        else
        {
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(p0), "=d"(p1) : "r"(a0));
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(q0), "=d"(q1) : "r"(a1));
        }
        v[0] = hiFirst ? q0 : p0; v[1] = hiFirst ? q1 : p1; v[NB - 1] = hiFirst ? p0 : q0;
    }
    else
    {
        if (VOL) asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "r"(a) : "memory");
        else     asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "r"(a));
    }
}
template <int NB>
__device__ __forceinline__ void sts_row(unsigned a, const double (&v)[NB], bool hiFirst)
{
    if (NB == 3)
    {
        const double z = 0.0;
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(a + (hiFirst ? 16u : 0u)), "d"(hiFirst ? v[NB - 1] : v[0]), "d"(hiFirst ? z : v[1]));
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(a + (hiFirst ? 0u : 16u)), "d"(hiFirst ? v[0] : v[NB - 1]), "d"(hiFirst ? v[1] : z));
    }
    else
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(a), "d"(v[0]), "d"(v[1]));
}
// one row = ONE global transaction (256 bits for NB = 3, 128 for NB = 2), at gpu scope: the hand-off between tiles
template <int NB>
__device__ __forceinline__ void publish_row(unsigned long long* p, const double (&v)[NB])
{
    if (NB == 3)
        asm volatile("st.relaxed.gpu.global.v4.b64 [%0], {%1, %2, %3, %4};"
                     :: "l"(p), "l"(__double_as_longlong(v[0])), "l"(__double_as_longlong(v[1])),
                        "l"(__double_as_longlong(v[NB - 1])), "l"(0LL) : "memory");
    else
        asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};"
                     :: "l"(p), "l"(__double_as_longlong(v[0])), "l"(__double_as_longlong(v[1])) : "memory");
}
template <int NB>
__device__ __forceinline__ void peek_row(const unsigned long long* p, unsigned long long (&w)[NB])
{
    if (NB == 3)
    {
        unsigned long long pad;
        asm volatile("ld.relaxed.gpu.global.v4.b64 {%0, %1, %2, %3}, [%4];"
                     : "=l"(w[0]), "=l"(w[1]), "=l"(w[NB - 1]), "=l"(pad) : "l"(p) : "memory");
    }
    else
        asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(w[0]), "=l"(w[1]) : "l"(p) : "memory");
}

template <int NB>
struct RowRegsV
{
    unsigned q8;        // byte offset of the row's value slot (scratch slot for idle lanes)
    int qg;             // tile-local row for the global publish, -1 = idle lane
    int4 c;             // absolute shared addresses of the four dependencies' value slots (formed one item ahead)
    double f[4];        // raw coefficients of the four dependencies (upper, or upper^2 for the factor)
    double d[NB];       // rD of the row per component (unused by the factor)
    double a[NB];       // start operand per component: rA | diag | y
};

// static data of one row, LOADS ONLY, one work item ahead of its use.  The products the reference forms first
// (MODE 0: rD*rA and rD*upper; MODE 2: rD*upper) are formed from these registers in the NEXT iteration, right after
// that iteration's dependency loads have been issued: they fill the LDS latency instead of delaying the loads.
template <int NB, int MODE, bool SW>
__device__ __forceinline__ RowRegsV<NB> load_row_v(unsigned sb, const SweepLayout& L, int2 item, int lane)
{
    constexpr int NS = VecStride<NB>::v;
    RowRegsV<NB> R;
    const bool active = lane < (item.y & 0xffff);
    const int q = active ? item.x + lane : L.T;
    R.qg = active ? q : -1;
    R.q8 = active ? 8u*NS*q : 8u*NS*(L.one + 1);
    const int4 c = lds_v4(sb + L.rowC + 16u*q);
    const int vb = (int)(sb + L.vals);
    R.c = make_int4(vb + NS*c.x, vb + NS*c.y, vb + NS*c.z, vb + NS*c.w);
    const double2 fa = lds_d2(sb + L.rowF + 32u*q);
    const double2 fb = lds_d2(sb + L.rowF + 32u*q + 16u);
    R.f[0] = fa.x; R.f[1] = fa.y; R.f[2] = fb.x; R.f[3] = fb.y;
    const bool hiFirst = SW && (lane & 4) != 0;
    lds_row<NB, false>(sb + L.vals + R.q8, R.a, hiFirst);
    if (MODE != 1) lds_row<NB, false>(sb + L.dv + 8u*NS*q, R.d, hiFirst);
    else
    {
#pragma unroll
        for (int k = 0; k < NB; k++) R.d[k] = 0.0;
    }
    return R;
}

// SW: bank-swizzled access order of the row slots (see lds_row).  Faster where the sweeps are latency-bound (0.322 ->
// 0.298 ms per precondition at 1 M cells) but 30 registers dearer (120: four CTAs per SM instead of five), which costs
// more than it gives where they are throughput-bound (1.01 -> 1.14 ms at 8 M cells): chosen by size on the host.
template <int NB, int MODE, bool SW>
__global__ void __launch_bounds__(SWEEP_TPB) k_dic_tile_v(TileArgs A, const double* __restrict__ diagOrRD,
                                                          double* rA, unsigned long long* y,
                                                          unsigned long long* out, double* __restrict__ rDout,
                                                          unsigned long long* tileTicket, double* partials, int partStride,
                                                          DevState* st, int gated,
                                                          const int4* __restrict__ tabC, const double* __restrict__ tabF,
                                                          FuseXR fuse, int histStride)
{
    constexpr int NS = VecStride<NB>::v;
    if (gated && all_done<NB>(st)) return;
    constexpr bool BWD = (MODE == 2), FACTOR = (MODE == 1);
    const bool fused = (MODE == 0) && fuse.x != nullptr;
    bool doXR = false;
    double alphaXR[NB], sumXR[NB];
    bool pend[NB];
#pragma unroll
    for (int k = 0; k < NB; k++)
    {
        pend[k] = fused && st[k].pending;
        doXR = doXR || pend[k];
        alphaXR[k] = pend[k] ? __ddiv_rn(st[k].wArA, st[k].wApA) : 0.0;
        sumXR[k] = 0.0;
    }
    extern __shared__ __align__(128) unsigned char smemRaw[];
    const int X = BWD ? A.maxExtB : A.maxExtF;
    const SweepLayout L = tabv_layout<NB>(A.tileRowsMax, X, A.maxItems, fused);
    // the shared window base goes through an opaque move: ptxas otherwise rematerialises it INSIDE the round loop
    // (S2R SR_CgaCtaId + LEA, ~25 cycles in front of the dependency loads of every round)
    unsigned sb;
    asm volatile("mov.u32 %0, %1;" : "=r"(sb) : "r"((unsigned)__cvta_generic_to_shared(smemRaw)));
    const unsigned vb = sb + L.vals;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long tBegin = (g_tileTimes && threadIdx.x == 0) ? gtime() : 0ULL;
    const int tile = take_tile(tileTicket, A.nTiles, (int*)(smemRaw + L.slot), BWD);
    const int r0 = A.tileStart[tile], nr = A.tileStart[tile + 1] - r0;
    const int ip = A.tileItemPtr[tile], nItems = A.tileItemPtr[tile + 1] - ip;
    const int xBase = BWD ? A.extBPtr[tile] : A.extFPtr[tile];
    const int nExt = (BWD ? A.extBPtr[tile + 1] : A.extFPtr[tile + 1]) - xBase;
    const int* extCol = BWD ? A.extBCol : A.extFCol;
    const int T = L.T;
    double* vals = (double*)(smemRaw + L.vals);
    const size_t g0 = (size_t)NS*r0;                   // flat offset of the tile in the interleaved vectors
    const int nFlat = NS*nr;

    // ---- staging: bulk copies only
    if (threadIdx.x == 0)
    {
        mbar_init(sb + L.bar, 1);
        const unsigned bytes = 32u*nr + 16u*nr + 8u*nFlat + (FACTOR ? 0u : 8u*nFlat) + (doXR ? 8u*nFlat : 0u);
        mbar_expect_tx(sb + L.bar, bytes);
        bulk_g2s(sb + L.rowF, tabF + 4*(size_t)r0, 32u*nr, sb + L.bar);
        bulk_g2s(sb + L.rowC, tabC + r0, 16u*nr, sb + L.bar);
        if (!FACTOR) bulk_g2s(sb + L.dv, diagOrRD + g0, 8u*nFlat, sb + L.bar);
        bulk_g2s(sb + L.vals, FACTOR ? (const void*)(diagOrRD + g0) : (BWD ? (const void*)(y + g0) : (const void*)(rA + g0)),
                 8u*nFlat, sb + L.bar);
        if (doXR) bulk_g2s(sb + L.cX, fuse.wA + g0, 8u*nFlat, sb + L.bar);
    }
    {
        int2* items = (int2*)(smemRaw + L.items);
        for (int i = threadIdx.x; i < nItems + 2; i += SWEEP_TPB)
            items[i] = (i < nItems) ? A.items[ip + (BWD ? nItems - 1 - i : i)] : make_int2(0, 0);
        for (int i = threadIdx.x; i < NS*nExt; i += SWEEP_TPB) ((unsigned long long*)vals)[NS*T + i] = SENT;
        if (threadIdx.x < 4)            // the dummy row of idle lanes (never touched by the bulk copies: nr <= T)
        {
            ((int*)(smemRaw + L.rowC))[4*T + threadIdx.x] = 8*L.one;
            ((double*)(smemRaw + L.rowF))[4*T + threadIdx.x] = 0.0;
        }
        if (threadIdx.x < NS)
        {
            vals[NS*L.one + threadIdx.x] = 1.0;
            vals[NS*(L.one + 1) + threadIdx.x] = 0.0;
            ((double*)(smemRaw + L.dv))[NS*T + threadIdx.x] = 0.0;
        }
    }
    __syncthreads();
    mbar_wait(sb + L.bar, 0);
    if (BWD && warp != 0)               // y has been read: re-arm it for the next forward sweep
        for (int i = threadIdx.x - 32; i < nFlat; i += FETCHERS) y[g0 + i] = SENT;
    if (doXR)
    {
        // x += alpha pA;  rA -= alpha wA;  sum|rA|  per component that has an update pending;  wA re-armed for all
        const double* bv = (const double*)(smemRaw + L.cX);
        for (int i = threadIdx.x; i < nFlat; i += SWEEP_TPB)
        {
            const int k = i % NS;
            const size_t g = g0 + i;
            double al = alphaXR[0]; bool pe = pend[0];
#pragma unroll
            for (int c = 1; c < NB; c++) { if (k == c) { al = alphaXR[c]; pe = pend[c]; } }
            if (k >= NB) pe = false;
            if (pe)
            {
                fuse.x[g] = __dadd_rn(fuse.x[g], __dmul_rn(al, fuse.pA[g]));
                const double rr = __dsub_rn(vals[i], __dmul_rn(al, bv[i]));
                vals[i] = rr;
                fuse.rA[g] = rr;
                const double m = fabs(rr);
#pragma unroll
                for (int c = 0; c < NB; c++) sumXR[c] += (k == c) ? m : 0.0;
            }
            fuse.wA[g] = SENT;
        }
        __syncthreads();                // the sweep warp forms its start values from the updated residual
    }

    if (g_tileTimes && threadIdx.x == 0 && MODE == 0) { g_tileTimes[3*tile] = tBegin; g_tileTimes[3*tile + 1] = gtime(); }
    if (warp == 0)
    {
        const unsigned ib = sb + L.items;
        const int syncBit = BWD ? (1 << 16) : (1 << 17);       // end of a round in sweep order
        const int one8 = (int)vb + NS*8*L.one;                 // address of the constant slot: an unused dependency
        const bool hiF = SW && (lane & 4) != 0;
        int2 d0 = lds_i2(ib), d1 = lds_i2(ib + 8u);
        RowRegsV<NB> cur = load_row_v<NB, MODE, SW>(sb, L, d0, lane);
        for (int i = 0; i < nItems; i++)
        {
            const int2 d2 = lds_i2(ib + 8u*(i + 2));
            // 1. the dependency loads of this item: the only loads on the round-to-round chain go first
            const unsigned a0 = (unsigned)cur.c.x, a1 = (unsigned)cur.c.y, a2 = (unsigned)cur.c.z, a3 = (unsigned)cur.c.w;
            double v[4][NB];
            lds_row<NB, false>(a0, v[0], hiF); lds_row<NB, false>(a1, v[1], hiF);
            lds_row<NB, false>(a2, v[2], hiF); lds_row<NB, false>(a3, v[3], hiF);
            // 2. static data of the next item (loads only) ...
            const RowRegsV<NB> nxt = load_row_v<NB, MODE, SW>(sb, L, d1, lane);
            // 3. ... and this item's coefficient products, from registers, while the dependency loads are in flight
            double F[4][NB], acc0[NB];
#pragma unroll
            for (int k = 0; k < NB; k++)
            {
                acc0[k] = (MODE == 0) ? mul_here(cur.d[k], cur.a[k]) : canon(cur.a[k]);
#pragma unroll
                for (int j = 0; j < 4; j++) F[j][k] = (MODE == 1) ? cur.f[j] : mul_here(cur.d[k], cur.f[j]);
            }
            while (true)
            {
                bool notYet = false;
#pragma unroll
                for (int k = 0; k < NB; k++)
                    notYet = notYet | hi_is_sent(v[0][k]) | hi_is_sent(v[1][k]) | hi_is_sent(v[2][k]) | hi_is_sent(v[3][k]);
                if (!notYet) break;
                lds_row<NB, true>(a0, v[0], hiF); lds_row<NB, true>(a1, v[1], hiF);
                lds_row<NB, true>(a2, v[2], hiF); lds_row<NB, true>(a3, v[3], hiF);
            }
            double acc[NB];
#pragma unroll
            for (int k = 0; k < NB; k++)
            {
                double a = acc0[k];
                if (BWD)
                {
                    const double t0 = __dmul_rn(F[0][k], v[0][k]), t1 = __dmul_rn(F[1][k], v[1][k]);
                    const double t2 = __dmul_rn(F[2][k], v[2][k]), t3 = __dmul_rn(F[3][k], v[3][k]);
                    a = __dsub_rn(__dsub_rn(__dsub_rn(__dsub_rn(a, t3), t2), t1), t0);
                }
                else if (FACTOR)
                {
                    const double t0 = (cur.c.x != one8) ? __ddiv_rn(F[0][k], v[0][k]) : 0.0;
                    const double t1 = (cur.c.y != one8) ? __ddiv_rn(F[1][k], v[1][k]) : 0.0;
                    const double t2 = (cur.c.z != one8) ? __ddiv_rn(F[2][k], v[2][k]) : 0.0;
                    const double t3 = (cur.c.w != one8) ? __ddiv_rn(F[3][k], v[3][k]) : 0.0;
                    a = __dsub_rn(__dsub_rn(__dsub_rn(__dsub_rn(a, t0), t1), t2), t3);
                }
                else
                {
                    const double t0 = __dmul_rn(F[0][k], v[0][k]), t1 = __dmul_rn(F[1][k], v[1][k]);
                    const double t2 = __dmul_rn(F[2][k], v[2][k]), t3 = __dmul_rn(F[3][k], v[3][k]);
                    a = __dsub_rn(__dsub_rn(__dsub_rn(__dsub_rn(a, t0), t1), t2), t3);
                }
                acc[k] = a;
            }
            sts_row<NB>(vb + cur.q8, acc, hiF);
            if (cur.qg >= 0) publish_row<NB>(out + g0 + (size_t)NS*cur.qg, acc);
            if (d0.y & syncBit) __syncwarp();
            cur = nxt; d0 = d1; d1 = d2;
        }
        if (g_tileTimes && lane == 0 && MODE == 0) g_tileTimes[3*tile + 2] = gtime();
    }
    else
    {
        // fetch lanes: one cross-tile entry = one row of another tile = ONE poll; delivered when all NB values are there
        constexpr int stride = FETCHERS;
        int i = (warp - 1)*32 + lane;
        const int ns = g_spinNs;
        const unsigned long long* p = (i < nExt) ? out + (size_t)NS*extCol[xBase + i] : nullptr;
        int colNext = (i + stride < nExt) ? extCol[xBase + i + stride] : 0;
        SpinWatch watch;
        bool bail = false;
        while (i < nExt)
        {
            unsigned long long w[NB];
            if (bail) { for (int k = 0; k < NB; k++) w[k] = 0ULL; }
            else peek_row<NB>(p, w);
            bool ok = true;
#pragma unroll
            for (int k = 0; k < NB; k++) ok = ok && (w[k] != SENT);
            if (ok)
            {
                watch.progress();
                const unsigned dst = vb + 8u*NS*(T + i);
                asm volatile("st.volatile.shared.v2.b64 [%0], {%1, %2};" :: "r"(dst), "l"(w[0]), "l"(w[1]) : "memory");
                if (NB == 3) sts_u64(dst + 16u, w[NB - 1]);
                i += stride;
                p = out + (size_t)NS*colNext;
                if (i + stride < nExt) colNext = extCol[xBase + i + stride];
            }
            else if (watch.expired(SPIN_TILE_FETCH_V)) bail = true;
            else if (ns > 0) __nanosleep(ns);
        }
    }
    if (fused)      // every CTA of a fused launch takes part (S_XR is a no-op for components with nothing pending)
        grid_reduce_finish_v<1, NB, SWEEP_TPB>(sumXR, partials, partStride, tile, A.nTiles, st, S_XR, fuse.history, histStride);
    if (FACTOR)
    {
        __syncthreads();
        for (int i = threadIdx.x; i < nFlat; i += SWEEP_TPB) rDout[g0 + i] = (i % NS < NB) ? __ddiv_rn(1.0, vals[i]) : 0.0;
    }
    if (BWD && partials)
    {
        __syncthreads();
        double dot[NB];
#pragma unroll
        for (int k = 0; k < NB; k++) dot[k] = 0.0;
        for (int i = threadIdx.x; i < nFlat; i += SWEEP_TPB)
        {
            const int k = i % NS;
            const double t = (k < NB) ? vals[i]*rA[g0 + i] : 0.0;
#pragma unroll
            for (int c = 0; c < NB; c++) dot[c] += (k == c) ? t : 0.0;
        }
        grid_reduce_finish_v<1, NB, SWEEP_TPB>(dot, partials, partStride, tile, A.nTiles, st, S_SWEEP, nullptr, 0);
    }
}

// fvMatrix::addBoundaryDiag / addBoundarySource [EXT fvMatrix.C] of the non-coupled patches for the three components of
// the D equation, in fvMatrix's order (a cell's boundary faces ascending = patches in index order, faces in patch
// order):  diagC[c][k] = diag[c] + sum internalCoeffs[bf][k],  bC[c][k] = source[c][k] + sum boundaryCoeffs[bf][k]
__global__ void __launch_bounds__(TPB) k_add_boundary3(int nCells, const int* __restrict__ cellBfStart, const int* __restrict__ cellBf,
                                                       const double* __restrict__ diag, const double* __restrict__ source,
                                                       const double* __restrict__ ic, const double* __restrict__ bc,
                                                       double* __restrict__ diagC, double* __restrict__ bC)
{
    for (int c = blockIdx.x*TPB + threadIdx.x; c < nCells; c += gridDim.x*TPB)
    {
        const double d0 = diag[c];
        double d[3] = {d0, d0, d0}, b[3] = {source[3*(size_t)c], source[3*(size_t)c + 1], source[3*(size_t)c + 2]};
        for (int q = cellBfStart[c]; q < cellBfStart[c + 1]; q++)
        {
            const size_t bf = (size_t)cellBf[q];
#pragma unroll
            for (int k = 0; k < 3; k++) { d[k] = __dadd_rn(d[k], ic[3*bf + k]); b[k] = __dadd_rn(b[k], bc[3*bf + k]); }
        }
#pragma unroll
        for (int k = 0; k < 3; k++) { diagC[3*(size_t)c + k] = d[k]; bC[3*(size_t)c + k] = b[k]; }
    }
}

// the same for the NB valid components of a 2-D (empty) or 3-D case: component k of the solve is direction cm[k] of D.
// Also packs x[c][k] = D[c][cm[k]] (psi.component(cmpt), fvMatrixSolve.C)
template <int NB>
__global__ void __launch_bounds__(TPB) k_add_boundary_v(int nCells, const int* __restrict__ cellBfStart, const int* __restrict__ cellBf,
                                                        const double* __restrict__ diag, const double* __restrict__ source,
                                                        const double* __restrict__ ic, const double* __restrict__ bc,
                                                        const double* __restrict__ D, int cm0, int cm1, int cm2,
                                                        double* __restrict__ diagC, double* __restrict__ bC, double* __restrict__ xC)
{
    const int cm[3] = {cm0, cm1, cm2};
    for (int c = blockIdx.x*TPB + threadIdx.x; c < nCells; c += gridDim.x*TPB)
    {
        const double d0 = diag[c];
        double d[NB], b[NB];
#pragma unroll
        for (int k = 0; k < NB; k++) { d[k] = d0; b[k] = source[3*(size_t)c + cm[k]]; }
        for (int q = cellBfStart[c]; q < cellBfStart[c + 1]; q++)
        {
            const size_t bf = (size_t)cellBf[q];
#pragma unroll
            for (int k = 0; k < NB; k++) { d[k] = __dadd_rn(d[k], ic[3*bf + cm[k]]); b[k] = __dadd_rn(b[k], bc[3*bf + cm[k]]); }
        }
#pragma unroll
        for (int k = 0; k < NB; k++)
        {
            diagC[(size_t)NB*c + k] = d[k]; bC[(size_t)NB*c + k] = b[k];
            if (xC) xC[(size_t)NB*c + k] = D[3*(size_t)c + cm[k]];
        }
    }
}

// psi.replace(cmpt, psiCmpt) for the solved components
template <int NB>
__global__ void __launch_bounds__(TPB) k_unpack_cmpts(int nCells, const double* __restrict__ xC, int cm0, int cm1, int cm2, double* __restrict__ D)
{
    const int cm[3] = {cm0, cm1, cm2};
    for (int c = blockIdx.x*TPB + threadIdx.x; c < nCells; c += gridDim.x*TPB)
    {
#pragma unroll
        for (int k = 0; k < NB; k++) D[3*(size_t)c + cm[k]] = xC[(size_t)NB*c + k];
    }
}

// raw face coefficients per row (shared by the components)
__global__ void __launch_bounds__(TPB) k_tab_raw(int nRows, const int4* __restrict__ idx, const double* __restrict__ Uval,
                                                 double* __restrict__ F)
{
    for (int r = blockIdx.x*TPB + threadIdx.x; r < nRows; r += gridDim.x*TPB)
    {
        const int4 i4 = idx[r];
        const int ii[4] = {i4.x, i4.y, i4.z, i4.w};
        double f[4];
#pragma unroll
        for (int j = 0; j < 4; j++) f[j] = (ii[j] >= 0) ? canon(Uval[ii[j]]) : 0.0;
        double2* o = (double2*)(F + 4*(size_t)r);
        o[0] = make_double2(f[0], f[1]);
        o[1] = make_double2(f[2], f[3]);
    }
}

// ------------------------------------------------------------------------------------------
// K11, NB components: flat loops over the interleaved vectors (fully coalesced), component = flat index mod NS
// (the padding lane of NB = 3 is carried along untouched)
// ------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(TPB) k_init_residual_v(const double* __restrict__ b, const double* __restrict__ wA,
                                                         const double* __restrict__ x, double* __restrict__ rA,
                                                         long long nFlat, double* partials, DevState* st)
{
    constexpr int NS = VecStride<NB>::v;
    double v[NB];
#pragma unroll
    for (int k = 0; k < NB; k++) v[k] = 0.0;
    for (long long i = (long long)blockIdx.x*TPB + threadIdx.x; i < nFlat; i += (long long)gridDim.x*TPB)
    {
        const int k = (int)(i % NS);
        if (k >= NB) { rA[i] = 0.0; continue; }
        rA[i] = __dsub_rn(b[i], wA[i]);
        const double t = x[i];
#pragma unroll
        for (int c = 0; c < NB; c++) v[c] += (k == c) ? t : 0.0;
    }
    grid_reduce_finish_v<1, NB, TPB>(v, partials, MAXPART, blockIdx.x, gridDim.x, st, S_INIT, nullptr, 0);
}

template <int NB>
__global__ void __launch_bounds__(TPB) k_norm_v(const double* __restrict__ wA, const double* __restrict__ t,
                                                const double* __restrict__ b, const double* __restrict__ rA,
                                                long long nFlat, double* partials, DevState* st)
{
    constexpr int NS = VecStride<NB>::v;
    double v[2*NB];
#pragma unroll
    for (int k = 0; k < 2*NB; k++) v[k] = 0.0;
    for (long long i = (long long)blockIdx.x*TPB + threadIdx.x; i < nFlat; i += (long long)gridDim.x*TPB)
    {
        const int k = (int)(i % NS);
        if (k >= NB) continue;
        const double tr = t[i];
        const double n0 = __dadd_rn(fabs(__dsub_rn(wA[i], tr)), fabs(__dsub_rn(b[i], tr)));
        const double n1 = fabs(rA[i]);
#pragma unroll
        for (int c = 0; c < NB; c++) { v[2*c] += (k == c) ? n0 : 0.0; v[2*c + 1] += (k == c) ? n1 : 0.0; }
    }
    grid_reduce_finish_v<2, NB, TPB>(v, partials, MAXPART, blockIdx.x, gridDim.x, st, S_NORM, nullptr, 0);
}

// one row (NS doubles, 16- or 32-byte aligned) per access: 256-bit loads/stores for NB = 3
template <int NB>
__device__ __forceinline__ void grow_load(const double* p, double (&v)[VecStride<NB>::v])
{
    if (NB == 3)
        asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[VecStride<NB>::v - 1]) : "l"(p));
    else
        asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "l"(p));
}
template <int NB>
__device__ __forceinline__ void grow_store(double* p, const double (&v)[VecStride<NB>::v])
{
    if (NB == 3)
        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" :: "l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[VecStride<NB>::v - 1]) : "memory");
    else
        asm volatile("st.global.v2.f64 [%0], {%1, %2};" :: "l"(p), "d"(v[0]), "d"(v[1]) : "memory");
}

template <int NB>
__global__ void __launch_bounds__(TPB) k_p_update_v(const double* __restrict__ wA, double* __restrict__ pA,
                                                    long long nFlat, const DevState* __restrict__ st)
{
    constexpr int NS = VecStride<NB>::v;
    if (all_done<NB>(st)) return;
    bool first[NB], skip[NB];
    double beta[NB];
#pragma unroll
    for (int k = 0; k < NB; k++)
    {
        skip[k] = st[k].done != 0;
        first[k] = (st[k].nIter == 0);
        beta[k] = first[k] ? 0.0 : __ddiv_rn(st[k].wArA, st[k].wArAold);
    }
    const long long nRows = nFlat/NS;
    for (long long r = (long long)blockIdx.x*TPB + threadIdx.x; r < nRows; r += (long long)gridDim.x*TPB)
    {
        double w[NS], p[NS];
        grow_load<NB>(wA + r*NS, w);
        grow_load<NB>(pA + r*NS, p);
#pragma unroll
        for (int k = 0; k < NB; k++)
            if (!skip[k]) p[k] = first[k] ? w[k] : __dadd_rn(w[k], __dmul_rn(beta[k], p[k]));
        if (NS > NB) p[NS - 1] = 0.0;
        grow_store<NB>(pA + r*NS, p);
    }
}

template <int NB>
__global__ void __launch_bounds__(TPB) k_xr_update_v(const double* __restrict__ pA, unsigned long long* wAbits,
                                                     double* __restrict__ x, double* __restrict__ rA,
                                                     long long nFlat, double* partials, DevState* st, double* history,
                                                     int histStride)
{
    constexpr int NS = VecStride<NB>::v;
    bool pend[NB], any = false;
    double alpha[NB], v[NB];
#pragma unroll
    for (int k = 0; k < NB; k++)
    {
        pend[k] = st[k].pending != 0;
        any = any || pend[k];
        alpha[k] = pend[k] ? __ddiv_rn(st[k].wArA, st[k].wApA) : 0.0;
        v[k] = 0.0;
    }
    if (!any) return;
    const long long nRows = nFlat/NS;
    double sent[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) sent[k] = __longlong_as_double((long long)SENT);
    for (long long r = (long long)blockIdx.x*TPB + threadIdx.x; r < nRows; r += (long long)gridDim.x*TPB)
    {
        double w[NS], p[NS], xx[NS], rr[NS];
        grow_load<NB>((const double*)wAbits + r*NS, w);
        grow_load<NB>(pA + r*NS, p);
        grow_load<NB>(x + r*NS, xx);
        grow_load<NB>(rA + r*NS, rr);
#pragma unroll
        for (int k = 0; k < NB; k++)
        {
            if (pend[k])
            {
                xx[k] = __dadd_rn(xx[k], __dmul_rn(alpha[k], p[k]));
                rr[k] = __dsub_rn(rr[k], __dmul_rn(alpha[k], w[k]));
                v[k] += fabs(rr[k]);
            }
        }
        grow_store<NB>(x + r*NS, xx);
        grow_store<NB>(rA + r*NS, rr);
        grow_store<NB>((double*)wAbits + r*NS, sent);
    }
    grid_reduce_finish_v<1, NB, TPB>(v, partials, MAXPART, blockIdx.x, gridDim.x, st, S_XR, history, histStride);
}

} // namespace gpupcg
