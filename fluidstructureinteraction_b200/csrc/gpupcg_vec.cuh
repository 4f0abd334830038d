// gpupcg_vec.cuh -- host orchestration of the component-batched solve (included at the end of gpupcg.cu).
// C ABI: gpupcg_set_matrix_vector[_device], gpupcg_solve_vector[_device], gpupcg_time_kernel_vector, gpupcg_vector_mode.
// Replaces the component loop of fvMatrix<vector>::solve [EXT fvMatrixSolve.C] as DEqn.solve() reaches it
// (unsTotalLagrangianStress.C:908): the NB systems share upper()/lduAddr(), differ in diag and source, and are
// independent, so one launch sequence advances all of them (kernels_vec.cuh).  Meshes the table variant does not
// cover (more than four faces per row and side), and multi-rank runs without peer-memory mailboxes, take NB
// sequential scalar solves on the same GPU kernels instead (never a CPU path).

namespace
{

template <int NB>
int vec_prepare(gpupcg_handle h)
{
    if (h->vNB == NB) return GPUPCG_OK;
    if (h->vNB != 0)
    {
        h->err = "gpupcg vector solve: the handle is already set up for a different number of components";
        return GPUPCG_ERR_STATE;
    }
    constexpr int NS = VecStride<NB>::v;
    const HostPlan& P = h->plan;
    const long long nR = P.nRows, nC = std::max(1, P.nCells), nI = std::max(1, P.nIfaceFaces);
    CK(dalloc(h->vStageX, nC*NB)); CK(dalloc(h->vStageB, nC*NB)); CK(dalloc(h->vStageD, nC*NB));
    CK(dalloc(h->vTmpX, nC)); CK(dalloc(h->vTmpB, nC));
    CK(dalloc(h->vBou, nI*NB));
    h->smemFwdV = tabv_layout<NB>(P.tileRowsMax, P.maxTileExtF, P.maxItems, true).bytes;
    h->smemFacV = tabv_layout<NB>(P.tileRowsMax, P.maxTileExtF, P.maxItems, false).bytes;
    h->smemBwdV = tabv_layout<NB>(P.tileRowsMax, P.maxTileExtB, P.maxItems, false).bytes;
    h->smemAmulV = 2*(size_t)((amul_layout_v<NB>(h->amulT, h->amulEU, h->amulEL, h->amulXStride, h->amulCStride).bytes + 127u) & ~127u);
    h->vBatched = (h->pencil || (h->useTab && std::max(h->smemFwdV, h->smemBwdV) <= h->smemCap)) && P.nTiles > 0
               && env_int("GPUPCG_VECTOR_BATCH", 1) && h->smemAmulV <= h->smemCap;
    if (h->vBatched)
    {
        if (h->pencil) { CK(pencil_attrs<NB>(P.pLong, h->smemCap)); }
        else
        {
        CK(cudaFuncSetAttribute((k_dic_tile_v<NB, 0, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemFwdV));
        CK(cudaFuncSetAttribute((k_dic_tile_v<NB, 1, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemFacV));
        CK(cudaFuncSetAttribute((k_dic_tile_v<NB, 2, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemBwdV));
        CK(cudaFuncSetAttribute((k_dic_tile_v<NB, 0, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemFwdV));
        CK(cudaFuncSetAttribute((k_dic_tile_v<NB, 2, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemBwdV));
        }
        // latency-bound sweeps (small sub-domains) take the bank-swizzled variant, throughput-bound ones the leaner one
        const int swEnv = env_int("GPUPCG_SWEEP_SWIZZLE", -1);
        h->vSwizzle = (swEnv >= 0) ? (swEnv != 0) : (P.nRows <= env_int("GPUPCG_FUSE_XR_MAXROWS", 3000000));
        CK(cudaFuncSetAttribute((k_amul_tile_v<NB, 0>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemAmulV));
        CK(cudaFuncSetAttribute((k_amul_tile_v<NB, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemAmulV));
        CK(cudaFuncSetAttribute((k_amul_tile_v<NB, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemAmulV));
        int occ = 1;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (k_amul_tile_v<NB, 1>), AMUL_TPB, h->smemAmulV));
        occ = std::max(1, std::min(occ, env_int("GPUPCG_AMUL_BPS", 16)));
        h->amulGridV = std::max(1, std::min(h->nAmulTiles, std::min(h->numSMs*occ, MAXPART)));
        CK(dalloc(h->vDiag, nR*NS)); CK(dalloc(h->vRD, nR*NS)); CK(dalloc(h->vX, nR*NS)); CK(dalloc(h->vB, nR*NS));
        CK(dalloc(h->vRA, nR*NS));   CK(dalloc(h->vWA, nR*NS)); CK(dalloc(h->vPA, nR*NS)); CK(dalloc(h->vY, nR*NS));
        CK(dalloc(h->vXNbr, nI*NB)); CK(dalloc(h->vSend, nI*NB));
        CK(cudaMemsetAsync(h->vXNbr, 0, sizeof(double)*nI*NB, h->stream));
        CK(dalloc(h->vPartials, 2LL*NB*h->partStride));
        CK(dalloc(h->vHistory, (long long)NB*h->historyCap));
        CK(dalloc(h->vState, NB));
        CK(cudaMemsetAsync(h->vState, 0, sizeof(DevState)*NB, h->stream));
        CK(cudaMallocHost((void**)&h->hvRing, 3*NB*sizeof(DevState)));
        CK(dalloc(h->dTabSq, 4LL*std::max<long long>(1, nR)));
        CK(dalloc(h->dTabUF, 4LL*std::max<long long>(1, nR)));
        CK(dalloc(h->dTabUB, 4LL*std::max<long long>(1, nR)));
    }
    h->vNB = NB;
    return GPUPCG_OK;
}

// Multi-rank runs need the peer-memory mailboxes (the batched kernels have no ncclAllReduce path), and every rank must
// take the same path (the batched and the sequential solves issue different collectives): the ranks agree once, with
// one all-reduce, on "batched everywhere" -- a rank whose sub-domain has polyhedral rows, or no rows at all, sends
// everybody to the sequential component solves.
inline bool vec_batched_now(gpupcg_handle h)
{
    if (!h->comm.active()) return h->vBatched;
    if (h->vAgreed < 0)
    {
        const double mine = (h->vBatched && h->dP2P != nullptr) ? 1.0 : 0.0;
        double sum = 0.0;
        double* d = nullptr;
        bool ok = cudaMalloc((void**)&d, sizeof(double)) == cudaSuccess
               && cudaMemcpyAsync(d, &mine, sizeof(double), cudaMemcpyHostToDevice, h->stream) == cudaSuccess
               && h->comm.allreduce_sum(d, 1, h->stream, h->err) == 0
               && cudaMemcpyAsync(&sum, d, sizeof(double), cudaMemcpyDeviceToHost, h->stream) == cudaSuccess
               && cudaStreamSynchronize(h->stream) == cudaSuccess;
        if (d) cudaFree(d);
        h->vAgreed = (ok && sum == (double)h->comm.nRanks()) ? 1 : 0;
    }
    return h->vAgreed == 1;
}

// Concurrent component solves.  The batched kernels advance the NB systems in lockstep on one stream: the latency-bound
// sweeps (SMs mostly idle) and the bandwidth-bound Amul / update kernels alternate.  The systems are independent, so on
// a single rank they can also run as NB scalar solves in flight at once -- one stream, one host thread and one set of
// vectors per component, the plan tables and upper() shared -- and the GPU overlaps one system's sweeps with another's
// streaming kernels (measured: 1 M cells 55.5 -> 47.6 ms per D-equation solve, 8 M 426 -> 388 ms, 250 k 20.1 -> 17.5 ms;
// profiles/r2c_concurrent_components.md).  Every component is then exactly the scalar gpupcg_solve of that system.
// GPUPCG_VECTOR_CONCURRENT=0/1 overrides the default (on for structured-block plans of 100 k ... 10 M cells).
inline bool vec_concurrent_now(gpupcg_handle h)
{
    if (h->comm.active() || h->nPatches > 0 || h->plan.nIfaceFaces > 0 || h->plan.nTiles == 0 || h->isSys) return false;
    const int env = env_int("GPUPCG_VECTOR_CONCURRENT", -1);
    if (env >= 0) return env != 0 && (h->pencil || h->useTab);
    // above ~10 M cells the kernels are bandwidth-bound and the batched ones win (they read upper() and the tables once for
    // three systems): 16 M cells 1032 (batched) vs 1107 ms, 64 M 6.8 vs 7.9 s
    return h->pencil && h->plan.nCells >= env_int("GPUPCG_VECTOR_CONCURRENT_MINCELLS", 100000)
                     && h->plan.nCells <= env_int("GPUPCG_VECTOR_CONCURRENT_MAXCELLS", 10000000);
}

template <int NB>
int vec_ensure_systems(gpupcg_handle h)
{
    if (!h->evFork) CK(cudaEventCreateWithFlags(&h->evFork, cudaEventDisableTiming));
    for (int k = 0; k < NB; k++)
        if (!h->sys[k]) { int rc = make_system(h, h->sys[k], NB); if (rc) return rc; }
    return GPUPCG_OK;
}

template <int NB>
int vec_set_matrix_device(gpupcg_handle h, const double* d_diagC, const double* d_upper, const double* d_bouC)
{
    int rc = vec_prepare<NB>(h); if (rc) return rc;
    const HostPlan& P = h->plan;
    cudaStream_t s = h->stream;
    if (d_upper)
    {
        LAUNCH(k_coeffs_in, grid_for(P.entriesU, h->streamGrid), s, d_upper, h->dUval, h->dUfaceOld, P.entriesU);
        h->uvalVersion++;
    }
    else if (!h->haveMatrix && !h->vHaveMatrix && P.nFaces > 0)
    {
        h->err = "gpupcg_set_matrix_vector: upper == NULL but no previous coefficients";
        return GPUPCG_ERR_STATE;
    }
    if (d_diagC != h->vStageD)
        CK(cudaMemcpyAsync(h->vStageD, d_diagC, sizeof(double)*P.nCells*NB, cudaMemcpyDeviceToDevice, s));
    if (d_bouC && P.nIfaceFaces > 0 && d_bouC != h->vBou)
        CK(cudaMemcpyAsync(h->vBou, d_bouC, sizeof(double)*P.nIfaceFaces*NB, cudaMemcpyDeviceToDevice, s));
    if (h->vBatched)
    {
        const long long nFlat = (long long)P.nRows*VecStride<NB>::v;
        LAUNCH((k_rows_in_v<NB>), grid_for(nFlat, h->streamGrid), s, h->vStageD, h->vDiag, h->dRowOld, nFlat, 1.0);
        if (h->pencil) { int rcp = pencil_coeffs(h); if (rcp) return rcp; }
        else if (h->vTabVersion != h->uvalVersion)
        {
            h->vTabVersion = h->uvalVersion;
            const int gT = grid_for(P.nRows, h->streamGrid);
            LAUNCH((k_tab_coeffs<0>), gT, s, P.nRows, h->dTabIF, h->dUval, nullptr, h->dTabSq, h->dState, 0);
            LAUNCH(k_tab_raw, gT, s, P.nRows, h->dTabIF, h->dUval, h->dTabUF);
            LAUNCH(k_tab_raw, gT, s, P.nRows, h->dTabIB, h->dUval, h->dTabUB);
        }
    }
    if (vec_concurrent_now(h))
    {
        // diag of every component into its system, on this stream (the systems' streams wait for evFork before a solve)
        rc = vec_ensure_systems<NB>(h); if (rc) return rc;
        if (h->pencil) { int rcp = pencil_coeffs(h); if (rcp) return rcp; }
        for (int k = 0; k < NB; k++)
        {
            gpupcg_handle c = h->sys[k];
            LAUNCH(k_cmpt_get, grid_for(P.nCells, h->streamGrid), s, h->vStageD, NB, k, c->dStageA, P.nCells);
            LAUNCH(k_rows_in, grid_for(P.nRows, h->streamGrid), s, c->dStageA, c->dDiag, h->dRowOld, P.nRows, 1.0);
            c->haveMatrix = true;
            c->uvalVersion = h->uvalVersion; c->penCoefVersion = h->penCoefVersion;     // shares dUval and the pencil tables
        }
    }
    CK(cudaGetLastError());
    h->vHaveMatrix = true;
    return GPUPCG_OK;
}

// verification mode (GPUPCG_ORDERED_SUMS): the kernel launched just before skipped its tree reduction; form the sums of
// every component in the caller's cell order and advance the scalar recurrences (kernels_ordered.cuh)
template <int NB>
int vec_ordered_step(gpupcg_handle h, int step)
{
    if (!h->orderedSums) return GPUPCG_OK;
    const OrderedVecs V{h->vX, h->vB, h->vRA, h->vWA, h->vPA};
    h->launches++;
    k_ordered_step<NB, VecStride<NB>::v><<<1, ORD_TPB, 0, h->stream>>>(h->vState, step, V, h->dPos, h->plan.nCells, h->vHistory, h->historyCap);
    CK(cudaGetLastError());
    return GPUPCG_OK;
}

template <int NB>
int vec_halo(gpupcg_handle h, const double* x, bool* overlapped)
{
    const HostPlan& P = h->plan;
    *overlapped = false;
    if (!h->comm.active() || P.nIfaceFaces == 0) return GPUPCG_OK;
    cudaStream_t cs = h->stream;
    if (h->commStream)
    {
        cs = h->commStream;
        CK(cudaEventRecord(h->evX, h->stream));
        CK(cudaStreamWaitEvent(cs, h->evX, 0));
        *overlapped = true;
    }
    LAUNCH((k_iface_gather_v<NB>), grid_for((long long)P.nIfaceFaces*NB, h->streamGrid), cs, x, h->dIfFaceRow, h->vSend, P.nIfaceFaces*NB);
    if (h->comm.exchange(h->vSend, h->vXNbr, P.patchStart, P.patchNbrRank, cs, h->err, NB) != 0) return GPUPCG_ERR_NCCL;
    if (*overlapped) { CK(cudaEventRecord(h->evHalo, cs)); }
    return GPUPCG_OK;
}

template <int NB>
int vec_amul(gpupcg_handle h, int mode, const double* x, double* Ax, bool xNbrReady)
{
    const HostPlan& P = h->plan;
    cudaStream_t s = h->stream;
    const bool ifaces = P.nIfaceFaces > 0;
    bool haloInFlight = false;
    if (ifaces && mode != 2 && !xNbrReady) { int rc = vec_halo<NB>(h, x, &haloInFlight); if (rc) return rc; }
    const unsigned int* mask = (ifaces && mode == 1) ? h->dIfMask : nullptr;
    if (h->nAmulTiles > 0)
    {
        h->launches++;
        if (mode == 0)
            k_amul_tile_v<NB, 0><<<h->amulGridV, AMUL_TPB, h->smemAmulV, s>>>(h->tiles, h->vDiag, h->dUval, h->dRowOld, mask, x, Ax, h->vPartials, h->partStride, h->vState);
        else if (mode == 1)
            k_amul_tile_v<NB, 1><<<h->amulGridV, AMUL_TPB, h->smemAmulV, s>>>(h->tiles, h->vDiag, h->dUval, h->dRowOld, mask, x, Ax, h->vPartials, h->partStride, h->vState);
        else
            k_amul_tile_v<NB, 2><<<h->amulGridV, AMUL_TPB, h->smemAmulV, s>>>(h->tiles, h->vDiag, h->dUval, h->dRowOld, mask, x, Ax, h->vPartials, h->partStride, h->vState);
    }
    if (ifaces)
    {
        if (haloInFlight) { CK(cudaStreamWaitEvent(s, h->evHalo, 0)); }
        const int gi = grid_for(h->nIfRows, h->streamGrid);
        if (mode == 1)
            LAUNCH((k_iface_update_v<NB, 1>), gi, s, h->dIfRow, h->dIfRowStart, h->dIfEntry, h->nIfRows, h->vBou, h->vXNbr, 0, x, Ax, h->vPartials, h->vState, 1);
        else
            LAUNCH((k_iface_update_v<NB, 0>), gi, s, h->dIfRow, h->dIfRowStart, h->dIfEntry, h->nIfRows, h->vBou, h->vXNbr, mode == 2 ? 1 : 0, x, Ax, h->vPartials, h->vState, 0);
    }
    CK(cudaGetLastError());
    return GPUPCG_OK;
}

template <int NB>
void vec_fill_sent(gpupcg_handle h, double* p)
{
    const long long n = (long long)h->plan.nRows*VecStride<NB>::v;
    LAUNCH(k_fill64, grid_for(n, h->streamGrid), h->stream, (unsigned long long*)p, SENT, n);
}

template <int NB>
int vec_factor(gpupcg_handle h, int gated)
{
    if (h->plan.nTiles == 0) return GPUPCG_OK;
    vec_fill_sent<NB>(h, h->vY);
    const FuseXR none{nullptr, nullptr, nullptr, nullptr, nullptr};
    if (h->pencil)
    {
        int rcp = pencil_coeffs(h); if (rcp) return rcp;
        rcp = launch_pencil<NB, 1>(h, h->vDiag, nullptr, nullptr, (unsigned long long*)h->vY, h->vRD, nullptr, h->vState, gated);
        if (rcp) return rcp;
    }
    else
    {
    h->launches++;
    k_dic_tile_v<NB, 1, false><<<h->sweepGrid, SWEEP_TPB, h->smemFacV, h->stream>>>(h->tiles, h->vDiag, nullptr, nullptr,
        (unsigned long long*)h->vY, h->vRD, h->dTileTicket, nullptr, 0, h->vState, gated, h->dTabCF, h->dTabSq, none, 0);
    }
    vec_fill_sent<NB>(h, h->vY);
    CK(cudaGetLastError());
    return GPUPCG_OK;
}

// y (h->vY) and z must hold SENT on entry; y is re-armed by the backward kernel
template <int NB>
int vec_sweeps(gpupcg_handle h, double* rA, double* z, bool reduce, int gated, bool fuseXR)
{
    if (h->plan.nTiles == 0) return GPUPCG_OK;
    unsigned long long* yb = (unsigned long long*)h->vY;
    unsigned long long* zb = (unsigned long long*)z;
    const FuseXR none{nullptr, nullptr, nullptr, nullptr, nullptr};
    if (h->pencil)
    {
        int rcp = pencil_coeffs(h); if (rcp) return rcp;
        if (fuseXR)
        {
            // the pending x/r update of the previous iteration rides in the forward sweep (its own warp per CTA)
            const FuseXR fxp{h->vX, h->vPA, zb, h->vRA, h->vHistory};
            rcp = launch_pencil<NB, 0>(h, h->vRD, rA, nullptr, yb, nullptr, h->vPartials, h->vState, gated, &fxp, h->historyCap);
            if (rcp) return rcp;
            rcp = vec_ordered_step<NB>(h, S_XR); if (rcp) return rcp;
        }
        else
        {
        rcp = launch_pencil<NB, 0>(h, h->vRD, rA, nullptr, yb, nullptr, nullptr, h->vState, gated); if (rcp) return rcp;
        }
        rcp = launch_pencil<NB, 2>(h, h->vRD, rA, yb, zb, nullptr, reduce ? h->vPartials : nullptr, h->vState, gated); if (rcp) return rcp;
        CK(cudaGetLastError());
        if (reduce) return vec_ordered_step<NB>(h, S_SWEEP);
        return GPUPCG_OK;
    }
    const FuseXR fx = fuseXR ? FuseXR{h->vX, h->vPA, zb, h->vRA, h->vHistory} : none;
    h->launches += 2;
    const size_t smF = fuseXR ? h->smemFwdV : h->smemFacV;
    double* part = reduce ? h->vPartials : nullptr;
    if (h->vSwizzle)
    {
        k_dic_tile_v<NB, 0, true><<<h->sweepGrid, SWEEP_TPB, smF, h->stream>>>(h->tiles, h->vRD, rA, nullptr, yb, nullptr,
            h->dTileTicket, h->vPartials, h->partStride, h->vState, gated, h->dTabCF, h->dTabUF, fx, h->historyCap);
        if (fuseXR) { int rco = vec_ordered_step<NB>(h, S_XR); if (rco) return rco; }
        k_dic_tile_v<NB, 2, true><<<h->sweepGrid, SWEEP_TPB, h->smemBwdV, h->stream>>>(h->tiles, h->vRD, rA, yb, zb, nullptr,
            h->dTileTicket, part, h->partStride, h->vState, gated, h->dTabCB, h->dTabUB, none, 0);
    }
    else
    {
        k_dic_tile_v<NB, 0, false><<<h->sweepGrid, SWEEP_TPB, smF, h->stream>>>(h->tiles, h->vRD, rA, nullptr, yb, nullptr,
            h->dTileTicket, h->vPartials, h->partStride, h->vState, gated, h->dTabCF, h->dTabUF, fx, h->historyCap);
        if (fuseXR) { int rco = vec_ordered_step<NB>(h, S_XR); if (rco) return rco; }
        k_dic_tile_v<NB, 2, false><<<h->sweepGrid, SWEEP_TPB, h->smemBwdV, h->stream>>>(h->tiles, h->vRD, rA, yb, zb, nullptr,
            h->dTileTicket, part, h->partStride, h->vState, gated, h->dTabCB, h->dTabUB, none, 0);
    }
    CK(cudaGetLastError());
    if (reduce) return vec_ordered_step<NB>(h, S_SWEEP);
    return GPUPCG_OK;
}

// PCG::solve of NB components on resident data: h->vX (in/out), h->vB
template <int NB>
int vec_solve_core(gpupcg_handle h, double tol, double relTol, int minIter, int maxIter, gpupcg_perf* perf,
                   double* history, int historyCap)
{
    const HostPlan& P = h->plan;
    cudaStream_t s = h->stream;
    const long long nFlat = (long long)P.nRows*VecStride<NB>::v;
    const int gN = grid_for(nFlat, h->streamGrid);
    int rc;

    DevState init;
    memset(&init, 0, sizeof(init));
    init.tol = tol; init.relTol = relTol; init.minIter = minIter; init.maxIter = maxIter;
    init.nGlobalCells = h->comm.active() ? h->comm.globalCells(P.nCells) : P.nCells;
    init.historyCap = std::min(h->historyCap, std::max(0, maxIter));
    init.multiRank = h->comm.active() ? 1 : 0;
    init.p2p = h->comm.active() ? h->dP2P : nullptr;
    init.ordered = h->orderedSums ? 1 : 0;
    init.wArA = K_MGREAT; init.wArAold = K_MGREAT;
    DevState* fin = h->hvRing + 2*NB;
    for (int k = 0; k < NB; k++) fin[k] = init;
    CK(cudaMemcpyAsync(h->vState, fin, sizeof(DevState)*NB, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(h->ev[0], s));

    rc = vec_amul<NB>(h, 0, h->vX, h->vWA, false); if (rc) return rc;
    LAUNCH((k_init_residual_v<NB>), gN, s, h->vB, h->vWA, h->vX, h->vRA, nFlat, h->vPartials, h->vState);
    rc = vec_ordered_step<NB>(h, S_INIT); if (rc) return rc;
    rc = vec_amul<NB>(h, 2, h->vX, h->vPA, false); if (rc) return rc;
    LAUNCH((k_norm_v<NB>), gN, s, h->vWA, h->vPA, h->vB, h->vRA, nFlat, h->vPartials, h->vState);
    rc = vec_ordered_step<NB>(h, S_NORM); if (rc) return rc;
    rc = vec_factor<NB>(h, 1); if (rc) return rc;
    vec_fill_sent<NB>(h, h->vWA);
    CK(cudaGetLastError());

    // host loop one batch ahead of the device (see solve_core): every kernel is gated on "all components done"
    const int batch = std::max(1, env_int("GPUPCG_CHECK_EVERY", 8));
    // The x/r update rides in the staging of the next forward sweep while the sweeps are latency-bound (their CTAs
    // have bandwidth to spare); on large sub-domains the sweeps are throughput-bound and a streaming kernel is cheaper.
    const int fuseEnv = env_int("GPUPCG_FUSE_XR", -1);
    // (pencil sweeps: the update can ride along too -- an extra warp per CTA, kernels_pencil.cuh FUSE -- but measured it only
    // moves the 59 us of k_xr_update_v into the forward sweep: off unless GPUPCG_FUSE_XR=1; profiles/r2_pencil_sweeps.md)
    const bool fuseXR = (fuseEnv >= 0) ? (fuseEnv != 0) : (!h->pencil && P.nRows <= env_int("GPUPCG_FUSE_XR_MAXROWS", 3000000));
    auto all_done_host = [&](const DevState* st) { bool d = true; for (int k = 0; k < NB; k++) d = d && st[k].done; return d; };
    auto enqueue_check = [&](int slot) -> cudaError_t
    {
        cudaError_t e = cudaMemcpyAsync(h->hvRing + slot*NB, h->vState, sizeof(DevState)*NB, cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) return e;
        return cudaEventRecord(h->evCheck[slot], s);
    };
    int launched = 0, slot = 0;
    CK(enqueue_check(slot));
    while (true)
    {
        const bool canLaunch = launched < maxIter;
        if (canLaunch)
        {
            const int n = std::min(batch, maxIter - launched);
            for (int i = 0; i < n; i++)
            {
                rc = vec_sweeps<NB>(h, h->vRA, h->vWA, true, 1, fuseXR); if (rc) return rc;
                LAUNCH((k_p_update_v<NB>), gN, s, h->vWA, h->vPA, nFlat, h->vState);
                rc = vec_amul<NB>(h, 1, h->vPA, h->vWA, false); if (rc) return rc;
                rc = vec_ordered_step<NB>(h, S_AMUL); if (rc) return rc;
                if (!fuseXR)
                {
                    LAUNCH((k_xr_update_v<NB>), gN, s, h->vPA, (unsigned long long*)h->vWA, h->vX, h->vRA, nFlat, h->vPartials,
                           h->vState, h->vHistory, h->historyCap);
                    rc = vec_ordered_step<NB>(h, S_XR); if (rc) return rc;
                }
            }
            CK(cudaGetLastError());
            launched += n;
            CK(enqueue_check(slot ^ 1));
        }
        CK(cudaEventSynchronize(h->evCheck[slot]));
        if (all_done_host(h->hvRing + slot*NB))
        {
            if (canLaunch) { slot ^= 1; CK(cudaEventSynchronize(h->evCheck[slot])); }
            break;
        }
        if (!canLaunch) break;
        slot ^= 1;
    }
    // the update of the last iteration launched has not been applied if the loop ran out of iterations
    LAUNCH((k_xr_update_v<NB>), gN, s, h->vPA, (unsigned long long*)h->vWA, h->vX, h->vRA, nFlat, h->vPartials, h->vState,
           h->vHistory, h->historyCap);
    rc = vec_ordered_step<NB>(h, S_XR); if (rc) return rc;
    CK(cudaMemcpyAsync(fin, h->vState, sizeof(DevState)*NB, cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(h->ev[1], s));
    CK(cudaStreamSynchronize(s));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
    for (int k = 0; k < NB; k++)
    {
        const DevState& st = fin[k];
        const int nHist = std::min(std::min(st.nIter, historyCap), h->historyCap);
        if (history && nHist > 0)
            CK(cudaMemcpy(history + (size_t)k*historyCap, h->vHistory + (size_t)k*h->historyCap, sizeof(double)*nHist, cudaMemcpyDeviceToHost));
        if (perf)
        {
            perf[k].initialResidual = st.initialResidual;
            perf[k].finalResidual = st.finalResidual;
            perf[k].normFactor = st.normFactor;
            perf[k].nIterations = st.nIter;
            perf[k].converged = st.converged;
            perf[k].singular = st.singular;
            perf[k].solve_ms = ms;      // the components share the launch sequence: device time of the whole vector solve
        }
    }
    return check_watchdog(h, "gpupcg_solve_vector");
}

// sequential scalar solves of the components on the same kernels (meshes / set-ups the batched kernels do not cover)
template <int NB>
int vec_solve_sequential(gpupcg_handle h, double* d_x, const double* d_b, double tol, double relTol, int minIter,
                         int maxIter, gpupcg_perf* perf, double* history, int historyCap)
{
    const HostPlan& P = h->plan;
    cudaStream_t s = h->stream;
    const int gC = grid_for(P.nCells, h->streamGrid);
    for (int k = 0; k < NB; k++)
    {
        LAUNCH(k_cmpt_get, gC, s, h->vStageD, NB, k, h->vTmpX, P.nCells);
        LAUNCH(k_rows_in, grid_for(P.nRows, h->streamGrid), s, h->vTmpX, h->dDiag, h->dRowOld, P.nRows, 1.0);
        if (P.nIfaceFaces > 0) LAUNCH(k_cmpt_get, grid_for(P.nIfaceFaces, h->streamGrid), s, h->vBou, NB, k, h->dBou, P.nIfaceFaces);
        h->haveMatrix = true;
        LAUNCH(k_cmpt_get, gC, s, d_x, NB, k, h->vTmpX, P.nCells);
        LAUNCH(k_cmpt_get, gC, s, d_b, NB, k, h->vTmpB, P.nCells);
        CK(cudaGetLastError());
        gpupcg_perf p;
        int rc = gpupcg_solve_device(h, h->vTmpX, h->vTmpB, tol, relTol, minIter, maxIter, &p,
                                     history ? history + (size_t)k*historyCap : nullptr, historyCap);
        if (rc) return rc;
        if (perf) perf[k] = p;
        LAUNCH(k_cmpt_put, gC, s, h->vTmpX, NB, k, d_x, P.nCells);
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));
    return GPUPCG_OK;
}

// NB scalar solves in flight at once (see vec_concurrent_now): system k on its own stream, driven by its own host thread
// (solve_core's host loop blocks on its stream's convergence snapshots).
template <int NB>
int vec_solve_concurrent(gpupcg_handle h, double* d_x, const double* d_b, double tol, double relTol, int minIter,
                         int maxIter, gpupcg_perf* perf, double* history, int historyCap)
{
    const HostPlan& P = h->plan;
    for (int k = 0; k < NB; k++)
        if (!h->sys[k] || !h->sys[k]->haveMatrix) { h->err = "gpupcg_solve_vector: component systems have no matrix"; return GPUPCG_ERR_STATE; }
    CK(cudaEventRecord(h->evFork, h->stream));      // d_x, d_b and the matrix are ready on the caller's stream
    int rcs[NB];
    gpupcg_perf ps[NB];
    gpupcg_handle parent = h;
    auto run = [&, parent](int k)
    {
        gpupcg_handle h = parent->sys[k];       // CK / LAUNCH act on the system from here on
        rcs[k] = [&]() -> int
        {
            CK(cudaSetDevice(h->device));
            cudaStream_t s = h->stream;
            CK(cudaStreamWaitEvent(s, parent->evFork, 0));
            const int gC = grid_for(P.nCells, h->streamGrid);
            LAUNCH(k_cmpt_get, gC, s, d_x, NB, k, h->dStageA, P.nCells);
            LAUNCH(k_cmpt_get, gC, s, d_b, NB, k, h->dStageB, P.nCells);
            CK(cudaGetLastError());
            int rc = gpupcg_solve_device(h, h->dStageA, h->dStageB, tol, relTol, minIter, maxIter, &ps[k],
                                         history ? history + (size_t)k*historyCap : nullptr, historyCap);
            if (rc) return rc;
            LAUNCH(k_cmpt_put, gC, s, h->dStageA, NB, k, d_x, P.nCells);
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(s));
            return GPUPCG_OK;
        }();
    };
    std::thread th[NB];
    bool started[NB] = {};
    for (int k = 1; k < NB; k++)
    {
        try { th[k] = std::thread(run, k); started[k] = true; }
        catch (const std::exception&) { started[k] = false; }      // no thread to be had: that system runs after the others, here
    }
    run(0);
    for (int k = 1; k < NB; k++) if (!started[k]) run(k);
    for (int k = 1; k < NB; k++) if (started[k]) th[k].join();
    for (int k = 0; k < NB; k++)
    {
        if (rcs[k]) { h->err = "component " + std::to_string(k) + ": " + h->sys[k]->err; return rcs[k]; }
        if (perf) perf[k] = ps[k];
    }
    return GPUPCG_OK;
}

template <int NB>
int vec_solve_device(gpupcg_handle h, double* d_x, const double* d_b, double tol, double relTol, int minIter, int maxIter,
                     gpupcg_perf* perf, double* history, int historyCap)
{
    const HostPlan& P = h->plan;
    if (h->vNB != NB || !h->vHaveMatrix) { h->err = "gpupcg_solve_vector: set_matrix_vector has not been called for this number of components"; return GPUPCG_ERR_STATE; }
    if (P.nIfaceFaces > 0 && !h->comm.active())
    {
        h->err = "gpupcg_solve_vector: this sub-domain has coupled interfaces but gpupcg_comm_init has not been called";
        return GPUPCG_ERR_STATE;
    }
    if (perf) memset(perf, 0, sizeof(*perf)*NB);
    if (vec_concurrent_now(h) && h->sys[0]) return vec_solve_concurrent<NB>(h, d_x, d_b, tol, relTol, minIter, maxIter, perf, history, historyCap);
    if (!vec_batched_now(h)) return vec_solve_sequential<NB>(h, d_x, d_b, tol, relTol, minIter, maxIter, perf, history, historyCap);
    cudaStream_t s = h->stream;
    const long long nFlat = (long long)P.nRows*VecStride<NB>::v;
    const int gN = grid_for(nFlat, h->streamGrid);
    LAUNCH((k_rows_in_v<NB>), gN, s, d_x, h->vX, h->dRowOld, nFlat, 0.0);
    LAUNCH((k_rows_in_v<NB>), gN, s, d_b, h->vB, h->dRowOld, nFlat, 0.0);
    int rc = vec_solve_core<NB>(h, tol, relTol, minIter, maxIter, perf, history, historyCap);
    if (rc) return rc;
    LAUNCH((k_rows_out_v<NB>), gN, s, h->vX, d_x, h->dRowOld, nFlat);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));
    return GPUPCG_OK;
}

template <int NB>
int vec_time_kernel(gpupcg_handle h, int which, int reps, int flushL2, double* msOut)
{
    if (h->vNB != NB || !h->vHaveMatrix) { h->err = "gpupcg_time_kernel_vector: set_matrix_vector has not been called"; return GPUPCG_ERR_STATE; }
    if (!vec_batched_now(h)) { h->err = "gpupcg_time_kernel_vector: the batched kernels are not in use on this handle"; return GPUPCG_ERR_UNSUPPORTED; }
    const HostPlan& P = h->plan;
    cudaStream_t s = h->stream;
    const long long nFlat = (long long)P.nRows*VecStride<NB>::v;
    const int gN = grid_for(nFlat, h->streamGrid);
    if (flushL2 && !h->dFlush)
    {
        h->flushN = (256LL << 20)/sizeof(float);
        CK(dalloc(h->dFlush, h->flushN));
    }
    DevState init;
    memset(&init, 0, sizeof(init));
    init.maxIter = 1 << 30; init.nGlobalCells = P.nCells; init.normFactor = 1.0; init.tol = 0.0;
    init.wArA = 1.0; init.wArAold = 1.0; init.wApA = 1.0e30; init.nIter = 1; init.initialResidual = 1.0;
    init.pending = (which == 3) ? 1 : 0;
    DevState* fin = h->hvRing + 2*NB;
    for (int k = 0; k < NB; k++) fin[k] = init;
    CK(cudaMemcpyAsync(h->vState, fin, sizeof(DevState)*NB, cudaMemcpyHostToDevice, s));
    int rc = vec_factor<NB>(h, 0); if (rc) return rc;
    LAUNCH((k_rows_in_v<NB>), gN, s, h->vStageD, h->vRA, h->dRowOld, nFlat, 0.0);     // deterministic, non-trivial content
    LAUNCH((k_rows_in_v<NB>), gN, s, h->vStageD, h->vPA, h->dRowOld, nFlat, 0.0);
    CK(cudaMemsetAsync(h->vX, 0, sizeof(double)*nFlat, s));
    vec_fill_sent<NB>(h, h->vWA);
    CK(cudaStreamSynchronize(s));
    double total = 0.0;
    for (int i = 0; i < reps + 1; i++)
    {
        if (which == 1) vec_fill_sent<NB>(h, h->vWA);
        if (which == 3) CK(cudaMemsetAsync(h->vWA, 0, sizeof(double)*nFlat, s));
        if (flushL2) LAUNCH(k_flush, h->streamGrid, s, h->dFlush, h->flushN, (float)i);
        CK(cudaEventRecord(h->ev[0], s));
        switch (which)
        {
            case 0: rc = vec_amul<NB>(h, 1, h->vPA, h->vWA, true); break;
            case 1: rc = vec_sweeps<NB>(h, h->vRA, h->vWA, true, 0, false); break;
            case 2: LAUNCH((k_p_update_v<NB>), gN, s, h->vRA, h->vPA, nFlat, h->vState); rc = 0; break;
            case 3: LAUNCH((k_xr_update_v<NB>), gN, s, h->vPA, (unsigned long long*)h->vWA, h->vX, h->vRA, nFlat,
                           h->vPartials, h->vState, nullptr, 0); rc = 0; break;
            case 4: rc = vec_factor<NB>(h, 0); break;
            default: h->err = "gpupcg_time_kernel_vector: unknown kernel id"; return GPUPCG_ERR_ARG;
        }
        if (rc) return rc;
        CK(cudaEventRecord(h->ev[1], s));
        CK(cudaStreamSynchronize(s));
        CK(cudaGetLastError());
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
        if (i > 0) total += ms;
        CK(cudaMemcpyAsync(h->vState, fin, sizeof(DevState)*NB, cudaMemcpyHostToDevice, s));
    }
    *msOut = total/reps;
    return GPUPCG_OK;
}

} // namespace

#define VEC_DISPATCH(nb, call2, call3)                                                              \
    ((nb) == 2 ? (call2) : (nb) == 3 ? (call3)                                                      \
               : (h->err = "gpupcg vector entry points: nCmpt must be 2 or 3 (use the scalar entry points for 1)", GPUPCG_ERR_ARG))

extern "C" int gpupcg_vector_mode(gpupcg_handle h, int nCmpt)
{
    if (!h) return GPUPCG_ERR_ARG;
    CK(cudaSetDevice(h->device));
    int rc = VEC_DISPATCH(nCmpt, vec_prepare<2>(h), vec_prepare<3>(h));
    if (rc) return rc;
    if (vec_concurrent_now(h)) return 2;
    return vec_batched_now(h) ? 1 : 0;
}

extern "C" int gpupcg_set_matrix_vector_device(gpupcg_handle h, int nCmpt, const double* d_diagCmpt, const double* d_upper,
                                               const double* d_bouCmpt)
{
    if (!h || !d_diagCmpt) return GPUPCG_ERR_ARG;
    CK(cudaSetDevice(h->device));
    return VEC_DISPATCH(nCmpt, vec_set_matrix_device<2>(h, d_diagCmpt, d_upper, d_bouCmpt),
                               vec_set_matrix_device<3>(h, d_diagCmpt, d_upper, d_bouCmpt));
}

extern "C" int gpupcg_set_matrix_vector(gpupcg_handle h, int nCmpt, const double* diagCmpt, const double* upper,
                                        const double* const* patchBouCoeffsCmpt)
{
    if (!h || !diagCmpt) { if (h) h->err = "gpupcg_set_matrix_vector: NULL diag"; return GPUPCG_ERR_ARG; }
    CK(cudaSetDevice(h->device));
    int rc = VEC_DISPATCH(nCmpt, vec_prepare<2>(h), vec_prepare<3>(h));
    if (rc) return rc;
    const HostPlan& P = h->plan;
    cudaStream_t s = h->stream;
    CK(cudaMemcpyAsync(h->vStageD, diagCmpt, sizeof(double)*P.nCells*nCmpt, cudaMemcpyHostToDevice, s));
    if (upper && P.nFaces > 0)
        CK(cudaMemcpyAsync(h->dStageF, upper, sizeof(double)*P.nFaces, cudaMemcpyHostToDevice, s));
    if (h->nPatches > 0)
    {
        if (!patchBouCoeffsCmpt) { h->err = "gpupcg_set_matrix_vector: NULL patchBouCoeffs with coupled patches"; return GPUPCG_ERR_ARG; }
        for (int p = 0; p < h->nPatches; p++)
        {
            const int n = P.patchStart[p + 1] - P.patchStart[p];
            if (n > 0)
                CK(cudaMemcpyAsync(h->vBou + (size_t)P.patchStart[p]*nCmpt, patchBouCoeffsCmpt[p], sizeof(double)*n*nCmpt, cudaMemcpyHostToDevice, s));
        }
    }
    rc = VEC_DISPATCH(nCmpt, vec_set_matrix_device<2>(h, h->vStageD, upper ? h->dStageF : nullptr, h->vBou),
                             vec_set_matrix_device<3>(h, h->vStageD, upper ? h->dStageF : nullptr, h->vBou));
    if (rc) return rc;
    CK(cudaStreamSynchronize(s));       // the caller may free/modify its arrays on return
    return GPUPCG_OK;
}

extern "C" int gpupcg_solve_vector_device(gpupcg_handle h, int nCmpt, double* d_x, const double* d_b, double tol,
                                          double relTol, int minIter, int maxIter, gpupcg_perf* perf, double* history,
                                          int historyCap)
{
    if (!h || !d_x || !d_b) return GPUPCG_ERR_ARG;
    CK(cudaSetDevice(h->device));
    return VEC_DISPATCH(nCmpt, vec_solve_device<2>(h, d_x, d_b, tol, relTol, minIter, maxIter, perf, history, historyCap),
                               vec_solve_device<3>(h, d_x, d_b, tol, relTol, minIter, maxIter, perf, history, historyCap));
}

extern "C" int gpupcg_solve_vector(gpupcg_handle h, int nCmpt, double* x, const double* b, double tol, double relTol,
                                   int minIter, int maxIter, gpupcg_perf* perf, double* history, int historyCap)
{
    if (!h || !x || !b) return GPUPCG_ERR_ARG;
    if (h->vNB != nCmpt || !h->vHaveMatrix) { h->err = "gpupcg_solve_vector: set_matrix_vector has not been called for this number of components"; return GPUPCG_ERR_STATE; }
    const HostPlan& P = h->plan;
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const size_t bytes = sizeof(double)*(size_t)P.nCells*nCmpt;
    CK(cudaEventRecord(h->ev[2], s));
    CK(cudaMemcpyAsync(h->vStageX, x, bytes, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h->vStageB, b, bytes, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(h->ev[3], s));
    int rc = gpupcg_solve_vector_device(h, nCmpt, h->vStageX, h->vStageB, tol, relTol, minIter, maxIter, perf, history, historyCap);
    if (rc) return rc;
    float msIn = 0.f, msOut = 0.f;
    cudaEventElapsedTime(&msIn, h->ev[2], h->ev[3]);
    CK(cudaEventRecord(h->ev[2], s));
    CK(cudaMemcpyAsync(x, h->vStageX, bytes, cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(h->ev[3], s));
    CK(cudaStreamSynchronize(s));
    cudaEventElapsedTime(&msOut, h->ev[2], h->ev[3]);
    if (perf) for (int k = 0; k < nCmpt; k++) { perf[k].h2d_ms = msIn; perf[k].d2h_ms = msOut; }
    return GPUPCG_OK;
}

extern "C" int gpupcg_time_kernel_vector(gpupcg_handle h, int nCmpt, int which, int reps, int flushL2, double* msOut)
{
    if (!h || !msOut || reps < 1) return GPUPCG_ERR_ARG;
    CK(cudaSetDevice(h->device));
    return VEC_DISPATCH(nCmpt, vec_time_kernel<2>(h, which, reps, flushL2, msOut), vec_time_kernel<3>(h, which, reps, flushL2, msOut));
}

// DEqn.solve() on the matrix the last gpupcg_assemble_D left on the device: the matrix never crosses PCIe.
#include "mesh_view.h"

extern "C" int gpupcg_solve_assembled_D(gpupcg_handle h, gpupcg_mesh m, double* D, double tol, double relTol, int minIter,
                                        int maxIter, gpupcg_perf* perf)
{
    if (!h || !m || !D) return GPUPCG_ERR_ARG;
    gpupcg_mesh_view v;
    int rc = gpupcgi_mesh_view(m, &v);
    if (rc) { h->err = std::string("gpupcg_solve_assembled_D: ") + gpupcg_mesh_last_error(m); return rc; }
    const HostPlan& P = h->plan;
    if (v.device != h->device || v.nCells != P.nCells || v.nIF != P.nFaces)
    {
        h->err = "gpupcg_solve_assembled_D: the solver handle was not created for this mesh (device / cells / internal faces differ)";
        return GPUPCG_ERR_ARG;
    }
    if (h->nPatches > 0) { h->err = "gpupcg_solve_assembled_D: coupled (processor) patches are not assembled on the device yet"; return GPUPCG_ERR_UNSUPPORTED; }
    CK(cudaSetDevice(h->device));
    rc = vec_prepare<3>(h); if (rc) return rc;
    cudaStream_t s = h->stream;
    const size_t bytes = sizeof(double)*3*(size_t)P.nCells;
    CK(cudaEventRecord(h->ev[2], s));
    CK(cudaMemcpyAsync(h->vStageX, D, bytes, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(h->ev[3], s));
    // addBoundaryDiag / addBoundarySource on the device, straight into the solver's staging buffers
    LAUNCH(k_add_boundary3, grid_for(P.nCells, h->streamGrid), s, P.nCells, v.cellBfStart, v.cellBf, v.diag, v.source, v.ic, v.bc,
           h->vStageD, h->vStageB);
    CK(cudaGetLastError());
    rc = vec_set_matrix_device<3>(h, h->vStageD, v.upper, nullptr); if (rc) return rc;
    rc = vec_solve_device<3>(h, h->vStageX, h->vStageB, tol, relTol, minIter, maxIter, perf, nullptr, 0); if (rc) return rc;
    float msIn = 0.f, msOut = 0.f;
    cudaEventElapsedTime(&msIn, h->ev[2], h->ev[3]);
    CK(cudaEventRecord(h->ev[2], s));
    CK(cudaMemcpyAsync(D, h->vStageX, bytes, cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(h->ev[3], s));
    CK(cudaStreamSynchronize(s));
    cudaEventElapsedTime(&msOut, h->ev[2], h->ev[3]);
    if (perf) for (int k = 0; k < 3; k++) { perf[k].h2d_ms = msIn; perf[k].d2h_ms = msOut; }
    return GPUPCG_OK;
}


// ---------------------------------------------------------------------------------------------------------------------
// unsTotalLagrangianStress::evolve() (unsTotalLagrangianStress.C:769-1009) with every field resident on the device:
// per outer iteration storePrevIter -> assembly (all branches) -> DEqn.solve() (addBoundaryDiag / addBoundarySource +
// the batched PCG over the valid components) -> correctBoundaryConditions -> D.relax() -> least-squares vol->point ->
// grad / fGrad -> det(I + gradDf) bounds -> residual().  The host sees 6 doubles per outer iteration (the convergence
// logic of :949-971 needs them) and nothing else.
// ---------------------------------------------------------------------------------------------------------------------
template <int NB>
static int evolve_impl(gpupcg_handle h, gpupcg_mesh m, const gpupcg_evolve_controls* ec, gpupcg_evolve_result* res,
                       double* residualHistory, int* pcgIterations, int historyCap)
{
    gpupcg_mesh_view v;
    const HostPlan& P = h->plan;
    int rc = vec_prepare<NB>(h); if (rc) return rc;
    cudaStream_t s = h->stream;
    double* hostScal = nullptr;
    CK(cudaMallocHost((void**)&hostScal, 8*sizeof(double)));
    void* prevStream = nullptr;
    gpupcgi_mesh_bind_stream(m, (void*)s, &prevStream);
    struct Restore { gpupcg_mesh m; void* st; double* hs; ~Restore() { gpupcgi_mesh_bind_stream(m, st, nullptr); cudaFreeHost(hs); } } restore{m, prevStream, hostScal};
    double* dD = nullptr;
    gpupcgi_mesh_D(m, &dD);
    int cm[3] = {0, 1, 2}, nv = 0;
    for (int k = 0; k < 3; k++) if (ec->validCmpt[k]) cm[nv++] = k;
    gpupcg_eqn_controls eq = ec->eqn;
    const bool nonLinear = ec->eqn.nonLinear != 0;
    bool enforceLinear = false, haveView = false;
    int iCorr = 0, totalIts = 0;
    double r = 1.0, maxRes = 0.0, initialResidual = 0.0, lastInit = 0.0;
    float msA = 0.f, msS = 0.f, msE = 0.f, t = 0.f;
    // own events: the solve records h->ev[] itself
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr, e3 = nullptr;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2)); CK(cudaEventCreate(&e3));
    struct EvGuard { cudaEvent_t a, b, c, d; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); cudaEventDestroy(c); cudaEventDestroy(d); } } evGuard{e0, e1, e2, e3};
    std::vector<gpupcg_perf> perf(NB);
    while (true)
    {
        eq.nonLinear = (nonLinear && !enforceLinear) ? 1 : 0;
        CK(cudaEventRecord(e0, s));
        if (ec->incremental)
        {
            rc = gpupcgi_mesh_snapshot_gradDfB(m);
            if (rc) { h->err = std::string("gpupcg_evolve_D: ") + gpupcg_mesh_last_error(m); return rc; }
        }
        rc = gpupcgi_mesh_outer_begin(m, &eq, ec->g);
        if (rc) { h->err = std::string("gpupcg_evolve_D: ") + gpupcg_mesh_last_error(m); return rc; }
        if (!haveView)      // the device pointers of the assembled equation do not change between outer iterations
        {
            rc = gpupcgi_mesh_view(m, &v);
            if (rc) { h->err = std::string("gpupcg_evolve_D: ") + gpupcg_mesh_last_error(m); return rc; }
            if (v.device != h->device || v.nCells != P.nCells || v.nIF != P.nFaces)
            {
                h->err = "gpupcg_evolve_D: the solver handle was not created for this mesh (device / cells / internal faces differ)";
                return GPUPCG_ERR_ARG;
            }
            haveView = true;
        }
        // fvMatrix<vector>::solve: addBoundaryDiag / addBoundarySource per valid component, psi.component(cmpt)
        h->launches++;
        k_add_boundary_v<NB><<<grid_for(P.nCells, h->streamGrid), TPB, 0, s>>>(P.nCells, v.cellBfStart, v.cellBf, v.diag, v.source, v.ic, v.bc,
                                                                               dD, cm[0], cm[1], cm[2], h->vStageD, h->vStageB, h->vStageX);
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1, s));
        rc = vec_set_matrix_device<NB>(h, h->vStageD, v.upper, nullptr); if (rc) return rc;
        rc = vec_solve_device<NB>(h, h->vStageX, h->vStageB, ec->tolerance, ec->relTol, ec->minIter, ec->maxIter, perf.data(), nullptr, 0);
        if (rc) return rc;
        h->launches++;
        k_unpack_cmpts<NB><<<grid_for(P.nCells, h->streamGrid), TPB, 0, s>>>(P.nCells, h->vStageX, cm[0], cm[1], cm[2], dD);
        CK(cudaGetLastError());
        CK(cudaEventRecord(e2, s));
        rc = gpupcgi_mesh_outer_end(m, &eq, ec->relaxD, ec->doRelax, eq.nonLinear, hostScal, ec->incremental);
        if (rc) { h->err = std::string("gpupcg_evolve_D: ") + gpupcg_mesh_last_error(m); return rc; }
        CK(cudaEventRecord(e3, s));
        CK(cudaStreamSynchronize(s));
        cudaEventElapsedTime(&t, e0, e1); msA += t;
        cudaEventElapsedTime(&t, e1, e2); msS += t;
        cudaEventElapsedTime(&t, e2, e3); msE += t;
        // keep the component with the largest initial residual, singular ones ignored (fvMatrixSolve.C)
        const gpupcg_perf* worst = &perf[0];
        for (int k = 1; k < NB; k++)
            if (perf[k].initialResidual > worst->initialResidual && !perf[k].singular) worst = &perf[k];
        for (int k = 0; k < NB; k++) totalIts += perf[k].nIterations;
        if (iCorr == 0) initialResidual = worst->initialResidual;
        lastInit = worst->initialResidual;
        if (eq.nonLinear && (hostScal[5] < 0.01 || hostScal[4] > 100)) enforceLinear = true;       // :939-945
        r = hostScal[0];
        if (r > maxRes) maxRes = r;
        double cur = maxRes*ec->relConvergenceTolerance;
        if (cur < ec->convergenceTolerance) cur = ec->convergenceTolerance;
        if (residualHistory && iCorr < historyCap) residualHistory[iCorr] = r;
        if (pcgIterations && iCorr < historyCap) for (int k = 0; k < NB; k++) pcgIterations[3*iCorr + k] = perf[k].nIterations;
        // total model: while (res > curConvergenceTolerance && ++iCorr < nCorr) (:967-971); incremental model: while (res >
        // convergenceTolerance && ++iCorr < nCorr) (unsIncrTotalLagrangianStress.C:1155-1159)
        if (!(r > (ec->incremental ? ec->convergenceTolerance : cur))) break;
        if (!(++iCorr < ec->nCorrectors)) break;
    }
    if (res)
    {
        res->nOuterIterations = iCorr; res->totalPcgIterations = totalIts; res->enforceLinear = enforceLinear ? 1 : 0;
        res->converged = (nonLinear && enforceLinear) ? 0 : 1;
        res->initialResidual = initialResidual; res->lastSolverInitialResidual = lastInit; res->maxRes = maxRes; res->res = r;
        res->assemble_ms = msA; res->solve_ms = msS; res->update_ms = msE;
    }
    return GPUPCG_OK;
}

extern "C" int gpupcg_evolve_D(gpupcg_handle h, gpupcg_mesh m, const gpupcg_evolve_controls* ec, gpupcg_evolve_result* res,
                               double* residualHistory, int* pcgIterations, int historyCap)
{
    if (!h || !m || !ec) return GPUPCG_ERR_ARG;
    gpupcg_mesh_view v;
    double* dD = nullptr;
    if (gpupcgi_mesh_D(m, &dD) || !dD) return GPUPCG_ERR_ARG;
    const HostPlan& P = h->plan;
    int nv = 0;
    for (int k = 0; k < 3; k++) nv += ec->validCmpt[k] ? 1 : 0;
    if (nv < 2) { h->err = "gpupcg_evolve_D: at least two solution directions are needed (validCmpt)"; return GPUPCG_ERR_UNSUPPORTED; }
    if (h->nPatches > 0) { h->err = "gpupcg_evolve_D: coupled (processor) patches are not assembled on the device yet"; return GPUPCG_ERR_UNSUPPORTED; }
    if (ec->nCorrectors < 1) { h->err = "gpupcg_evolve_D: nCorrectors < 1"; return GPUPCG_ERR_ARG; }
    CK(cudaSetDevice(h->device));
    // the mesh must be the one the solver handle was built for: checked against the view after the first assembly
    (void)v; (void)P;
    return VEC_DISPATCH(nv, evolve_impl<2>(h, m, ec, res, residualHistory, pcgIterations, historyCap),
                            evolve_impl<3>(h, m, ec, res, residualHistory, pcgIterations, historyCap));
}
