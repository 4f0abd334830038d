// gpupcg.cu -- C ABI (include/gpupcg.h) over the sm_100a kernels in kernels.cuh.
// Host orchestration only: no arithmetic of the solve happens on the CPU, and there is no CPU
// fallback -- without a CUDA device every entry point fails with GPUPCG_ERR_NODEVICE.
#include "../../include/gpupcg.h"
#include "kernels.cuh"
#include "kernels_vec.cuh"
#include "kernels_pencil.cuh"
#include "kernels_ordered.cuh"
#include "plan.h"
#include "comm.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <stdexcept>
#include <thread>
#include <vector>

using namespace gpupcg;

static thread_local std::string g_createError;

struct gpupcg_handle_s
{
    int device = 0;
    int numSMs = 0;
    cudaStream_t ownStream = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::string err;

    HostPlan plan;          // index arrays are released after upload; sizes stay
    int nPatches = 0;
    int sweepGrid = 0, streamGrid = 0;
    bool haveMatrix = false;
    long long launches = 0;

    // plan (device)
    SliceInfo* dSlices = nullptr;
    int *dUfaceOld = nullptr, *dRowOld = nullptr;
    int* dPos = nullptr;            // [nCells] cell -> row, only in the ordered-sums verification mode
    bool orderedSums = false;       // GPUPCG_ORDERED_SUMS=1 at create: global sums in the caller's cell order (kernels_ordered.cuh)
    int *dIfRow = nullptr, *dIfRowStart = nullptr, *dIfEntry = nullptr, *dIfFaceRow = nullptr;
    unsigned int* dIfMask = nullptr;
    int nIfRows = 0;
    // tiling of the DIC sweeps (device)
    int *dTileStart = nullptr, *dTileItemPtr = nullptr, *dItems = nullptr;
    unsigned short *dColL16 = nullptr, *dIdxL16 = nullptr, *dColU16 = nullptr;
    int *dExtFPtr = nullptr, *dExtFCol = nullptr, *dExtFIdx = nullptr, *dExtBPtr = nullptr, *dExtBCol = nullptr;
    AmulTileMeta* dAmulMeta = nullptr; unsigned short* dAmulU16 = nullptr; unsigned int* dAmulL32 = nullptr;
    int *dAmulExtX = nullptr, *dAmulExtC = nullptr;
    int amulXStride = AMUL_TPB, amulCStride = AMUL_TPB, nAmulTiles = 0, amulT = 32, amulEU = 0, amulEL = 0;
    // table variant of the sweeps (<= 4 faces per row and side): static index tables + per-matrix coefficient tables
    bool useTab = false;
    int4 *dTabCF = nullptr, *dTabCB = nullptr, *dTabIF = nullptr, *dTabIB = nullptr;
    double *dTabF = nullptr, *dTabB = nullptr;
    size_t smemFwdTab = 0, smemBwdTab = 0;
    // streaming pencil sweeps (structured blocks, plan.h): per-pencil meta, coefficient index / value tables
    bool pencil = false;
    int* dPenMeta = nullptr; int *dPenIdxF = nullptr, *dPenIdxB = nullptr;
    double *dPenCoefF = nullptr, *dPenCoefB = nullptr;
    long long penCoefVersion = -1;
    PencilArgs pen{};
    unsigned int* dSmCounter = nullptr;
    unsigned long long* dTileTicket = nullptr;
    unsigned long long* dTileTrace = nullptr;
    size_t smemFwd = 0, smemBwd = 0, smemAmul = 0;
    long long traceLen = 0;
    DevState* hStateRing = nullptr;     // pinned, two snapshots of the device state (pipelined host loop)
    cudaEvent_t evCheck[2] = {nullptr, nullptr};
    cudaStream_t commStream = nullptr;  // halo swaps, overlapped with the interior Amul (multi-rank only)
    cudaEvent_t evX = nullptr, evHalo = nullptr;
    P2PDev* dP2P = nullptr;            // scalar all-reduces through peer-memory mailboxes (owned by comm)
    std::string p2pNote;
    int amulGrid = 1;               // persistent CTAs of the Amul kernel
    bool hasOvf = false;            // some row has more than four faces on one side (polyhedral cells)
    int partStride = MAXPART;
    TileArgs tiles;

    // matrix + vectors (device, wavefront order unless noted)
    double *dDiag = nullptr, *dUval = nullptr, *dRD = nullptr;
    double *dX = nullptr, *dB = nullptr, *dRA = nullptr, *dWA = nullptr, *dPA = nullptr, *dY = nullptr;
    double *dStageA = nullptr, *dStageB = nullptr;      // [nCells] original order
    double *dStageF = nullptr;                          // [nFaces] original order
    double *dBou = nullptr, *dXNbr = nullptr, *dSend = nullptr;     // [nIfaceFaces] concat order
    double *dPartials = nullptr, *dHistory = nullptr;
    DevState* dState = nullptr;
    DevState* hState = nullptr;     // pinned
    double* hHistory = nullptr;     // pinned
    int historyCap = 0;
    float* dFlush = nullptr;
    long long flushN = 0;

    Comm comm;

    // ---- component-batched ("vector") solve: fvMatrix<vector>::solve's NB systems advanced together (gpupcg_vec.cuh)
    long long uvalVersion = 0, vTabVersion = -1;    // dUval writes / the version the vector path's row tables were built from
    int vNB = 0;                    // components the buffers below are sized for (0: none yet)
    bool vBatched = false;          // batched kernels usable (table variant, shared memory fits); else sequential scalar solves
    bool vHaveMatrix = false;
    int vAgreed = -1;               // multi-rank: 1 every rank runs the batched kernels, 0 none does, -1 not asked yet
    bool vSwizzle = false;          // bank-swizzled sweep variant (small, latency-bound sub-domains)
    size_t smemCap = 0;
    double *vDiag = nullptr, *vRD = nullptr, *vX = nullptr, *vB = nullptr, *vRA = nullptr, *vWA = nullptr, *vPA = nullptr, *vY = nullptr;
    double *vStageX = nullptr, *vStageB = nullptr, *vStageD = nullptr;     // [nCells*NB] caller's order
    double *vTmpX = nullptr, *vTmpB = nullptr;                              // [nCells] sequential fall-back
    double *vBou = nullptr, *vXNbr = nullptr, *vSend = nullptr;            // [nIfaceFaces*NB]
    double *vPartials = nullptr, *vHistory = nullptr;
    double *dTabSq = nullptr, *dTabUF = nullptr, *dTabUB = nullptr;        // [nRows*4] upper^2 / upper per row and side
    DevState* vState = nullptr;
    DevState* hvRing = nullptr;     // pinned: 3 x NB states
    size_t smemFwdV = 0, smemBwdV = 0, smemFacV = 0, smemAmulV = 0;
    int amulGridV = 1;

    // ---- concurrent component solves (gpupcg_vec.cuh, vec_solve_concurrent): the NB systems as NB scalar solves in flight
    // at once, each on its own stream with its own vectors and state, sharing this handle's plan tables and upper().  A
    // system is a handle of its own (isSys) that owns only what make_system allocates.
    gpupcg_handle_s* sys[3] = {nullptr, nullptr, nullptr};
    bool isSys = false;
    cudaEvent_t evFork = nullptr;
    int penSmemKB = 0, penCtasPerSM = 0;        // > 0: ring budget / CTAs per SM of the pencil sweeps of a system (NB systems'
                                                // persistent CTAs must fit an SM together)
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char b_[512];                                                                          \
            snprintf(b_, sizeof(b_), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),       \
                     __FILE__, __LINE__);                                                          \
            h->err = b_;                                                                           \
            return GPUPCG_ERR_CUDA;                                                                \
        }                                                                                          \
    } while (0)

// every kernel launch of the library goes through here (counted: bench.py's gpu_launches)
#define LAUNCH(kern, grid, strm, ...)                                                              \
    do { h->launches++; kern<<<(grid), TPB, 0, (strm)>>>(__VA_ARGS__); } while (0)

template <class T>
static cudaError_t dalloc(T*& p, long long n)
{
    p = nullptr;
    if (n <= 0) n = 1;
    return cudaMalloc((void**)&p, (size_t)n*sizeof(T));
}

template <class T>
static cudaError_t upload(T*& d, const std::vector<T>& v, cudaStream_t s)
{
    cudaError_t e = dalloc(d, (long long)v.size());
    if (e != cudaSuccess || v.empty()) return e;
    return cudaMemcpyAsync(d, v.data(), v.size()*sizeof(T), cudaMemcpyHostToDevice, s);
}

static int env_int(const char* name, int dflt)
{
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

static int grid_for(long long n, int cap)
{
    long long g = (n + TPB - 1)/TPB;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

// ------------------------------------------------------------------------------------------------
extern "C" const char* gpupcg_last_error(gpupcg_handle h)
{
    return h ? h->err.c_str() : g_createError.c_str();
}

extern "C" int gpupcg_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" int gpupcg_device_report(int device, char* buf, int bufLen)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device >= n) return GPUPCG_ERR_NODEVICE;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return GPUPCG_ERR_CUDA;
    if (buf && bufLen > 0)
    {
        snprintf(buf, bufLen, "%s sm_%d%d SMs=%d L2=%dMB mem=%.1fGB", p.name, p.major, p.minor,
                 p.multiProcessorCount, p.l2CacheSize >> 20, p.totalGlobalMem/1073741824.0);
    }
    return GPUPCG_OK;
}

extern "C" int gpupcg_host_register(void* ptr, unsigned long long bytes)
{
    return cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault) == cudaSuccess ? GPUPCG_OK : GPUPCG_ERR_CUDA;
}

extern "C" int gpupcg_host_unregister(void* ptr)
{
    return cudaHostUnregister(ptr) == cudaSuccess ? GPUPCG_OK : GPUPCG_ERR_CUDA;
}

extern "C" int gpupcg_destroy(gpupcg_handle h)
{
    if (!h) return GPUPCG_ERR_ARG;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (auto& c : h->sys) if (c) { gpupcg_destroy(c); c = nullptr; }
    if (h->evFork) cudaEventDestroy(h->evFork);
    if (h->isSys)       // a component system: the plan tables, upper() and the pencil coefficients belong to the parent
    {
        void* own[] = {h->dDiag, h->dRD, h->dX, h->dB, h->dRA, h->dWA, h->dPA, h->dY, h->dStageA, h->dStageB, h->dBou, h->dXNbr,
                       h->dSend, h->dPartials, h->dHistory, h->dState, h->dTileTicket, h->dTabF, h->dTabB, h->dSmCounter};
        for (void* p : own) if (p) cudaFree(p);
        if (h->hState) cudaFreeHost(h->hState);
        if (h->hStateRing) cudaFreeHost(h->hStateRing);
        if (h->hHistory) cudaFreeHost(h->hHistory);
        for (auto& e : h->evCheck) if (e) cudaEventDestroy(e);
        for (auto& e : h->ev) if (e) cudaEventDestroy(e);
        if (h->ownStream) cudaStreamDestroy(h->ownStream);
        delete h;
        return GPUPCG_OK;
    }
    h->comm.destroy();
    void* ptrs[] = {h->dSlices, h->dUfaceOld, h->dRowOld, h->dPos, h->dIfRow,
                    h->dIfRowStart, h->dIfEntry, h->dIfFaceRow, h->dIfMask, h->dDiag, h->dUval, h->dRD,
                    h->dX, h->dB, h->dRA, h->dWA, h->dPA, h->dY, h->dStageA, h->dStageB, h->dStageF,
                    h->dBou, h->dXNbr, h->dSend, h->dPartials, h->dHistory, h->dState, h->dFlush,
                    h->dTileStart, h->dTileItemPtr, h->dItems, h->dColL16, h->dIdxL16, h->dColU16,
                    h->dExtFPtr, h->dExtFCol, h->dExtFIdx, h->dExtBPtr, h->dExtBCol, h->dTileTicket,
                    h->dAmulMeta, h->dAmulU16, h->dAmulL32, h->dAmulExtX, h->dAmulExtC,
                    h->dTabCF, h->dTabCB, h->dTabIF, h->dTabIB, h->dTabF, h->dTabB,
                    h->vDiag, h->vRD, h->vX, h->vB, h->vRA, h->vWA, h->vPA, h->vY, h->vStageX, h->vStageB, h->vStageD,
                    h->vTmpX, h->vTmpB, h->vBou, h->vXNbr, h->vSend, h->vPartials, h->vHistory, h->dTabSq, h->dTabUF,
                    h->dTabUB, h->vState, h->dPenMeta, h->dPenIdxF, h->dPenIdxB, h->dPenCoefF, h->dPenCoefB, h->dSmCounter};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (h->hvRing) cudaFreeHost(h->hvRing);
    if (h->hState) cudaFreeHost(h->hState);
    if (h->hStateRing) cudaFreeHost(h->hStateRing);
    for (auto& e : h->evCheck) if (e) cudaEventDestroy(e);
    if (h->evX) cudaEventDestroy(h->evX);
    if (h->evHalo) cudaEventDestroy(h->evHalo);
    if (h->commStream) cudaStreamDestroy(h->commStream);
    if (h->hHistory) cudaFreeHost(h->hHistory);
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    if (h->ownStream) cudaStreamDestroy(h->ownStream);
    delete h;
    return GPUPCG_OK;
}

// after a synchronised solve: did a device-side wait give up?  (kernels.cuh: SpinWatch)
static int check_watchdog(gpupcg_handle h, const char* where)
{
    unsigned int f = 0;
    CK(cudaMemcpyFromSymbol(&f, g_spinFault, sizeof(f)));
    if (!f) return GPUPCG_OK;
    const unsigned int zero = 0u;
    CK(cudaMemcpyToSymbol(g_spinFault, &zero, sizeof(zero)));
    static const char* what[] = {"", "a peer rank's mailbox flag (cross-rank reduction)", "a cross-tile value (tile sweeps)",
                                 "a cross-tile row (batched tile sweeps)", "a neighbour pencil's row (pencil sweeps)", "a published value"};
    h->err = std::string(where) + ": a device-side wait for " + what[f < 6 ? f : 5] +
             " timed out (GPUPCG_SPIN_TIMEOUT_MS); the kernels ran to their end on zeros and the result is invalid";
    return GPUPCG_ERR_STATE;
}

static int round_up(int v, int m) { return ((std::max(v, 1) + m - 1)/m)*m; }

// The Amul kernel's view of the matrix.  Its tiles are its own: every sweep tile cut into pieces of <= amulRows rows
// (the sweeps want long tiles -- few hops -- while Amul wants many small ones in flight).  Per tile: contiguous
// ranges, operand codes that index the kernel's shared arrays directly (xv = [tile x | cross-tile x: neighbour side,
// then owner side], cv = [tile U block | cross-tile neighbour-side coefficients]) and fixed-stride gather lists.
static int build_amul_view(gpupcg_handle h, cudaStream_t s)
{
    const HostPlan& P = h->plan;
    // 256 rows per Amul tile; 128 on small sub-domains, where more and smaller CTAs hide the fetch latency better
    // (batched Amul 55.1 -> 52.3 us at 1 M cells; 0.283 -> 0.295 ms at 8 M cells, hence by size)
    const int amulRows = std::max(32, (env_int("GPUPCG_AMUL_ROWS", P.nCells <= 3000000 ? 128 : 256)/32)*32);
    std::vector<int> start;                         // first row of every Amul tile, + nRows
    for (int t = 0; t < P.nTiles; t++)
        for (int r = P.tileStart[t]; r < P.tileStart[t + 1]; r += amulRows) start.push_back(r);
    start.push_back(P.nTiles > 0 ? P.tileStart[P.nTiles] : 0);
    const int nT = (int)start.size() - 1;
    auto offU = [&](int sl) { return sl < P.nSlices ? P.slices[sl].offU : (int)P.Ucol.size(); };
    auto offL = [&](int sl) { return sl < P.nSlices ? P.slices[sl].offL : (int)P.Lcol.size(); };
    // pass 1: sizes
    int T = 32, EU = 0, EL = 0, maxX = 0, maxC = 0;
    std::vector<int> nXF(std::max(1, nT), 0), nXB(std::max(1, nT), 0);
    for (int t = 0; t < nT; t++)
    {
        const int r0 = start[t], r1 = start[t + 1];
        const int eU0 = offU(r0/32), eU1 = offU(r1/32), eL0 = offL(r0/32), eL1 = offL(r1/32);
        for (int e = eL0; e < eL1; e++) { const int c = P.Lcol[e]; if (c >= 0 && (c < r0 || c >= r1)) nXF[t]++; }
        for (int e = eU0; e < eU1; e++) { const int c = P.Ucol[e]; if (c >= 0 && (c < r0 || c >= r1)) nXB[t]++; }
        T = std::max(T, r1 - r0); EU = std::max(EU, eU1 - eU0); EL = std::max(EL, eL1 - eL0);
        maxX = std::max(maxX, nXF[t] + nXB[t]); maxC = std::max(maxC, nXF[t]);
    }
    const int XS = round_up(maxX, AMUL_TPB), CS = round_up(maxC, AMUL_TPB);
    if (T + XS >= 0x8000 || EU + CS >= 0x8000)
    {
        h->err = "gpupcg_create: Amul tile too large for its 15-bit operand codes (rows with too many faces)";
        return GPUPCG_ERR_UNSUPPORTED;
    }
    h->nAmulTiles = nT; h->amulT = T; h->amulEU = EU; h->amulEL = EL; h->amulXStride = XS; h->amulCStride = CS;
    // pass 2: codes and gather lists
    std::vector<AmulTileMeta> meta(std::max(1, nT));
    std::vector<unsigned short> U16(P.Ucol.size());
    std::vector<unsigned int> L32(P.Lcol.size());
    std::vector<int> extX((size_t)std::max(1, nT)*XS, -1), extC((size_t)std::max(1, nT)*CS, -1);
    for (int t = 0; t < nT; t++)
    {
        const int r0 = start[t], r1 = start[t + 1];
        const int eU0 = offU(r0/32), eU1 = offU(r1/32), eL0 = offL(r0/32), eL1 = offL(r1/32);
        meta[t] = AmulTileMeta{r0, r1 - r0, eU0, eU1 - eU0, eL0, eL1 - eL0, nXF[t] + nXB[t], nXF[t]};
        int kF = 0, kB = 0;
        for (int e = eL0; e < eL1; e++)
        {
            const int c = P.Lcol[e];
            if (c < 0) L32[e] = 0x80000000u;
            else if (c >= r0 && c < r1) L32[e] = (unsigned)(c - r0) | ((unsigned)(P.Lidx[e] - eU0) << 16);
            else
            {
                extX[(size_t)t*XS + kF] = c;
                extC[(size_t)t*CS + kF] = P.Lidx[e];
                L32[e] = (unsigned)(T + kF) | ((unsigned)(EU + kF) << 16);
                kF++;
            }
        }
        for (int e = eU0; e < eU1; e++)
        {
            const int c = P.Ucol[e];
            if (c < 0) U16[e] = 0x8000u;
            else if (c >= r0 && c < r1) U16[e] = (unsigned short)(c - r0);
            else
            {
                extX[(size_t)t*XS + nXF[t] + kB] = c;
                U16[e] = (unsigned short)(T + nXF[t] + kB);
                kB++;
            }
        }
    }
    CK(upload(h->dAmulMeta, meta, s));
    CK(upload(h->dAmulU16, U16, s));
    CK(upload(h->dAmulL32, L32, s));
    CK(upload(h->dAmulExtX, extX, s));
    CK(upload(h->dAmulExtC, extC, s));
    CK(cudaStreamSynchronize(s));
    return GPUPCG_OK;
}

// Row tables of the table variant, from the plan's 16-bit tile codes: per row and sweep direction four value offsets
// (bytes into the sweep CTA's value array: tile row, cross-tile slot, or the constant slot) and the four U-array
// indices their coefficients come from (-1: unused slot).
static int build_sweep_tables(gpupcg_handle h, cudaStream_t s)
{
    const HostPlan& P = h->plan;
    const int T = P.tileRowsMax, oneF = T + P.maxTileExtF, oneB = T + P.maxTileExtB;
    std::vector<int4> cF(P.nRows), cB(P.nRows), iF(P.nRows), iB(P.nRows);
    for (int t = 0; t < P.nTiles; t++)
    {
        const int r0 = P.tileStart[t], r1 = P.tileStart[t + 1];
        const int eU0 = P.slices[r0/32].offU;
        for (int r = r0; r < r1; r++)
        {
            const SliceInfo& si = P.slices[r/32];
            int c4[4], i4[4];
            for (int j = 0; j < 4; j++)
            {
                c4[j] = 8*oneF; i4[j] = -1;
                if (j >= si.widthL) continue;
                const int e = si.offL + 32*j + (r & 31);
                const unsigned code = P.colL16[e];
                if (code == 0xFFFFu) continue;
                if (code & 0x8000u) { c4[j] = 8*(T + (int)(code & 0x7FFFu)); i4[j] = P.extFIdx[P.extFPtr[t] + (code & 0x7FFFu)]; }
                else { c4[j] = 8*(int)code; i4[j] = eU0 + P.idxL16[e]; }
            }
            cF[r] = make_int4(c4[0], c4[1], c4[2], c4[3]); iF[r] = make_int4(i4[0], i4[1], i4[2], i4[3]);
            for (int j = 0; j < 4; j++)
            {
                c4[j] = 8*oneB; i4[j] = -1;
                if (j >= si.widthU) continue;
                const int e = si.offU + 32*j + (r & 31);
                const unsigned code = P.colU16[e];
                if (code == 0xFFFFu) continue;
                c4[j] = (code & 0x8000u) ? 8*(T + (int)(code & 0x7FFFu)) : 8*(int)code;
                i4[j] = e;
            }
            cB[r] = make_int4(c4[0], c4[1], c4[2], c4[3]); iB[r] = make_int4(i4[0], i4[1], i4[2], i4[3]);
        }
    }
    CK(upload(h->dTabCF, cF, s)); CK(upload(h->dTabCB, cB, s)); CK(upload(h->dTabIF, iF, s)); CK(upload(h->dTabIB, iB, s));
    CK(dalloc(h->dTabF, 4LL*std::max(1, P.nRows))); CK(dalloc(h->dTabB, 4LL*std::max(1, P.nRows)));
    CK(cudaStreamSynchronize(s));
    return GPUPCG_OK;
}

template <int NB> static cudaError_t pencil_attrs(int pLong, size_t cap);      // defined with the pencil launch helpers below

static int create_impl(gpupcg_handle h, int device, int nCells, int nFaces, const int* l, const int* u,
                       int nPatches, const int* patchSizes, const int* const* patchFaceCells,
                       const int* patchNbrRank)
{
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0)
    {
        h->err = "gpupcg_create: no CUDA device visible (this library has no CPU fallback)";
        return GPUPCG_ERR_NODEVICE;
    }
    if (device < 0 || device >= nDev)
    {
        h->err = "gpupcg_create: device index out of range";
        return GPUPCG_ERR_NODEVICE;
    }
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
    {
        h->err = std::string("gpupcg_create: device '") + prop.name +
                 "' is not a Blackwell (sm_100) class GPU; the kernels are built for sm_100a only";
        return GPUPCG_ERR_NODEVICE;
    }
    h->device = device;
    h->numSMs = prop.multiProcessorCount;
    h->nPatches = nPatches;

    // tile size: as large as the shared-memory carve-up of one CTA allows (several CTAs per SM wanted)
    // default tile: 512 rows for bricks of a structured block (16 x 4 x 8), 256 for the greedy tiles of general meshes
    const bool tileRowsGiven = getenv("GPUPCG_TILE_ROWS") && *getenv("GPUPCG_TILE_ROWS");
    // 256 rows (bricks 8 x 4 x 8): what the component-batched sweeps want (their CTAs carry three systems' tiles in
    // shared memory); the scalar sweeps alone are ~8 % faster at 1 M cells with 512 (16 x 4 x 8)
    int tileRows = env_int("GPUPCG_TILE_ROWS", 256);
    const size_t smemCap = (size_t)std::min<long long>(prop.sharedMemPerBlockOptin, 200*1024);
    h->smemCap = smemCap;
    int rc;
    while (true)
    {
        rc = build_plan(nCells, nFaces, l, u, nPatches, patchSizes, patchFaceCells, patchNbrRank, tileRows,
                        env_int("GPUPCG_TILE_MODE", 3), h->plan, h->err);
        if (rc != GPUPCG_OK && rc != -100) return rc;
        if (rc == GPUPCG_OK && h->plan.pencil) { h->pencil = true; h->useTab = false; break; }
        if (rc == GPUPCG_OK && !h->plan.bricks && !tileRowsGiven && tileRows > 256) { tileRows = 256; continue; }
        const HostPlan& Q = h->plan;
        h->smemFwd = sweep_layout(Q.tileRowsMax, Q.maxTileExtF, Q.maxItems, std::max(Q.maxTileEntriesU, Q.maxTileEntriesL), Q.maxTileEntriesL).bytes;
        h->smemBwd = sweep_layout(Q.tileRowsMax, Q.maxTileExtB, Q.maxItems, Q.maxTileEntriesU, Q.maxTileEntriesL).bytes;
        h->useTab = (rc == GPUPCG_OK) && Q.maxRowL <= 4 && Q.maxRowU <= 4 && env_int("GPUPCG_SWEEP_TABLES", 1);
        h->smemFwdTab = tab_layout(Q.tileRowsMax, Q.maxTileExtF, Q.maxItems).bytes;
        h->smemBwdTab = tab_layout(Q.tileRowsMax, Q.maxTileExtB, Q.maxItems).bytes;
        const size_t need = h->useTab ? std::max(h->smemFwdTab, h->smemBwdTab) : std::max(h->smemFwd, h->smemBwd);
        const bool codesFit = true;
        if (rc == GPUPCG_OK && need <= smemCap && codesFit) break;
        if (tileRows <= 32) break;
        tileRows /= 2;
    }
    if (!h->pencil && (rc != GPUPCG_OK || (h->useTab ? std::max(h->smemFwdTab, h->smemBwdTab) : std::max(h->smemFwd, h->smemBwd)) > smemCap))
    {
        h->err = "gpupcg_create: a single 32-row tile does not fit in shared memory (rows with too many faces)";
        return GPUPCG_ERR_UNSUPPORTED;
    }
    HostPlan& P = h->plan;

    CK(cudaStreamCreateWithFlags(&h->ownStream, cudaStreamNonBlocking));
    h->stream = h->ownStream;
    for (auto& e : h->ev) CK(cudaEventCreate(&e));
    {
        int ns = env_int("GPUPCG_SPIN_NS", 0);
        CK(cudaMemcpyToSymbol(g_spinNs, &ns, sizeof(int)));
        // watchdog of the device-side waits (kernels.cuh: SpinWatch)
        const unsigned long long tns = 1000000ULL*(unsigned long long)std::max(0, env_int("GPUPCG_SPIN_TIMEOUT_MS", 20000));
        const unsigned int mask = (unsigned int)std::max(0, env_int("GPUPCG_SPIN_CHECK_MASK", 1023)), zero = 0u;
        CK(cudaMemcpyToSymbol(g_spinTimeoutNs, &tns, sizeof(tns)));
        CK(cudaMemcpyToSymbol(g_spinCheckMask, &mask, sizeof(mask)));
        CK(cudaMemcpyToSymbol(g_spinFault, &zero, sizeof(zero)));
    }
    if (h->pencil)
    {
        const HostPlan& Q = h->plan;
        h->pen.nPencils = Q.nTiles; h->pen.w1 = Q.pW1; h->pen.w2 = Q.pW2; h->pen.n0 = Q.pN0; h->pen.steps = Q.pSteps;
        CK(upload(h->dPenMeta, Q.pMeta, h->stream));
        CK(upload(h->dPenIdxF, Q.pIdxF, h->stream));
        CK(upload(h->dPenIdxB, Q.pIdxB, h->stream));
        CK(dalloc(h->dPenCoefF, 3LL*std::max(1, Q.nRows))); CK(dalloc(h->dPenCoefB, 3LL*std::max(1, Q.nRows)));
        h->pen.meta = h->dPenMeta; h->pen.coefF = h->dPenCoefF; h->pen.coefB = h->dPenCoefB;
        CK(pencil_attrs<1>(Q.pLong, smemCap));
        CK(dalloc(h->dSmCounter, 1024));
        CK(cudaMemsetAsync(h->dSmCounter, 0, sizeof(unsigned int)*1024, h->stream));
        h->pen.smCounter = env_int("GPUPCG_PENCIL_ROTATE", 1) ? h->dSmCounter : nullptr;
        CK(cudaStreamSynchronize(h->stream));
        std::vector<int>().swap(h->plan.pIdxF); std::vector<int>().swap(h->plan.pIdxB);
    }
    else
    {
    CK(cudaFuncSetAttribute(k_dic_tile<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemFwd));
    CK(cudaFuncSetAttribute(k_dic_tile<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemFwd));
    CK(cudaFuncSetAttribute(k_dic_tile<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemBwd));
    CK(cudaFuncSetAttribute(k_dic_tile<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemFwd));
    CK(cudaFuncSetAttribute(k_dic_tile<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemFwd));
    CK(cudaFuncSetAttribute(k_dic_tile<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemBwd));
    h->hasOvf = (P.maxRowL > 4 || P.maxRowU > 4);
    if (h->useTab)
    {
        CK(cudaFuncSetAttribute((k_dic_tile<0, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemFwdTab));
        CK(cudaFuncSetAttribute((k_dic_tile<1, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemFwdTab));
        CK(cudaFuncSetAttribute((k_dic_tile<2, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemBwdTab));
    }
    }   // !pencil
    h->sweepGrid = std::max(1, P.nTiles);
    int sbps = std::max(1, env_int("GPUPCG_STREAM_BPS", 8));
    h->streamGrid = std::min(h->numSMs*sbps, MAXPART);

    cudaStream_t s = h->stream;
    CK(upload(h->dSlices, P.slices, s));
    CK(upload(h->dUfaceOld, P.UfaceOld, s));
    CK(upload(h->dRowOld, P.rowOld, s));
    h->orderedSums = env_int("GPUPCG_ORDERED_SUMS", 0) != 0;
    if (h->orderedSums) CK(upload(h->dPos, P.pos, s));
    CK(upload(h->dIfRow, P.ifRow, s));
    CK(upload(h->dIfRowStart, P.ifRowStart, s));
    CK(upload(h->dIfEntry, P.ifEntry, s));
    CK(upload(h->dIfFaceRow, P.ifFaceRow, s));
    CK(upload(h->dTileStart, P.tileStart, s));
    CK(upload(h->dTileItemPtr, P.tileItemPtr, s));
    CK(upload(h->dItems, P.items, s));
    CK(upload(h->dColL16, P.colL16, s));
    CK(upload(h->dIdxL16, P.idxL16, s));
    CK(upload(h->dColU16, P.colU16, s));
    CK(upload(h->dExtFPtr, P.extFPtr, s));
    CK(upload(h->dExtFCol, P.extFCol, s));
    CK(upload(h->dExtFIdx, P.extFIdx, s));
    CK(upload(h->dExtBPtr, P.extBPtr, s));
    CK(upload(h->dExtBCol, P.extBCol, s));
    { int rcv = build_amul_view(h, s); if (rcv) return rcv; }
    if (h->useTab) { int rcv = build_sweep_tables(h, s); if (rcv) return rcv; }
    {
        h->smemAmul = 2*(size_t)((amul_layout(h->amulT, h->amulEU, h->amulEL, h->amulXStride, h->amulCStride).bytes + 127u) & ~127u);
        if (h->smemAmul > smemCap)
        {
            h->err = "gpupcg_create: an Amul tile does not fit in shared memory (rows with too many faces); lower GPUPCG_AMUL_ROWS";
            return GPUPCG_ERR_UNSUPPORTED;
        }
        CK(cudaFuncSetAttribute(k_amul_tile<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemAmul));
        CK(cudaFuncSetAttribute(k_amul_tile<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemAmul));
        CK(cudaFuncSetAttribute(k_amul_tile<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemAmul));
        int occ = 1;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_amul_tile<1>, AMUL_TPB, h->smemAmul));
        occ = std::max(1, std::min(occ, env_int("GPUPCG_AMUL_BPS", 16)));
        h->amulGrid = std::max(1, std::min(h->nAmulTiles, std::min(h->numSMs*occ, MAXPART)));
    }
    if (env_int("GPUPCG_TILE_TRACE", 0))
    {
        unsigned long long* d = nullptr;
        h->traceLen = 3LL*std::max(1, P.nTiles) + P.nRows + (long long)P.extFCol.size();
        CK(dalloc(d, h->traceLen));
        CK(cudaMemsetAsync(d, 0, sizeof(unsigned long long)*h->traceLen, s));
        CK(cudaMemcpyToSymbol(g_tileTimes, &d, sizeof(d)));
        h->dTileTrace = d;
    }
    CK(dalloc(h->dTileTicket, 1));
    CK(cudaMemsetAsync(h->dTileTicket, 0, sizeof(unsigned long long), s));
    h->nIfRows = (int)P.ifRow.size();
    {
        std::vector<unsigned int> mask(P.nSlices > 0 ? P.nSlices : 1, 0u);
        for (int r : P.ifRow) mask[r >> 5] |= (1u << (r & 31));
        CK(upload(h->dIfMask, mask, s));
        CK(cudaStreamSynchronize(s));
    }

    const long long nR = P.nRows;
    CK(dalloc(h->dDiag, nR));  CK(dalloc(h->dUval, P.entriesU)); CK(dalloc(h->dRD, nR));
    CK(dalloc(h->dX, nR));     CK(dalloc(h->dB, nR));            CK(dalloc(h->dRA, nR));
    CK(dalloc(h->dWA, nR));    CK(dalloc(h->dPA, nR));           CK(dalloc(h->dY, nR));
    CK(dalloc(h->dStageA, nCells)); CK(dalloc(h->dStageB, nCells)); CK(dalloc(h->dStageF, nFaces));
    CK(dalloc(h->dBou, P.nIfaceFaces)); CK(dalloc(h->dXNbr, P.nIfaceFaces)); CK(dalloc(h->dSend, P.nIfaceFaces));
    h->partStride = std::max(MAXPART, P.nTiles);
    CK(dalloc(h->dPartials, 4LL*h->partStride));
    h->historyCap = 4096;
    CK(dalloc(h->dHistory, h->historyCap));
    CK(dalloc(h->dState, 1));
    CK(cudaMallocHost((void**)&h->hState, sizeof(DevState)));
    CK(cudaMallocHost((void**)&h->hStateRing, 2*sizeof(DevState)));
    for (auto& e : h->evCheck) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CK(cudaMallocHost((void**)&h->hHistory, sizeof(double)*h->historyCap));
    memset(h->hState, 0, sizeof(DevState));
    CK(cudaMemsetAsync(h->dState, 0, sizeof(DevState), s));
    CK(cudaMemsetAsync(h->dXNbr, 0, sizeof(double)*std::max(1, P.nIfaceFaces), s));
    CK(cudaStreamSynchronize(s));

    TileArgs T;
    T.slices = h->dSlices; T.tileStart = h->dTileStart; T.tileItemPtr = h->dTileItemPtr; T.items = (const int2*)h->dItems;
    T.colL16 = h->dColL16; T.idxL16 = h->dIdxL16; T.colU16 = h->dColU16;
    T.extFPtr = h->dExtFPtr; T.extFCol = h->dExtFCol; T.extFIdx = h->dExtFIdx;
    T.extBPtr = h->dExtBPtr; T.extBCol = h->dExtBCol;
    T.nTiles = P.nTiles; T.tileRowsMax = P.tileRowsMax; T.maxItems = P.maxItems;
    T.maxEU = P.maxTileEntriesU; T.maxEL = P.maxTileEntriesL; T.maxExtF = P.maxTileExtF; T.maxExtB = P.maxTileExtB;
    T.amulMeta = h->dAmulMeta; T.amulU16 = h->dAmulU16; T.amulL32 = h->dAmulL32;
    T.amulExtX = h->dAmulExtX; T.amulExtC = h->dAmulExtC; T.amulXStride = h->amulXStride; T.amulCStride = h->amulCStride;
    T.nAmulTiles = h->nAmulTiles; T.amulT = h->amulT; T.amulEU = h->amulEU; T.amulEL = h->amulEL;
    h->tiles = T;

    // the big host index arrays are no longer needed
    std::vector<int>().swap(P.Ucol); std::vector<int>().swap(P.UfaceOld);
    std::vector<int>().swap(P.Lidx); std::vector<int>().swap(P.Lcol);
    std::vector<int>().swap(P.LcolTile); std::vector<int>().swap(P.UcolTile);
    std::vector<int>().swap(P.extFCol); std::vector<int>().swap(P.extBCol); std::vector<int>().swap(P.extFIdx);
    std::vector<unsigned short>().swap(P.colL16); std::vector<unsigned short>().swap(P.idxL16);
    std::vector<unsigned short>().swap(P.colU16);
    std::vector<int>().swap(P.roundStart); std::vector<int>().swap(P.items);
    return GPUPCG_OK;
}

extern "C" int gpupcg_create(gpupcg_handle* out, int device, int nCells, int nFaces,
                             const int* lowerAddr, const int* upperAddr, int nPatches,
                             const int* patchSizes, const int* const* patchFaceCells, const int* patchNbrRank)
{
    if (!out) { g_createError = "gpupcg_create: NULL handle pointer"; return GPUPCG_ERR_ARG; }
    *out = nullptr;
    gpupcg_handle h = new gpupcg_handle_s();
    int rc = create_impl(h, device, nCells, nFaces, lowerAddr, upperAddr, nPatches, patchSizes,
                         patchFaceCells, patchNbrRank);
    if (rc != GPUPCG_OK)
    {
        g_createError = h->err;
        gpupcg_destroy(h);
        return rc;
    }
    *out = h;
    return GPUPCG_OK;
}

// A component system of the concurrent vector solve: a handle that shares h's plan tables (read-only), upper() and the
// pencil coefficient tables, and owns its stream, events, vectors, state and -- table variant -- the per-matrix coefficient
// tables.  Single-rank handles only (no interfaces).
static int make_system(gpupcg_handle h, gpupcg_handle& out, int nSystems)
{
    out = nullptr;
    gpupcg_handle c = new gpupcg_handle_s();
    struct Guard { gpupcg_handle c, p; ~Guard() { if (c) { p->err = "component system: " + c->err; gpupcg_destroy(c); } } } guard{c, h};
    c->isSys = true;
    c->device = h->device; c->numSMs = h->numSMs; c->nPatches = 0;
    c->sweepGrid = h->sweepGrid; c->streamGrid = h->streamGrid;
    {   // the sizes of the plan, none of its arrays
        const HostPlan& P = h->plan; HostPlan& Q = c->plan;
        Q.nCells = P.nCells; Q.nFaces = P.nFaces; Q.nRows = P.nRows; Q.nSlices = P.nSlices; Q.nLevels = P.nLevels;
        Q.bricks = P.bricks; Q.pencil = P.pencil; Q.pLong = P.pLong; Q.pStride1 = P.pStride1;
        Q.pW1 = P.pW1; Q.pW2 = P.pW2; Q.pN0 = P.pN0; Q.pSteps = P.pSteps;
        Q.maxLevelRows = P.maxLevelRows; Q.entriesU = P.entriesU; Q.entriesL = P.entriesL;
        Q.tileRowsMax = P.tileRowsMax; Q.nTiles = P.nTiles; Q.maxRounds = P.maxRounds;
        Q.maxTileEntriesL = P.maxTileEntriesL; Q.maxTileEntriesU = P.maxTileEntriesU;
        Q.maxTileExtF = P.maxTileExtF; Q.maxTileExtB = P.maxTileExtB; Q.maxRowL = P.maxRowL; Q.maxRowU = P.maxRowU;
        Q.maxItems = P.maxItems; Q.nIfaceFaces = 0;
    }
    // shared, read-only
    c->dSlices = h->dSlices; c->dUfaceOld = h->dUfaceOld; c->dRowOld = h->dRowOld; c->dPos = h->dPos; c->orderedSums = h->orderedSums;
    c->dIfRow = h->dIfRow; c->dIfRowStart = h->dIfRowStart; c->dIfEntry = h->dIfEntry; c->dIfFaceRow = h->dIfFaceRow; c->dIfMask = h->dIfMask;
    c->nIfRows = 0;
    c->amulXStride = h->amulXStride; c->amulCStride = h->amulCStride; c->nAmulTiles = h->nAmulTiles; c->amulT = h->amulT;
    c->amulEU = h->amulEU; c->amulEL = h->amulEL; c->amulGrid = h->amulGrid; c->smemAmul = h->smemAmul;
    c->useTab = h->useTab; c->dTabCF = h->dTabCF; c->dTabCB = h->dTabCB; c->dTabIF = h->dTabIF; c->dTabIB = h->dTabIB;
    c->smemFwdTab = h->smemFwdTab; c->smemBwdTab = h->smemBwdTab; c->smemFwd = h->smemFwd; c->smemBwd = h->smemBwd;
    c->smemCap = h->smemCap; c->hasOvf = h->hasOvf; c->partStride = h->partStride; c->tiles = h->tiles;
    c->pencil = h->pencil; c->dPenMeta = h->dPenMeta; c->dPenIdxF = h->dPenIdxF; c->dPenIdxB = h->dPenIdxB;
    c->dPenCoefF = h->dPenCoefF; c->dPenCoefB = h->dPenCoefB; c->pen = h->pen;
    c->dUval = h->dUval;
    // the systems' persistent sweep CTAs must be resident together: nSystems rings per SM, one CTA per SM and system
    c->penSmemKB = std::max(24, std::min(100, (int)(h->smemCap/1024)/std::max(1, nSystems) - 2));
    c->penCtasPerSM = 1;
    h = c;      // CK / LAUNCH below act on the system
    CK(cudaStreamCreateWithFlags(&h->ownStream, cudaStreamNonBlocking));
    h->stream = h->ownStream;
    for (auto& e : h->ev) CK(cudaEventCreate(&e));
    for (auto& e : h->evCheck) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    const HostPlan& P = h->plan;
    const long long nR = P.nRows, nC = std::max(1, P.nCells);
    CK(dalloc(h->dDiag, nR)); CK(dalloc(h->dRD, nR));
    CK(dalloc(h->dX, nR));    CK(dalloc(h->dB, nR));  CK(dalloc(h->dRA, nR));
    CK(dalloc(h->dWA, nR));   CK(dalloc(h->dPA, nR)); CK(dalloc(h->dY, nR));
    CK(dalloc(h->dStageA, nC)); CK(dalloc(h->dStageB, nC));
    CK(dalloc(h->dBou, 1)); CK(dalloc(h->dXNbr, 1)); CK(dalloc(h->dSend, 1));
    CK(dalloc(h->dPartials, 4LL*h->partStride));
    h->historyCap = 4096;
    CK(dalloc(h->dHistory, h->historyCap));
    CK(dalloc(h->dState, 1));
    CK(cudaMallocHost((void**)&h->hState, sizeof(DevState)));
    CK(cudaMallocHost((void**)&h->hStateRing, 2*sizeof(DevState)));
    CK(cudaMallocHost((void**)&h->hHistory, sizeof(double)*h->historyCap));
    memset(h->hState, 0, sizeof(DevState));
    CK(cudaMemsetAsync(h->dState, 0, sizeof(DevState), h->stream));
    CK(cudaMemsetAsync(h->dXNbr, 0, sizeof(double), h->stream));
    CK(dalloc(h->dTileTicket, 1));
    CK(cudaMemsetAsync(h->dTileTicket, 0, sizeof(unsigned long long), h->stream));
    if (h->pencil)
    {
        // The per-SM role counter is the parent's, shared by the systems (GPUPCG_SYS_OWN_COUNTER=1: one each): the sweep CTAs of
        // the systems that share an SM then take different schedulers for their sweep warps instead of colliding by chance.
        if (env_int("GPUPCG_SYS_OWN_COUNTER", 0))
        {
            CK(dalloc(h->dSmCounter, 1024));
            CK(cudaMemsetAsync(h->dSmCounter, 0, sizeof(unsigned int)*1024, h->stream));
            h->pen.smCounter = h->pen.smCounter ? h->dSmCounter : nullptr;
        }
    }
    else if (h->useTab) { CK(dalloc(h->dTabF, 4LL*std::max(1, P.nRows))); CK(dalloc(h->dTabB, 4LL*std::max(1, P.nRows))); }
    CK(cudaStreamSynchronize(h->stream));
    guard.c = nullptr;
    out = c;
    return GPUPCG_OK;
}

static void fill_info(const HostPlan& P, gpupcg_plan_info* info)
{
    memset(info, 0, sizeof(*info));
    info->nCells = P.nCells; info->nFaces = P.nFaces; info->nRows = P.nRows; info->nSlices = P.nSlices;
    info->nLevels = P.nLevels; info->maxLevelRows = P.maxLevelRows;
    info->entriesUpper = P.entriesU; info->entriesLower = P.entriesL;
    info->nTiles = P.nTiles; info->tileRowsMax = P.tileRowsMax; info->maxRounds = P.maxRounds;
    info->extForward = P.extFPtr.empty() ? 0 : P.extFPtr.back();
    info->extBackward = P.extBPtr.empty() ? 0 : P.extBPtr.back();
    info->roundEntries = (int)P.roundStart.size();
}

extern "C" int gpupcg_plan_export(int nCells, int nFaces, const int* l, const int* u, int tileRows, int tileMode,
                                  gpupcg_plan_info* info, int** arrays)
{
    HostPlan P;
    std::string err;
    int rc = build_plan(nCells, nFaces, l, u, 0, nullptr, nullptr, nullptr, tileRows, tileMode, P, err);
    if (rc != GPUPCG_OK) { g_createError = err; return rc; }
    if (info) fill_info(P, info);
    if (arrays)
    {
        const std::vector<int>* src[GPUPCG_PLAN_NARRAYS] = {
            &P.rowOld, nullptr, &P.Ucol, &P.UfaceOld, &P.Lidx, &P.Lcol, &P.tileStart, &P.tileRoundPtr,
            &P.roundStart, &P.LcolTile, &P.UcolTile, &P.extFPtr, &P.extFCol, &P.extBPtr, &P.extBCol};
        for (int k = 0; k < GPUPCG_PLAN_NARRAYS; k++)
        {
            if (!arrays[k]) continue;
            if (k == 1) memcpy(arrays[k], P.slices.data(), sizeof(SliceInfo)*P.slices.size());
            else if (!src[k]->empty()) memcpy(arrays[k], src[k]->data(), sizeof(int)*src[k]->size());
        }
    }
    return GPUPCG_OK;
}

extern "C" int gpupcg_get_plan_info(gpupcg_handle h, gpupcg_plan_info* info)
{
    if (!h || !info) return GPUPCG_ERR_ARG;
    const HostPlan& P = h->plan;
    memset(info, 0, sizeof(*info));
    info->nCells = P.nCells; info->nFaces = P.nFaces; info->nRows = P.nRows; info->nSlices = P.nSlices;
    info->nLevels = P.nLevels; info->maxLevelRows = P.maxLevelRows;
    info->entriesUpper = P.entriesU; info->entriesLower = P.entriesL;
    info->nInterfaceFaces = P.nIfaceFaces; info->nInterfaces = h->nPatches;
    info->sweepGrid = h->sweepGrid; info->streamGrid = h->streamGrid; info->numSMs = h->numSMs;
    info->nTiles = P.nTiles; info->tileRowsMax = P.tileRowsMax; info->maxRounds = P.maxRounds;
    info->sweepSmemBytes = (int)(h->useTab ? std::max(h->smemFwdTab, h->smemBwdTab) : std::max(h->smemFwd, h->smemBwd));
    info->extForward = h->plan.extFPtr.empty() ? 0 : h->plan.extFPtr.back();
    info->extBackward = h->plan.extBPtr.empty() ? 0 : h->plan.extBPtr.back();
    return GPUPCG_OK;
}

// debug: per-tile {CTA start, staging done, sweep done} globaltimer stamps of the last forward sweep
extern "C" int gpupcg_debug_tile_trace(gpupcg_handle h, unsigned long long* out3n)
{
    if (!h || !out3n || !h->dTileTrace) return GPUPCG_ERR_ARG;
    CK(cudaMemcpy(out3n, h->dTileTrace, sizeof(unsigned long long)*3*h->plan.nTiles, cudaMemcpyDeviceToHost));
    return GPUPCG_OK;
}

#ifdef GPUPCG_DEEPTRACE
// trace build only (tools/): tile stamps, then the publish time of every row, then the delivery time of every
// cross-tile entry of the last forward sweep
extern "C" int gpupcg_debug_deep_trace(gpupcg_handle h, unsigned long long* out, long long cap)
{
    if (!h || !out || !h->dTileTrace || cap < h->traceLen) return GPUPCG_ERR_ARG;
    CK(cudaMemcpy(out, h->dTileTrace, sizeof(unsigned long long)*h->traceLen, cudaMemcpyDeviceToHost));
    return GPUPCG_OK;
}
#endif

extern "C" int gpupcg_launch_count(gpupcg_handle h, long long* count)
{
    if (!h || !count) return GPUPCG_ERR_ARG;
    *count = h->launches;
    for (auto c : h->sys) if (c) *count += c->launches;
    return GPUPCG_OK;
}

extern "C" int gpupcg_set_stream(gpupcg_handle h, void* cudaStream)
{
    if (!h) return GPUPCG_ERR_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    h->stream = cudaStream ? (cudaStream_t)cudaStream : h->ownStream;
    return GPUPCG_OK;
}

// ------------------------------------------------------------------------------------------------
// matrix
// ------------------------------------------------------------------------------------------------
static int matrix_from_staging(gpupcg_handle h, bool haveUpper)
{
    const HostPlan& P = h->plan;
    cudaStream_t s = h->stream;
    LAUNCH(k_rows_in, grid_for(P.nRows, h->streamGrid), s, h->dStageA, h->dDiag, h->dRowOld, P.nRows, 1.0);
    if (haveUpper)
    {
        LAUNCH(k_coeffs_in, grid_for(P.entriesU, h->streamGrid), s, h->dStageF, h->dUval, h->dUfaceOld, P.entriesU);
        h->uvalVersion++;
    }
    CK(cudaGetLastError());
    h->haveMatrix = true;
    return GPUPCG_OK;
}

extern "C" int gpupcg_set_matrix(gpupcg_handle h, const double* diag, const double* upper,
                                 const double* lower, const double* const* patchBouCoeffs)
{
    if (!h || !diag) { if (h) h->err = "gpupcg_set_matrix: NULL diag"; return GPUPCG_ERR_ARG; }
    if (lower && lower != upper)
    {
        h->err = "gpupcg_set_matrix: asymmetric matrix (lower != upper); gpuPCG is a symMatrix solver";
        return GPUPCG_ERR_UNSUPPORTED;
    }
    if (!upper && !h->haveMatrix && !h->vHaveMatrix && h->plan.nFaces > 0)
    {
        h->err = "gpupcg_set_matrix: upper == NULL but no previous coefficients";
        return GPUPCG_ERR_STATE;
    }
    const HostPlan& P = h->plan;
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    CK(cudaMemcpyAsync(h->dStageA, diag, sizeof(double)*P.nCells, cudaMemcpyHostToDevice, s));
    if (upper && P.nFaces > 0)
        CK(cudaMemcpyAsync(h->dStageF, upper, sizeof(double)*P.nFaces, cudaMemcpyHostToDevice, s));
    if (h->nPatches > 0)
    {
        if (!patchBouCoeffs) { h->err = "gpupcg_set_matrix: NULL patchBouCoeffs with coupled patches"; return GPUPCG_ERR_ARG; }
        for (int p = 0; p < h->nPatches; p++)
        {
            const int n = P.patchStart[p + 1] - P.patchStart[p];
            if (n > 0)
                CK(cudaMemcpyAsync(h->dBou + P.patchStart[p], patchBouCoeffs[p], sizeof(double)*n, cudaMemcpyHostToDevice, s));
        }
    }
    int rc = matrix_from_staging(h, upper != nullptr);
    if (rc != GPUPCG_OK) return rc;
    CK(cudaStreamSynchronize(s));      // the caller may free/modify its arrays on return
    return GPUPCG_OK;
}

extern "C" int gpupcg_set_matrix_device(gpupcg_handle h, const double* d_diag, const double* d_upper,
                                        const double* d_bou)
{
    if (!h || !d_diag) return GPUPCG_ERR_ARG;
    const HostPlan& P = h->plan;
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    LAUNCH(k_rows_in, grid_for(P.nRows, h->streamGrid), s, d_diag, h->dDiag, h->dRowOld, P.nRows, 1.0);
    if (d_upper)
    {
        LAUNCH(k_coeffs_in, grid_for(P.entriesU, h->streamGrid), s, d_upper, h->dUval, h->dUfaceOld, P.entriesU);
        h->uvalVersion++;
    }
    else if (!h->haveMatrix && P.nFaces > 0) { h->err = "gpupcg_set_matrix_device: no coefficients yet"; return GPUPCG_ERR_STATE; }
    if (d_bou && P.nIfaceFaces > 0)
        CK(cudaMemcpyAsync(h->dBou, d_bou, sizeof(double)*P.nIfaceFaces, cudaMemcpyDeviceToDevice, s));
    CK(cudaGetLastError());
    h->haveMatrix = true;
    return GPUPCG_OK;
}

// ------------------------------------------------------------------------------------------------
// building blocks (all enqueue on h->stream)
// ------------------------------------------------------------------------------------------------
static inline void fill_sent(gpupcg_handle h, double* p)
{
    LAUNCH(k_fill64, grid_for(h->plan.nRows, h->streamGrid), h->stream, (unsigned long long*)p, SENT, h->plan.nRows);
}

// halo: gather x[faceCells] -> exchange -> dXNbr.  Single rank: nothing to do.
// The swap runs on the side stream h->commStream (processorFvPatchField::initInterfaceMatrixUpdate posts the
// sends before the interior Amul in the reference too): it starts when x is ready on the main stream (evX), the
// interior rows' Amul runs meanwhile on the main stream, and k_iface_update waits for evHalo.
static int halo_exchange(gpupcg_handle h, const double* x, bool* overlapped)
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
    LAUNCH(k_iface_gather, grid_for(P.nIfaceFaces, h->streamGrid), cs, x, h->dIfFaceRow, h->dSend, P.nIfaceFaces);
    if (h->comm.exchange(h->dSend, h->dXNbr, P.patchStart, P.patchNbrRank, cs, h->err) != 0) return GPUPCG_ERR_NCCL;
    if (*overlapped) { CK(cudaEventRecord(h->evHalo, cs)); }
    return GPUPCG_OK;
}

static int allreduce_step(gpupcg_handle h, int step, int count)
{
    if (h->orderedSums)
    {
        // verification mode: the kernel before this call skipped its tree reduction; form the sums in cell order here
        const OrderedVecs V{h->dX, h->dB, h->dRA, h->dWA, h->dPA};
        h->launches++;
        k_ordered_step<1, 1><<<1, ORD_TPB, 0, h->stream>>>(h->dState, step, V, h->dPos, h->plan.nCells, h->dHistory, 0);
        CK(cudaGetLastError());
    }
    if (!h->comm.active()) return GPUPCG_OK;
    if (h->dP2P) return GPUPCG_OK;      // summed across ranks inside the kernel that finished the reduction
    if (h->comm.allreduce_sum((double*)((char*)h->dState + offsetof(DevState, red)), count, h->stream, h->err) != 0)
        return GPUPCG_ERR_NCCL;
    h->launches++; k_scalar_step<<<1, 32, 0, h->stream>>>(h->dState, step, h->dHistory);
    return GPUPCG_OK;
}

// Ax = A x  (+ interfaces).  mode 0/1/2 as k_amul.  xNbrReady: dXNbr already holds neighbour values.
static int enqueue_amul(gpupcg_handle h, int mode, const double* x, double* Ax, bool xNbrReady)
{
    const HostPlan& P = h->plan;
    cudaStream_t s = h->stream;
    const bool ifaces = P.nIfaceFaces > 0;
    int rc;
    bool haloInFlight = false;
    if (ifaces && mode != 2 && !xNbrReady) { rc = halo_exchange(h, x, &haloInFlight); if (rc) return rc; }
    const unsigned int* mask = (ifaces && mode == 1) ? h->dIfMask : nullptr;
    if (h->nAmulTiles > 0)
    {
        h->launches++;
        if (mode == 0)
            k_amul_tile<0><<<h->amulGrid, AMUL_TPB, h->smemAmul, s>>>(h->tiles, h->dDiag, h->dUval, h->dRowOld, mask, x, Ax, h->dPartials, h->partStride, h->dState);
        else if (mode == 1)
            k_amul_tile<1><<<h->amulGrid, AMUL_TPB, h->smemAmul, s>>>(h->tiles, h->dDiag, h->dUval, h->dRowOld, mask, x, Ax, h->dPartials, h->partStride, h->dState);
        else
            k_amul_tile<2><<<h->amulGrid, AMUL_TPB, h->smemAmul, s>>>(h->tiles, h->dDiag, h->dUval, h->dRowOld, mask, x, Ax, h->dPartials, h->partStride, h->dState);
    }
    if (ifaces)
    {
        if (haloInFlight) { CK(cudaStreamWaitEvent(s, h->evHalo, 0)); }
        const int gi = grid_for(h->nIfRows, h->streamGrid);
        if (mode == 1)
            LAUNCH(k_iface_update<1>, gi, s, h->dIfRow, h->dIfRowStart, h->dIfEntry, h->nIfRows, h->dBou, h->dXNbr, 0, x, Ax, h->dPartials, h->dState, 1);
        else
            LAUNCH(k_iface_update<0>, gi, s, h->dIfRow, h->dIfRowStart, h->dIfEntry, h->nIfRows, h->dBou, h->dXNbr, mode == 2 ? 1 : 0, x, Ax, h->dPartials, h->dState, 0);
    }
    CK(cudaGetLastError());
    return GPUPCG_OK;
}

// ---- streaming pencil sweeps (kernels_pencil.cuh): one CTA per pencil -------------------------------------------------
template <int NB, int MODE, int LONG>
static cudaError_t pencil_attr(size_t cap)
{
    cudaError_t e = cudaFuncSetAttribute((k_dic_pencil<NB, MODE, LONG, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap);
    if (!e && NB > 1)
        e = cudaFuncSetAttribute((k_dic_pencil<NB, MODE, LONG, NB>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap);
    if (MODE == 0)                  // forward sweep carrying the x/r update
    {
        if (!e) e = cudaFuncSetAttribute((k_dic_pencil<NB, 0, LONG, 1, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap);
        if (!e && NB > 1) e = cudaFuncSetAttribute((k_dic_pencil<NB, 0, LONG, NB, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap);
    }
    return e;
}
template <int NB>
static cudaError_t pencil_attrs(int pLong, size_t cap)
{
    cudaError_t e = cudaSuccess;
    if (pLong == 0) { e = pencil_attr<NB, 0, 0>(cap); if (!e) e = pencil_attr<NB, 1, 0>(cap); if (!e) e = pencil_attr<NB, 2, 0>(cap); }
    else if (pLong == 1) { e = pencil_attr<NB, 0, 1>(cap); if (!e) e = pencil_attr<NB, 1, 1>(cap); if (!e) e = pencil_attr<NB, 2, 1>(cap); }
    else { e = pencil_attr<NB, 0, 2>(cap); if (!e) e = pencil_attr<NB, 1, 2>(cap); if (!e) e = pencil_attr<NB, 2, 2>(cap); }
    return e;
}

// coefficient tables of the pencil sweeps follow dUval
static int pencil_coeffs(gpupcg_handle h)
{
    if (h->penCoefVersion == h->uvalVersion) return GPUPCG_OK;
    const long long n3 = 3LL*h->plan.nRows;
    LAUNCH(k_pencil_coeffs, grid_for(n3, h->streamGrid), h->stream, n3, h->dPenIdxF, h->dUval, h->dPenCoefF);
    LAUNCH(k_pencil_coeffs, grid_for(n3, h->streamGrid), h->stream, n3, h->dPenIdxB, h->dUval, h->dPenCoefB);
    CK(cudaGetLastError());
    h->penCoefVersion = h->uvalVersion;
    return GPUPCG_OK;
}

template <int NB, int MODE>
static int launch_pencil(gpupcg_handle h, const double* diagOrRD, const double* rA, unsigned long long* y, unsigned long long* out,
                         double* rDout, double* partials, DevState* st, int gated, const FuseXR* fx = nullptr, int histStride = 0)
{
    const bool withDot = (MODE == 2) && partials != nullptr;
    const bool fuse = (MODE == 0) && fx != nullptr && fx->x != nullptr;
    PencilArgs A = h->pen;
    A.fxX = nullptr; A.fxPA = nullptr; A.fxWA = nullptr; A.fxRA = nullptr; A.fxHistory = nullptr; A.fxHistStride = 0;
    if (fuse) { A.fxX = fx->x; A.fxPA = fx->pA; A.fxWA = fx->wA; A.fxRA = fx->rA; A.fxHistory = fx->history; A.fxHistStride = histStride; }
    // ring stages: as many as fit the budget (default ~100 KB: two CTAs per SM), 2 ... 8
    const size_t cap = std::min<size_t>(h->smemCap, (size_t)(h->penSmemKB > 0 ? h->penSmemKB : env_int("GPUPCG_PENCIL_SMEM_KB", 100))*1024);
    int nst = std::max(2, std::min(8, env_int("GPUPCG_PENCIL_STAGES", 8)));
    while (nst > 2 && pen_smem_bytes<NB, MODE>(nst, A.w1, A.w2, withDot, fuse) > cap) nst--;
    A.nStages = nst;
    A.fetchWindow = env_int("GPUPCG_PENCIL_FETCH_WINDOW", 1);     // 1: measured best (profiles/r2_pencil_sweeps.md); 0 = adaptive
    A.fetchSleep = std::max(0, env_int("GPUPCG_PENCIL_FETCH_SLEEP", 0));      // measured: profiles/r2_pencil_sweeps.md (knob matrix)
    const size_t smem = pen_smem_bytes<NB, MODE>(nst, A.w1, A.w2, withDot, fuse);
    if (smem > h->smemCap) { h->err = "pencil sweep: the ring does not fit in shared memory"; return GPUPCG_ERR_UNSUPPORTED; }
    h->launches++;
    // one sweep warp per component (default for the batched solves) or one warp for all of them
    // -- on latency-bound sub-domains (about one pencil per SM); on large ones the sweeps are paced by the hand-offs and by
    // DRAM, and two unsplit CTAs per SM do better (1 M cells: 244 vs 262 us, 8 M: 925 vs 737 us; profiles/r2_pencil_sweeps.md)
    const bool splitDefault = A.nPencils <= 2*h->numSMs;
    const bool split = NB > 1 && env_int("GPUPCG_PENCIL_SPLIT", splitDefault ? 1 : 0) != 0;
    // persistent CTAs, all resident (the kernel's static round-robin over the wavefront order relies on it).  A sweep warp
    // is issue-bound: beyond one per scheduler they slow each other down and, a pencil being paced by the slowest pencil
    // upstream, the whole wavefront with them -- so by default no more CTAs per SM than keep the sweep warps <= 4.
    int occ = 0, grid = 1;
    const int perSmWanted = h->penCtasPerSM > 0 ? h->penCtasPerSM
                          : std::max(1, env_int("GPUPCG_PENCIL_CTAS_PER_SM", split ? std::max(1, 4/NB) : 2));
#define PEN_THREADS(SW_, FUSE_) ((SW_ + 1 + PEN_FW + ((FUSE_) ? 1 : 0))*32)
#define PEN_RUN(LONG_, SW_, FUSE_) do { \
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (k_dic_pencil<NB, MODE, LONG_, SW_, FUSE_>), PEN_THREADS(SW_, FUSE_), smem)); \
        if (occ < 1) { h->err = "pencil sweep: no CTA fits an SM"; return GPUPCG_ERR_UNSUPPORTED; } \
        grid = std::min(A.nPencils, h->numSMs*std::min(occ, perSmWanted)); \
        k_dic_pencil<NB, MODE, LONG_, SW_, FUSE_><<<grid, PEN_THREADS(SW_, FUSE_), smem, h->stream>>>(A, diagOrRD, rA, y, out, rDout, \
            h->dTileTicket, partials, h->partStride, st, gated); } while (0)
#define PEN_DISPATCH(LONG_) do { \
        if (MODE == 0 && fuse) { if (split) PEN_RUN(LONG_, NB, (MODE == 0)); else PEN_RUN(LONG_, 1, (MODE == 0)); } \
        else { if (split) PEN_RUN(LONG_, NB, false); else PEN_RUN(LONG_, 1, false); } } while (0)
    switch (h->plan.pLong)
    {
        case 0: PEN_DISPATCH(0); break;
        case 1: PEN_DISPATCH(1); break;
        default: PEN_DISPATCH(2); break;
    }
#undef PEN_DISPATCH
#undef PEN_RUN
#undef PEN_THREADS
    return GPUPCG_OK;
}

#define LAUNCH_SWEEP(kern, smem, ...)                                                              \
    do { h->launches++; kern<<<h->sweepGrid, SWEEP_TPB, (smem), h->stream>>>(__VA_ARGS__); } while (0)

static int enqueue_factor(gpupcg_handle h, int gated)
{
    if (h->plan.nTiles == 0) return GPUPCG_OK;
    fill_sent(h, h->dY);
    if (h->pencil)
    {
        int rcp = pencil_coeffs(h); if (rcp) return rcp;
        rcp = launch_pencil<1, 1>(h, h->dDiag, nullptr, nullptr, (unsigned long long*)h->dY, h->dRD, nullptr, h->dState, gated);
        if (rcp) return rcp;
    }
    else if (h->useTab)
    {
        const int gT = grid_for(h->plan.nRows, h->streamGrid);
        LAUNCH((k_tab_coeffs<0>), gT, h->stream, h->plan.nRows, h->dTabIF, h->dUval, nullptr, h->dTabF, h->dState, gated);
        LAUNCH_SWEEP((k_dic_tile<1, false, true>), h->smemFwdTab, h->tiles, h->dDiag, h->dUval, nullptr, nullptr,
                     (unsigned long long*)h->dY, h->dRD, h->dTileTicket, nullptr, 0, h->dState, gated, h->dTabCF, h->dTabF);
        LAUNCH((k_tab_coeffs<1>), gT, h->stream, h->plan.nRows, h->dTabIF, h->dUval, h->dRD, h->dTabF, h->dState, gated);
        LAUNCH((k_tab_coeffs<1>), gT, h->stream, h->plan.nRows, h->dTabIB, h->dUval, h->dRD, h->dTabB, h->dState, gated);
    }
    else if (h->hasOvf)
        LAUNCH_SWEEP((k_dic_tile<1, true>), h->smemFwd, h->tiles, h->dDiag, h->dUval, nullptr, nullptr,
                     (unsigned long long*)h->dY, h->dRD, h->dTileTicket, nullptr, 0, h->dState, gated);
    else
        LAUNCH_SWEEP((k_dic_tile<1, false>), h->smemFwd, h->tiles, h->dDiag, h->dUval, nullptr, nullptr,
                     (unsigned long long*)h->dY, h->dRD, h->dTileTicket, nullptr, 0, h->dState, gated);
    fill_sent(h, h->dY);
    CK(cudaGetLastError());
    return GPUPCG_OK;
}

// y (h->dY) and z must hold SENT on entry; y is re-armed by the backward kernel.
static int allreduce_step(gpupcg_handle h, int step, int count);

static int enqueue_sweeps(gpupcg_handle h, const double* rA, double* z, bool reduce, int gated, bool fuseXR = false)
{
    if (h->plan.nTiles == 0) return GPUPCG_OK;
    unsigned long long* yb = (unsigned long long*)h->dY;
    unsigned long long* zb = (unsigned long long*)z;
    double* part = reduce ? h->dPartials : nullptr;
    if (h->pencil)
    {
        int rcp = pencil_coeffs(h); if (rcp) return rcp;
        if (fuseXR)
        {
            // the pending x/r update of the previous iteration rides in the forward sweep (its update warp); S_XR follows
            FuseXR fx{h->dX, h->dPA, zb, h->dRA, h->dHistory};
            rcp = launch_pencil<1, 0>(h, h->dRD, rA, nullptr, yb, nullptr, h->dPartials, h->dState, gated, &fx, 0); if (rcp) return rcp;
            rcp = allreduce_step(h, S_XR, 1); if (rcp) return rcp;
        }
        else { rcp = launch_pencil<1, 0>(h, h->dRD, rA, nullptr, yb, nullptr, nullptr, h->dState, gated); if (rcp) return rcp; }
        rcp = launch_pencil<1, 2>(h, h->dRD, rA, yb, zb, nullptr, part, h->dState, gated); if (rcp) return rcp;
    }
    else if (h->useTab)
    {
        if (fuseXR)
        {
            // the pending x/r update of the previous iteration rides in this sweep's staging; its sum|rA| goes through
            // the same S_XR step (all-reduced first on several ranks)
            FuseXR fx{h->dX, h->dPA, zb, h->dRA, h->dHistory};
            LAUNCH_SWEEP((k_dic_tile<0, false, true>), h->smemFwdTab, h->tiles, h->dRD, h->dUval, rA, nullptr, yb, nullptr,
                         h->dTileTicket, h->dPartials, h->partStride, h->dState, gated, h->dTabCF, h->dTabF, fx);
            int rcx = allreduce_step(h, S_XR, 1); if (rcx) return rcx;
        }
        else
        LAUNCH_SWEEP((k_dic_tile<0, false, true>), h->smemFwdTab, h->tiles, h->dRD, h->dUval, rA, nullptr, yb, nullptr,
                     h->dTileTicket, nullptr, 0, h->dState, gated, h->dTabCF, h->dTabF);
        LAUNCH_SWEEP((k_dic_tile<2, false, true>), h->smemBwdTab, h->tiles, h->dRD, h->dUval, rA, yb, zb, nullptr,
                     h->dTileTicket, part, h->partStride, h->dState, gated, h->dTabCB, h->dTabB);
    }
    else if (h->hasOvf)
    {
        LAUNCH_SWEEP((k_dic_tile<0, true>), h->smemFwd, h->tiles, h->dRD, h->dUval, rA, nullptr, yb, nullptr,
                     h->dTileTicket, nullptr, 0, h->dState, gated);
        LAUNCH_SWEEP((k_dic_tile<2, true>), h->smemBwd, h->tiles, h->dRD, h->dUval, rA, yb, zb, nullptr,
                     h->dTileTicket, part, h->partStride, h->dState, gated);
    }
    else
    {
        LAUNCH_SWEEP((k_dic_tile<0, false>), h->smemFwd, h->tiles, h->dRD, h->dUval, rA, nullptr, yb, nullptr,
                     h->dTileTicket, nullptr, 0, h->dState, gated);
        LAUNCH_SWEEP((k_dic_tile<2, false>), h->smemBwd, h->tiles, h->dRD, h->dUval, rA, yb, zb, nullptr,
                     h->dTileTicket, part, h->partStride, h->dState, gated);
    }
    CK(cudaGetLastError());
    return GPUPCG_OK;
}

// ------------------------------------------------------------------------------------------------
// PCG::solve on resident data: h->dX (in/out), h->dB.
// ------------------------------------------------------------------------------------------------
static int solve_core(gpupcg_handle h, double tol, double relTol, int minIter, int maxIter,
                      gpupcg_perf* perf, double* history, int historyCap)
{
    const HostPlan& P = h->plan;
    if (P.nIfaceFaces > 0 && !h->comm.active())
    {
        h->err = "gpupcg_solve: this sub-domain has coupled interfaces but gpupcg_comm_init has not been called";
        return GPUPCG_ERR_STATE;
    }
    cudaStream_t s = h->stream;
    const int gN = grid_for(P.nRows, h->streamGrid);
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
    *h->hState = init;
    CK(cudaMemcpyAsync(h->dState, h->hState, sizeof(DevState), cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(h->ev[0], s));

    // wA = A x ; rA = b - wA ; xRef = gAverage(x)
    rc = enqueue_amul(h, 0, h->dX, h->dWA, false); if (rc) return rc;
    LAUNCH(k_init_residual, gN, s, h->dB, h->dWA, h->dX, h->dRA, P.nRows, h->dPartials, h->dState);
    rc = allreduce_step(h, S_INIT, 1); if (rc) return rc;
    // pA = A (xRef field) ; normFactor ; initial residual ; stop()?
    rc = enqueue_amul(h, 2, h->dX, h->dPA, false); if (rc) return rc;
    LAUNCH(k_norm, gN, s, h->dWA, h->dPA, h->dB, h->dRA, P.nRows, h->dPartials, h->dState);
    rc = allreduce_step(h, S_NORM, 2); if (rc) return rc;
    // preconditioner::New -> DIC: calcReciprocalD (skipped on device when already converged)
    rc = enqueue_factor(h, 1); if (rc) return rc;
    fill_sent(h, h->dWA);
    CK(cudaGetLastError());

    // Host loop, one batch ahead of the device: batch k+1 is enqueued BEFORE the host waits for the state snapshot
    // taken after batch k-1, so the stream never runs dry while the host looks at `done`.  Every kernel of the
    // iteration is gated on st->done on the device, so a batch enqueued after convergence is a string of no-ops.
    const int batch = std::max(1, env_int("GPUPCG_CHECK_EVERY", 8));
    // The x/r update of an iteration rides in the forward sweep of the next one.  Tile sweeps (table variant): GPUPCG_FUSE_XR,
    // on.  Pencil sweeps of a scalar solve: GPUPCG_FUSE_XR_PENCIL, on for single-rank handles -- the sweep is latency-bound, its
    // update warp streams the update under it: concurrent component solves 47.3 -> 45.5 ms at 1 M cells, 422 -> 398 ms at 8 M,
    // 16.3 -> 15.7 ms at 250 k; one scalar solve 2.4 % (profiles/r2c_concurrent_components.md).
    const bool fuseXR = (h->useTab && env_int("GPUPCG_FUSE_XR", 1))
                     || (h->pencil && env_int("GPUPCG_FUSE_XR_PENCIL", h->comm.active() ? 0 : 1));
    auto enqueue_check = [&](int slot) -> cudaError_t
    {
        cudaError_t e = cudaMemcpyAsync(h->hStateRing + slot, h->dState, sizeof(DevState), cudaMemcpyDeviceToHost, s);
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
                rc = enqueue_sweeps(h, h->dRA, h->dWA, true, 1, fuseXR); if (rc) return rc;
                rc = allreduce_step(h, S_SWEEP, 1); if (rc) return rc;
                LAUNCH(k_p_update, gN, s, h->dWA, h->dPA, P.nRows, h->dState);
                rc = enqueue_amul(h, 1, h->dPA, h->dWA, false); if (rc) return rc;
                rc = allreduce_step(h, S_AMUL, 2); if (rc) return rc;
                if (!fuseXR)
                {
                    LAUNCH(k_xr_update, gN, s, h->dPA, (unsigned long long*)h->dWA, h->dX, h->dRA, P.nRows,
                                                   h->dPartials, h->dState, h->dHistory);
                    rc = allreduce_step(h, S_XR, 1); if (rc) return rc;
                }
            }
            CK(cudaGetLastError());
            launched += n;
            CK(enqueue_check(slot ^ 1));
        }
        CK(cudaEventSynchronize(h->evCheck[slot]));
        if (h->hStateRing[slot].done)
        {
            if (canLaunch)      // the batch enqueued above ran as no-ops; its snapshot is the same, final state
            {
                slot ^= 1;
                CK(cudaEventSynchronize(h->evCheck[slot]));
            }
            break;
        }
        if (!canLaunch) break;  // maxIter launched and the last snapshot has been read
        slot ^= 1;
    }
    *h->hState = h->hStateRing[slot];
    if (fuseXR)
    {
        // the update of the last iteration launched has not been applied if the loop ran out of iterations
        LAUNCH(k_xr_update, gN, s, h->dPA, (unsigned long long*)h->dWA, h->dX, h->dRA, P.nRows,
                                       h->dPartials, h->dState, h->dHistory);
        rc = allreduce_step(h, S_XR, 1); if (rc) return rc;
        CK(cudaMemcpyAsync(h->hState, h->dState, sizeof(DevState), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    CK(cudaEventRecord(h->ev[1], s));
    const DevState& st = *h->hState;
    const int nHist = std::min(std::min(st.nIter, historyCap), h->historyCap);
    if (history && nHist > 0)
    {
        CK(cudaMemcpyAsync(h->hHistory, h->dHistory, sizeof(double)*nHist, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    if (history && nHist > 0) memcpy(history, h->hHistory, sizeof(double)*nHist);
    if (perf)
    {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
        perf->initialResidual = st.initialResidual;
        perf->finalResidual = st.finalResidual;
        perf->normFactor = st.normFactor;
        perf->nIterations = st.nIter;
        perf->converged = st.converged;
        perf->singular = st.singular;
        perf->solve_ms = ms;
    }
    return check_watchdog(h, "gpupcg_solve");
}

extern "C" int gpupcg_solve_device(gpupcg_handle h, double* d_x, const double* d_b, double tol, double relTol,
                                   int minIter, int maxIter, gpupcg_perf* perf, double* history, int historyCap)
{
    if (!h || !d_x || !d_b) return GPUPCG_ERR_ARG;
    if (!h->haveMatrix) { h->err = "gpupcg_solve: set_matrix has not been called"; return GPUPCG_ERR_STATE; }
    const HostPlan& P = h->plan;
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const int gN = grid_for(P.nRows, h->streamGrid);
    if (perf) memset(perf, 0, sizeof(*perf));
    LAUNCH(k_rows_in, gN, s, d_x, h->dX, h->dRowOld, P.nRows, 0.0);
    LAUNCH(k_rows_in, gN, s, d_b, h->dB, h->dRowOld, P.nRows, 0.0);
    int rc = solve_core(h, tol, relTol, minIter, maxIter, perf, history, historyCap);
    if (rc) return rc;
    LAUNCH(k_rows_out, gN, s, h->dX, d_x, h->dRowOld, P.nRows);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));
    return GPUPCG_OK;
}

extern "C" int gpupcg_solve(gpupcg_handle h, double* x, const double* b, double tol, double relTol,
                            int minIter, int maxIter, gpupcg_perf* perf, double* history, int historyCap)
{
    if (!h || !x || !b) return GPUPCG_ERR_ARG;
    if (!h->haveMatrix) { h->err = "gpupcg_solve: set_matrix has not been called"; return GPUPCG_ERR_STATE; }
    const HostPlan& P = h->plan;
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    CK(cudaEventRecord(h->ev[2], s));
    CK(cudaMemcpyAsync(h->dStageA, x, sizeof(double)*P.nCells, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h->dStageB, b, sizeof(double)*P.nCells, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(h->ev[3], s));
    gpupcg_perf p;
    int rc = gpupcg_solve_device(h, h->dStageA, h->dStageB, tol, relTol, minIter, maxIter, &p, history, historyCap);
    if (rc) return rc;
    float msIn = 0.f, msOut = 0.f;
    cudaEventElapsedTime(&msIn, h->ev[2], h->ev[3]);
    CK(cudaEventRecord(h->ev[2], s));
    CK(cudaMemcpyAsync(x, h->dStageA, sizeof(double)*P.nCells, cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(h->ev[3], s));
    CK(cudaStreamSynchronize(s));
    cudaEventElapsedTime(&msOut, h->ev[2], h->ev[3]);
    p.h2d_ms = msIn; p.d2h_ms = msOut;
    if (perf) *perf = p;
    return GPUPCG_OK;
}

// ------------------------------------------------------------------------------------------------
// kernels in isolation
// ------------------------------------------------------------------------------------------------
extern "C" int gpupcg_amul(gpupcg_handle h, const double* x, double* Ax, const double* xNbr)
{
    if (!h || !x || !Ax) return GPUPCG_ERR_ARG;
    if (!h->haveMatrix) { h->err = "gpupcg_amul: set_matrix has not been called"; return GPUPCG_ERR_STATE; }
    const HostPlan& P = h->plan;
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const int gN = grid_for(P.nRows, h->streamGrid);
    CK(cudaMemcpyAsync(h->dStageA, x, sizeof(double)*P.nCells, cudaMemcpyHostToDevice, s));
    LAUNCH(k_rows_in, gN, s, h->dStageA, h->dX, h->dRowOld, P.nRows, 0.0);
    bool nbrReady = false;
    if (P.nIfaceFaces > 0 && xNbr)
    {
        CK(cudaMemcpyAsync(h->dXNbr, xNbr, sizeof(double)*P.nIfaceFaces, cudaMemcpyHostToDevice, s));
        nbrReady = true;
    }
    int rc = enqueue_amul(h, 0, h->dX, h->dWA, nbrReady); if (rc) return rc;
    LAUNCH(k_rows_out, gN, s, h->dWA, h->dStageB, h->dRowOld, P.nRows);
    CK(cudaMemcpyAsync(Ax, h->dStageB, sizeof(double)*P.nCells, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    return GPUPCG_OK;
}

extern "C" int gpupcg_dic_factor(gpupcg_handle h, double* rD)
{
    if (!h || !rD) return GPUPCG_ERR_ARG;
    if (!h->haveMatrix) { h->err = "gpupcg_dic_factor: set_matrix has not been called"; return GPUPCG_ERR_STATE; }
    const HostPlan& P = h->plan;
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    int rc = enqueue_factor(h, 0); if (rc) return rc;
    LAUNCH(k_rows_out, grid_for(P.nRows, h->streamGrid), s, h->dRD, h->dStageB, h->dRowOld, P.nRows);
    CK(cudaMemcpyAsync(rD, h->dStageB, sizeof(double)*P.nCells, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    return GPUPCG_OK;
}

extern "C" int gpupcg_precondition(gpupcg_handle h, const double* rA, double* wA)
{
    if (!h || !rA || !wA) return GPUPCG_ERR_ARG;
    if (!h->haveMatrix) { h->err = "gpupcg_precondition: set_matrix has not been called"; return GPUPCG_ERR_STATE; }
    const HostPlan& P = h->plan;
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const int gN = grid_for(P.nRows, h->streamGrid);
    int rc = enqueue_factor(h, 0); if (rc) return rc;
    CK(cudaMemcpyAsync(h->dStageA, rA, sizeof(double)*P.nCells, cudaMemcpyHostToDevice, s));
    LAUNCH(k_rows_in, gN, s, h->dStageA, h->dRA, h->dRowOld, P.nRows, 0.0);
    fill_sent(h, h->dWA);
    rc = enqueue_sweeps(h, h->dRA, h->dWA, false, 0); if (rc) return rc;
    LAUNCH(k_rows_out, gN, s, h->dWA, h->dStageB, h->dRowOld, P.nRows);
    CK(cudaMemcpyAsync(wA, h->dStageB, sizeof(double)*P.nCells, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    return GPUPCG_OK;
}

extern "C" int gpupcg_time_kernel(gpupcg_handle h, int which, int reps, int flushL2, double* msOut)
{
    if (!h || !msOut || reps < 1) return GPUPCG_ERR_ARG;
    if (!h->haveMatrix) { h->err = "gpupcg_time_kernel: set_matrix has not been called"; return GPUPCG_ERR_STATE; }
    const HostPlan& P = h->plan;
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const int gN = grid_for(P.nRows, h->streamGrid);
    if (flushL2 && !h->dFlush)
    {
        h->flushN = (256LL << 20)/sizeof(float);
        CK(dalloc(h->dFlush, h->flushN));
    }
    // state: not done, benign scalars, deterministic vectors
    DevState init;
    memset(&init, 0, sizeof(init));
    init.maxIter = 1 << 30; init.nGlobalCells = P.nCells; init.normFactor = 1.0; init.tol = 0.0;
    init.wArA = 1.0; init.wArAold = 1.0; init.wApA = 1.0e30; init.nIter = 1; init.initialResidual = 1.0;
    init.pending = (which == 3) ? 1 : 0;
    *h->hState = init;
    CK(cudaMemcpyAsync(h->dState, h->hState, sizeof(DevState), cudaMemcpyHostToDevice, s));
    int rc = enqueue_factor(h, 0); if (rc) return rc;
    LAUNCH(k_rows_in, gN, s, h->dStageA, h->dRA, h->dRowOld, P.nRows, 0.0);   // whatever the staging holds
    LAUNCH(k_rows_in, gN, s, h->dStageA, h->dPA, h->dRowOld, P.nRows, 0.0);
    CK(cudaMemsetAsync(h->dX, 0, sizeof(double)*P.nRows, s));
    fill_sent(h, h->dWA);
    CK(cudaStreamSynchronize(s));

    double total = 0.0;
    for (int i = 0; i < reps + 1; i++)      // first pass is an untimed warm-up
    {
        if (which == 1) fill_sent(h, h->dWA);
        if (which == 3) CK(cudaMemsetAsync(h->dWA, 0, sizeof(double)*P.nRows, s));
        if (flushL2) LAUNCH(k_flush, h->streamGrid, s, h->dFlush, h->flushN, (float)i);
        CK(cudaEventRecord(h->ev[0], s));
        switch (which)
        {
            case 0: rc = enqueue_amul(h, 1, h->dPA, h->dWA, true); break;
            case 1: rc = enqueue_sweeps(h, h->dRA, h->dWA, true, 0); break;
            case 2: LAUNCH(k_p_update, gN, s, h->dRA, h->dPA, P.nRows, h->dState); rc = 0; break;
            case 3: LAUNCH(k_xr_update, gN, s, h->dPA, (unsigned long long*)h->dWA, h->dX, h->dRA, P.nRows,
                                                   h->dPartials, h->dState, nullptr); rc = 0; break;
            case 4: rc = enqueue_factor(h, 0); break;
            default: h->err = "gpupcg_time_kernel: unknown kernel id"; return GPUPCG_ERR_ARG;
        }
        if (rc) return rc;
        CK(cudaEventRecord(h->ev[1], s));
        CK(cudaStreamSynchronize(s));
        CK(cudaGetLastError());
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
        if (i > 0) total += ms;
        // keep the scalar state benign (k_xr_update / k_amul<1> advance it)
        *h->hState = init;
        CK(cudaMemcpyAsync(h->dState, h->hState, sizeof(DevState), cudaMemcpyHostToDevice, s));
    }
    *msOut = total/reps;
    return GPUPCG_OK;
}

// ------------------------------------------------------------------------------------------------
// multi-GPU
// ------------------------------------------------------------------------------------------------
extern "C" int gpupcg_comm_unique_id(void* id128)
{
    return Comm::unique_id(id128) == 0 ? GPUPCG_OK : GPUPCG_ERR_NCCL;
}

extern "C" int gpupcg_comm_init(gpupcg_handle h, int rank, int nRanks, const void* id128)
{
    if (!h || !id128 || rank < 0 || rank >= nRanks) return GPUPCG_ERR_ARG;
    CK(cudaSetDevice(h->device));
    if (h->comm.init(rank, nRanks, id128, h->plan.nCells, h->stream, h->err) != 0) return GPUPCG_ERR_NCCL;
    if (h->comm.emptyRanks() > 0)
    {
        // a rank without rows launches no reducing kernel and would never take part in the per-step reductions (mismatched
        // collectives on the NCCL path, peers spinning on its mailbox flag on the P2P path).  Every rank sees the same count
        // and refuses together.
        h->err = "gpupcg_comm_init: " + std::to_string(h->comm.emptyRanks()) + " rank(s) own a sub-domain without cells; "
                 "decompose over fewer ranks (every rank needs at least one cell)";
        h->comm.destroy();
        return GPUPCG_ERR_UNSUPPORTED;
    }
    h->dP2P = env_int("GPUPCG_P2P_REDUCE", 1) ? h->comm.setup_p2p(h->stream, h->p2pNote) : nullptr;
    if (!h->dP2P && env_int("GPUPCG_P2P_REDUCE", 1) == 2)    // 2: insist
    {
        h->err = "gpupcg_comm_init: " + h->p2pNote;
        return GPUPCG_ERR_NCCL;
    }
    if (!h->commStream && env_int("GPUPCG_OVERLAP_HALO", 1))
    {
        CK(cudaStreamCreateWithFlags(&h->commStream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&h->evX, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->evHalo, cudaEventDisableTiming));
    }
    return GPUPCG_OK;
}

#include "gpupcg_vec.cuh"
