// comm.cpp -- see comm.h.  NCCL entry points are bound at run time (dlopen) against the minimal
// prototype set below (stable since NCCL 2.7: ncclSend/ncclRecv).
#include "comm.h"

#include <dlfcn.h>
#include <cstdio>
#include <cstring>

namespace gpupcg
{

namespace
{
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void* ncclComm_p;
enum { NCCL_SUCCESS = 0 };
enum { NCCL_INT8 = 0, NCCL_INT64 = 4, NCCL_FLOAT64 = 8 };    // ncclDataType_t values
enum { NCCL_SUM = 0 };                         // ncclRedOp_t

struct Api
{
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId_t*) = nullptr;
    int (*CommInitRank)(ncclComm_p*, int, ncclUniqueId_t, int) = nullptr;
    int (*CommDestroy)(ncclComm_p) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_p, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};

Api& api()
{
    static Api a;
    if (a.lib) return a;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names)
    {
        a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (a.lib) break;
    }
    if (!a.lib) return a;
#define BIND(field, sym) *(void**)(&a.field) = dlsym(a.lib, sym)
    BIND(GetUniqueId, "ncclGetUniqueId");
    BIND(CommInitRank, "ncclCommInitRank");
    BIND(CommDestroy, "ncclCommDestroy");
    BIND(AllReduce, "ncclAllReduce");
    BIND(AllGather, "ncclAllGather");
    BIND(Send, "ncclSend");
    BIND(Recv, "ncclRecv");
    BIND(GroupStart, "ncclGroupStart");
    BIND(GroupEnd, "ncclGroupEnd");
    BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.Send && a.Recv
        && a.GroupStart && a.GroupEnd;
    return a;
}

int fail(std::string& err, const char* what, int code)
{
    char b[256];
    Api& a = api();
    snprintf(b, sizeof(b), "%s failed: %s", what, (a.GetErrorString ? a.GetErrorString(code) : "nccl error"));
    err = b;
    return -1;
}
} // namespace

int Comm::unique_id(void* id128)
{
    Api& a = api();
    if (!a.ok || !id128) return -1;
    ncclUniqueId_t id;
    if (a.GetUniqueId(&id) != NCCL_SUCCESS) return -1;
    memcpy(id128, &id, sizeof(id));
    return 0;
}

int Comm::init(int rank, int nRanks, const void* id128, int nCellsLocal, cudaStream_t s, std::string& err)
{
    Api& a = api();
    if (!a.ok) { err = "gpupcg_comm_init: libnccl.so.2 not found or incomplete"; return -1; }
    destroy();
    ncclUniqueId_t id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_p c = nullptr;
    int rc = a.CommInitRank(&c, nRanks, id, rank);
    if (rc != NCCL_SUCCESS) return fail(err, "ncclCommInitRank", rc);
    comm_ = c;
    rank_ = rank;
    nRanks_ = nRanks;
    // global cell count (gAverage denominator) and the number of ranks without cells: one int64 all-reduce at init
    long long* d = nullptr;
    if (cudaMalloc((void**)&d, 2*sizeof(long long)) != cudaSuccess) { err = "cudaMalloc failed in comm init"; return -1; }
    long long v[2] = {nCellsLocal, nCellsLocal == 0 ? 1 : 0};
    cudaMemcpyAsync(d, v, sizeof(v), cudaMemcpyHostToDevice, s);
    rc = a.AllReduce(d, d, 2, NCCL_INT64, NCCL_SUM, (ncclComm_p)comm_, s);
    if (rc != NCCL_SUCCESS) { cudaFree(d); return fail(err, "ncclAllReduce(int64)", rc); }
    cudaMemcpyAsync(v, d, sizeof(v), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    cudaFree(d);
    globalCells_ = v[0];
    emptyRanks_ = (int)v[1];
    return 0;
}

P2PDev* Comm::setup_p2p(cudaStream_t s, std::string& note)
{
    Api& a = api();
    if (!comm_ || !a.AllGather || nRanks_ > P2P_MAX_RANKS) { note = "p2p reduce: not available (no ncclAllGather / too many ranks)"; return nullptr; }
    if (p2pDev_) return p2pDev_;
    if (cudaMalloc((void**)&mbox_, sizeof(P2PMailbox)) != cudaSuccess) { note = "p2p reduce: cudaMalloc failed"; cudaGetLastError(); return nullptr; }
    cudaMemsetAsync(mbox_, 0, sizeof(P2PMailbox), s);
    cudaIpcMemHandle_t mine;
    bool ok = cudaIpcGetMemHandle(&mine, mbox_) == cudaSuccess;
    // every rank must take part in the gather even if its own handle failed: a zeroed handle marks the failure
    unsigned char* d = nullptr;
    const size_t hb = sizeof(cudaIpcMemHandle_t) + 8;       // handle + "valid" byte (padded)
    std::vector<unsigned char> sendb(hb, 0), all(hb*nRanks_, 0);
    if (ok) { memcpy(sendb.data(), &mine, sizeof(mine)); sendb[sizeof(mine)] = 1; } else cudaGetLastError();
    if (cudaMalloc((void**)&d, hb*(nRanks_ + 1)) != cudaSuccess) { note = "p2p reduce: cudaMalloc failed"; cudaGetLastError(); return nullptr; }
    cudaMemcpyAsync(d, sendb.data(), hb, cudaMemcpyHostToDevice, s);
    int rc = a.AllGather(d, d + hb, hb, NCCL_INT8, (ncclComm_p)comm_, s);
    if (rc == NCCL_SUCCESS) cudaMemcpyAsync(all.data(), d + hb, hb*nRanks_, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    cudaFree(d);
    if (rc != NCCL_SUCCESS) { note = "p2p reduce: ncclAllGather of the IPC handles failed"; return nullptr; }
    P2PDev h;
    memset(&h, 0, sizeof(h));
    h.rank = rank_; h.nRanks = nRanks_; h.seq = 0;
    bool allOk = true;
    for (int r = 0; r < nRanks_; r++) allOk = allOk && all[hb*r + sizeof(cudaIpcMemHandle_t)] == 1;
    for (int r = 0; allOk && r < nRanks_; r++)
    {
        if (r == rank_) { h.peer[r] = mbox_; continue; }
        cudaIpcMemHandle_t hr;
        memcpy(&hr, &all[hb*r], sizeof(hr));
        void* ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, hr, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); allOk = false; break; }
        opened_.push_back(ptr);
        h.peer[r] = (P2PMailbox*)ptr;
    }
    // all ranks must agree: one more tiny all-reduce of the success flag
    long long* flag = nullptr;
    long long v = allOk ? 1 : 0;
    if (cudaMalloc((void**)&flag, sizeof(long long)) == cudaSuccess)
    {
        cudaMemcpyAsync(flag, &v, sizeof(v), cudaMemcpyHostToDevice, s);
        rc = a.AllReduce(flag, flag, 1, NCCL_INT64, NCCL_SUM, (ncclComm_p)comm_, s);
        cudaMemcpyAsync(&v, flag, sizeof(v), cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s);
        cudaFree(flag);
    }
    else { cudaGetLastError(); v = 0; }
    if (rc != NCCL_SUCCESS || v != nRanks_)
    {
        note = "p2p reduce: peer mapping (cudaIpcOpenMemHandle) not available on every rank; using ncclAllReduce";
        return nullptr;
    }
    if (cudaMalloc((void**)&p2pDev_, sizeof(P2PDev)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    cudaMemcpyAsync(p2pDev_, &h, sizeof(h), cudaMemcpyHostToDevice, s);
    cudaStreamSynchronize(s);
    return p2pDev_;
}

void Comm::destroy()
{
    for (void* p : opened_) cudaIpcCloseMemHandle(p);
    opened_.clear();
    if (p2pDev_) { cudaFree(p2pDev_); p2pDev_ = nullptr; }
    if (mbox_) { cudaFree(mbox_); mbox_ = nullptr; }
    if (comm_)
    {
        api().CommDestroy((ncclComm_p)comm_);
        comm_ = nullptr;
    }
}

int Comm::exchange(const double* send, double* recv, const std::vector<int>& patchStart,
                   const std::vector<int>& nbr, cudaStream_t s, std::string& err, int nb)
{
    Api& a = api();
    int rc = a.GroupStart();
    if (rc != NCCL_SUCCESS) return fail(err, "ncclGroupStart", rc);
    for (size_t p = 0; p + 1 < patchStart.size(); p++)
    {
        const size_t n = (size_t)nb*(size_t)(patchStart[p + 1] - patchStart[p]);
        if (n == 0 || nbr[p] < 0) continue;
        const size_t off = (size_t)nb*(size_t)patchStart[p];
        rc = a.Send(send + off, n, NCCL_FLOAT64, nbr[p], (ncclComm_p)comm_, s);
        if (rc != NCCL_SUCCESS) return fail(err, "ncclSend", rc);
        rc = a.Recv(recv + off, n, NCCL_FLOAT64, nbr[p], (ncclComm_p)comm_, s);
        if (rc != NCCL_SUCCESS) return fail(err, "ncclRecv", rc);
    }
    rc = a.GroupEnd();
    if (rc != NCCL_SUCCESS) return fail(err, "ncclGroupEnd", rc);
    return 0;
}

int Comm::allreduce_sum(double* buf, int count, cudaStream_t s, std::string& err)
{
    Api& a = api();
    int rc = a.AllReduce(buf, buf, (size_t)count, NCCL_FLOAT64, NCCL_SUM, (ncclComm_p)comm_, s);
    if (rc != NCCL_SUCCESS) return fail(err, "ncclAllReduce", rc);
    return 0;
}

} // namespace gpupcg
