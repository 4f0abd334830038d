// comm.h -- the two cross-rank steps of the path, over NCCL on NVLink 5 / NVSwitch:
//   * processorFvPatchField::initInterfaceMatrixUpdate / updateInterfaceMatrix  -> pairwise halo swap
//   * gSumProd / gSumMag / gSum / gAverage (Pstream reduce)                      -> all-reduce of 1-2 doubles
// NCCL is resolved with dlopen at comm_init, so single-GPU use has no NCCL dependency (and in a
// torch process the already loaded torch-bundled libnccl is the one that answers).
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>

namespace gpupcg
{

// One-shot scalar all-reduce over NVLink peer memory (the three gSumProd/gSumMag of a PCG iteration are 8-16 byte
// all-reduces: pure latency).  Every rank owns a mailbox; the thread that finishes a grid reduction writes its partial
// sums straight into every peer's mailbox (P2P stores through NVSwitch, then a release flag = sequence number), waits
// for the peers' flags in its OWN mailbox, and sums the values in rank order -- the same bits on every rank, no
// kernel launch, no NCCL call on the critical path.  Two buffers by sequence parity: a rank can run at most one
// reduction ahead of the slowest one.
static constexpr int P2P_MAX_RANKS = 16;
struct P2PMailbox
{
    double val[2][P2P_MAX_RANKS][8];       // up to 2 sums x 3 components per step (vector solve)
    unsigned long long flag[2][P2P_MAX_RANKS];
};
struct P2PDev
{
    int rank, nRanks;
    unsigned long long seq;
    P2PMailbox* peer[P2P_MAX_RANKS];        // peer[r]: rank r's mailbox mapped into this process (peer[rank]: own)
};

class Comm
{
public:
    bool active() const { return comm_ != nullptr; }
    int rank() const { return rank_; }
    int nRanks() const { return nRanks_; }
    long long globalCells(int /*local*/) const { return globalCells_; }
    int emptyRanks() const { return emptyRanks_; }      // ranks whose sub-domain has no cells (known to every rank after init)

    static int unique_id(void* id128);
    int init(int rank, int nRanks, const void* id128, int nCellsLocal, cudaStream_t s, std::string& err);
    void destroy();

    // send[patchStart[p]..) -> rank nbr[p];  recv[patchStart[p]..) <- rank nbr[p]   (nb values per face, interleaved)
    int exchange(const double* send, double* recv, const std::vector<int>& patchStart,
                 const std::vector<int>& nbr, cudaStream_t s, std::string& err, int nb = 1);
    int allreduce_sum(double* buf, int count, cudaStream_t s, std::string& err);

    // peer-memory mailboxes (cudaIpc handles swapped with one ncclAllGather); returns the device-side descriptor or
    // nullptr when peer mapping is not available (the NCCL all-reduce path stays in use then)
    P2PDev* setup_p2p(cudaStream_t s, std::string& note);

private:
    void* comm_ = nullptr;
    int rank_ = 0, nRanks_ = 1;
    long long globalCells_ = 0;
    int emptyRanks_ = 0;
    P2PMailbox* mbox_ = nullptr;
    P2PDev* p2pDev_ = nullptr;
    std::vector<void*> opened_;
};

} // namespace gpupcg
