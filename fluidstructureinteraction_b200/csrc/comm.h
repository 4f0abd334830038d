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

class Comm
{
public:
    bool active() const { return comm_ != nullptr; }
    int rank() const { return rank_; }
    int nRanks() const { return nRanks_; }
    long long globalCells(int /*local*/) const { return globalCells_; }

    static int unique_id(void* id128);
    int init(int rank, int nRanks, const void* id128, int nCellsLocal, cudaStream_t s, std::string& err);
    void destroy();

    // send[patchStart[p]..) -> rank nbr[p];  recv[patchStart[p]..) <- rank nbr[p]
    int exchange(const double* send, double* recv, const std::vector<int>& patchStart,
                 const std::vector<int>& nbr, cudaStream_t s, std::string& err);
    int allreduce_sum(double* buf, int count, cudaStream_t s, std::string& err);

private:
    void* comm_ = nullptr;
    int rank_ = 0, nRanks_ = 1;
    long long globalCells_ = 0;
};

} // namespace gpupcg
