/*---------------------------------------------------------------------------*\
    gpuPCG.C -- foam-extend-3.1 lduMatrix::solver adaptor over include/gpupcg.h
\*---------------------------------------------------------------------------*/

#include "gpuPCG.H"
#include "gpupcg.h"

#include <cstdlib>
#include <map>
#include <vector>

#ifndef FOAMLIKE
#include "processorLduInterfaceField.H"
#include "Pstream.H"
#endif

// * * * * * * * * * * * * * * Static Data Members * * * * * * * * * * * * * //

namespace Foam
{
    defineTypeNameAndDebug(gpuPCG, 0);

    lduMatrix::solver::addsymMatrixConstructorToTable<gpuPCG>
        addgpuPCGSymMatrixConstructorToTable_;


// One device handle per lduAddressing: fvMatrix::solve constructs a fresh solver object for every
// component of every outer iteration, the mesh (and so the schedule built by gpupcg_create) lives on.
struct gpuPCGCacheEntry
{
    gpupcg_handle handle;
    label nCells;
    label nFaces;
    const scalar* upperPtr;
    scalar upperCheck;
    bool commReady;
};

static std::map<const lduAddressing*, gpuPCGCacheEntry> gpuPCGCache_;

// Hash of EVERY coefficient of upper() (four interleaved FNV-1a lanes over the 64-bit words, position dependent): one
// O(F) host pass, a fraction of the H2D copy it can save.  fvMatrix builds a fresh lduMatrix every outer iteration and
// the allocator routinely hands back the same address, so the pointer alone says nothing about the values.
static scalar upperHash(const scalarField& upper)
{
    const label n = upper.size();
    const unsigned long long* w = reinterpret_cast<const unsigned long long*>(upper.begin());
    const unsigned long long prime = 1099511628211ULL;
    unsigned long long h0 = 14695981039346656037ULL, h1 = h0 ^ 0x9E3779B97F4A7C15ULL, h2 = h0 ^ 0xC2B2AE3D27D4EB4FULL,
                       h3 = h0 ^ 0x165667B19E3779F9ULL;
    label i = 0;
    for (; i + 4 <= n; i += 4)
    {
        h0 = (h0 ^ w[i])*prime; h1 = (h1 ^ w[i + 1])*prime; h2 = (h2 ^ w[i + 2])*prime; h3 = (h3 ^ w[i + 3])*prime;
    }
    for (; i < n; i++) h0 = (h0 ^ w[i])*prime;
    unsigned long long h = (h0 ^ (h1*prime)) ^ ((h2 ^ (h3*prime))*prime) ^ (unsigned long long)n;
    h &= 0x000FFFFFFFFFFFFFULL;                 // exactly representable as a scalar (the cache stores it as one)
    return scalar(h);
}

// `reuseUpper yes`: skip the upload of upper() when the array the device already holds has the same address, size and
// hash.  Off by default: a solve with stale coefficients is silent, an extra copy is not.
static bool upperUnchanged(gpuPCGCacheEntry& entry, const scalarField& upper, const bool reuse)
{
    if (!reuse)
    {
        entry.upperPtr = NULL;
        return false;
    }
    const scalar check = upperHash(upper);
    const bool same = entry.upperPtr == upper.begin() && entry.upperCheck == check;
    entry.upperPtr = upper.begin();
    entry.upperCheck = check;
    return same;
}

// Device of this rank: the node-local rank where the MPI launcher exports it (ranks are then free to be placed unevenly
// or over several nodes), else rank mod visible devices
static int defaultDevice(const int nDev)
{
    static const char* vars[] = {"OMPI_COMM_WORLD_LOCAL_RANK", "MV2_COMM_WORLD_LOCAL_RANK", "MPI_LOCALRANKID",
                                 "SLURM_LOCALID", "PMI_LOCAL_RANK", "LOCAL_RANK"};
    for (unsigned v = 0; v < sizeof(vars)/sizeof(vars[0]); v++)
    {
        const char* e = getenv(vars[v]);
        if (e && *e) return (nDev > 0) ? (atoi(e) % nDev) : 0;
    }
    return (nDev > 0) ? (Pstream::myProcNo() % nDev) : 0;
}

} // End namespace Foam


// * * * * * * * * * * * * * * * * Constructors  * * * * * * * * * * * * * * //

Foam::gpuPCG::gpuPCG
(
    const word& fieldName,
    const lduMatrix& matrix,
    const FieldField<Field, scalar>& coupleBouCoeffs,
    const FieldField<Field, scalar>& coupleIntCoeffs,
    const lduInterfaceFieldPtrsList& interfaces,
    const dictionary& dict
)
:
    lduMatrix::solver
    (
        fieldName,
        matrix,
        coupleBouCoeffs,
        coupleIntCoeffs,
        interfaces,
        dict
    ),
    preconName_("DIC"),
    reuseUpper_(false),
    device_(-1)
{
    if (dict.found("preconditioner"))
    {
        if (dict.isDict("preconditioner"))
        {
            preconName_ = word(dict.subDict("preconditioner").lookup("preconditioner"));
        }
        else
        {
            preconName_ = word(dict.lookup("preconditioner"));
        }
    }

    if (preconName_ != "DIC")
    {
        FatalErrorIn("gpuPCG::gpuPCG(...)")
            << "preconditioner " << preconName_ << " is not available on the GPU; "
            << "gpuPCG implements foam-extend's DIC (DICPreconditioner) only"
            << abort(FatalError);
    }

    word reuse("no");
    if (dict.readIfPresent("reuseUpper", reuse))
    {
        reuseUpper_ = (reuse == "yes" || reuse == "on" || reuse == "true");
    }
    dict.readIfPresent("device", device_);
}


// * * * * * * * * * * * * * * * Member Functions  * * * * * * * * * * * * * //

void Foam::gpuPCG::clearCache(const lduAddressing* addr)
{
    if (addr)
    {
        std::map<const lduAddressing*, gpuPCGCacheEntry>::iterator it = gpuPCGCache_.find(addr);
        if (it != gpuPCGCache_.end())
        {
            gpupcg_destroy(it->second.handle);
            gpuPCGCache_.erase(it);
        }
        return;
    }
    for
    (
        std::map<const lduAddressing*, gpuPCGCacheEntry>::iterator it = gpuPCGCache_.begin();
        it != gpuPCGCache_.end();
        ++it
    )
    {
        gpupcg_destroy(it->second.handle);
    }
    gpuPCGCache_.clear();
}


namespace Foam
{

// Coupled (processor) interfaces of a sub-domain, in interface order
struct gpuPCGPatches
{
    std::vector<int> ids, sizes, nbr;
    std::vector<const int*> faceCells;
};

// Device handle for an lduAddressing (created once per mesh) + its communicator in a parallel run
static gpuPCGCacheEntry& gpuPCGAcquire
(
    const lduMatrix& matrix_,
    const lduInterfaceFieldPtrsList& interfaces_,
    const label device_,
    const label nCells,
    gpuPCGPatches& P
)
{
    const lduAddressing& addr = matrix_.lduAddr();
    const label nFaces = addr.lowerAddr().size();
    std::vector<int>& patchIds = P.ids;
    std::vector<int>& patchSizes = P.sizes;
    std::vector<int>& patchNbr = P.nbr;
    std::vector<const int*>& patchFaceCells = P.faceCells;

    // --- Coupled (processor) interfaces of this sub-domain, in interface order
    forAll(interfaces_, patchI)
    {
        if (interfaces_.set(patchI))
        {
            if (!isA<processorLduInterfaceField>(interfaces_[patchI]))
            {
                FatalErrorIn("gpuPCG::solve(...)")
                    << "coupled interface " << patchI << " of type " << interfaces_[patchI].type()
                    << " is not a processor interface: only processorFvPatch halos run on the GPU"
                    << abort(FatalError);
            }
            const processorLduInterfaceField& pif =
                refCast<const processorLduInterfaceField>(interfaces_[patchI]);
            if (pif.doTransform())
            {
                // processorFvPatchField::updateInterfaceMatrix applies transformCoupleField(pnf, cmpt) on rotational
                // (processor-cyclic) patches; the device halo update does not
                FatalErrorIn("gpuPCG::solve(...)")
                    << "processor interface " << patchI << " transforms its neighbour field (doTransform()): "
                    << "not supported on the GPU" << abort(FatalError);
            }
            patchIds.push_back(patchI);
            patchSizes.push_back(addr.patchAddr(patchI).size());
            patchFaceCells.push_back(addr.patchAddr(patchI).begin());
            patchNbr.push_back(pif.neighbProcNo());
        }
    }
    const int nPatches = patchIds.size();

    // --- Device handle for this addressing (created once per mesh)
    std::map<const lduAddressing*, gpuPCGCacheEntry>::iterator it = gpuPCGCache_.find(&addr);
    if (it != gpuPCGCache_.end() && (it->second.nCells != nCells || it->second.nFaces != nFaces))
    {
        gpupcg_destroy(it->second.handle);      // same address, different mesh: rebuild
        gpuPCGCache_.erase(it);
        it = gpuPCGCache_.end();
    }
    if (it == gpuPCGCache_.end())
    {
        gpuPCGCacheEntry e;
        e.nCells = nCells;
        e.nFaces = nFaces;
        e.upperPtr = NULL;
        e.upperCheck = 0;
        e.commReady = false;
        // one decomposePar sub-domain per GPU of the box
        const int nDev = gpupcg_device_count();
        const int device = (device_ >= 0) ? device_ : defaultDevice(nDev);
        const int rc = gpupcg_create
        (
            &e.handle, device, nCells, nFaces,
            addr.lowerAddr().begin(), addr.upperAddr().begin(),
            nPatches,
            nPatches ? &patchSizes[0] : NULL,
            nPatches ? &patchFaceCells[0] : NULL,
            nPatches ? &patchNbr[0] : NULL
        );
        if (rc != GPUPCG_OK)
        {
            FatalErrorIn("gpuPCG::solve(...)")
                << "gpupcg_create failed (" << rc << "): " << gpupcg_last_error(NULL)
                << abort(FatalError);
        }
        it = gpuPCGCache_.insert(std::make_pair(&addr, e)).first;
    }
    gpuPCGCacheEntry& entry = it->second;

#ifndef FOAMLIKE
    if (Pstream::parRun() && !entry.commReady)
    {
        // ncclUniqueId from the master, broadcast through Pstream
        List<char> id(128);
        if (Pstream::master())
        {
            if (gpupcg_comm_unique_id(id.begin()) != GPUPCG_OK)
            {
                FatalErrorIn("gpuPCG::solve(...)") << "ncclGetUniqueId failed" << abort(FatalError);
            }
        }
        Pstream::scatter(id);
        if (gpupcg_comm_init(entry.handle, Pstream::myProcNo(), Pstream::nProcs(), id.begin()) != GPUPCG_OK)
        {
            FatalErrorIn("gpuPCG::solve(...)")
                << "gpupcg_comm_init failed: " << gpupcg_last_error(entry.handle) << abort(FatalError);
        }
        entry.commReady = true;
    }
#endif

    return entry;
}

} // End namespace Foam


Foam::lduMatrix::solverPerformance Foam::gpuPCG::solve
(
    scalarField& x,
    const scalarField& b,
    const direction cmpt
) const
{
    const label nCells = x.size();
    gpuPCGPatches P;
    gpuPCGCacheEntry& entry = gpuPCGAcquire(matrix_, interfaces_, device_, nCells, P);
    const std::vector<int>& patchIds = P.ids;
    const int nPatches = patchIds.size();

    // --- Matrix: diag always (addBoundaryDiag changes it per component), upper only when it changed
    const scalarField& upper = matrix_.upper();
    const bool sameUpper = upperUnchanged(entry, upper, reuseUpper_);

    std::vector<const scalar*> bouPtrs(nPatches);
    for (int p = 0; p < nPatches; p++)
    {
        bouPtrs[p] = coupleBouCoeffs_[patchIds[p]].begin();
    }

    int rc = gpupcg_set_matrix
    (
        entry.handle,
        matrix_.diag().begin(),
        sameUpper ? NULL : upper.begin(),
        NULL,
        nPatches ? &bouPtrs[0] : NULL
    );
    if (rc != GPUPCG_OK)
    {
        FatalErrorIn("gpuPCG::solve(...)")
            << "gpupcg_set_matrix failed (" << rc << "): " << gpupcg_last_error(entry.handle)
            << abort(FatalError);
    }

    // --- Solve
    gpupcg_perf perf;
    rc = gpupcg_solve
    (
        entry.handle, x.begin(), b.begin(),
        tolerance_, relTolerance_, minIter_, maxIter_,
        &perf, NULL, 0
    );
    if (rc != GPUPCG_OK)
    {
        FatalErrorIn("gpuPCG::solve(...)")
            << "gpupcg_solve failed (" << rc << "): " << gpupcg_last_error(entry.handle)
            << abort(FatalError);
    }

    // --- Setup class containing solver performance data (same fields PCG::solve fills)
    lduMatrix::solverPerformance solverPerf
    (
        preconName_ + typeName,
        fieldName_,
        perf.initialResidual,
        perf.finalResidual,
        perf.nIterations,
        perf.converged != 0,
        perf.singular != 0
    );

    return solverPerf;
}


void Foam::gpuPCG::solveVector
(
    const label nCmpt,
    scalar* x,
    const scalar* b,
    const scalar* diagCmpt,
    const std::vector<const scalar*>& coupleBouCoeffsCmpt,
    std::vector<lduMatrix::solverPerformance>& perfCmpt
) const
{
    static const char* cmptNames[3] = {"x", "y", "z"};
    const label nCells = matrix_.lduAddr().size();
    gpuPCGPatches P;
    gpuPCGCacheEntry& entry = gpuPCGAcquire(matrix_, interfaces_, device_, nCells, P);
    const int nPatches = P.ids.size();

    const scalarField& upper = matrix_.upper();
    const bool sameUpper = upperUnchanged(entry, upper, reuseUpper_);

    std::vector<const scalar*> bouPtrs(nPatches);
    for (int p = 0; p < nPatches; p++)
    {
        bouPtrs[p] = coupleBouCoeffsCmpt[P.ids[p]];
    }

    int rc = gpupcg_set_matrix_vector
    (
        entry.handle, nCmpt, diagCmpt,
        sameUpper ? NULL : upper.begin(),
        nPatches ? &bouPtrs[0] : NULL
    );
    if (rc != GPUPCG_OK)
    {
        FatalErrorIn("gpuPCG::solveVector(...)")
            << "gpupcg_set_matrix_vector failed (" << rc << "): " << gpupcg_last_error(entry.handle)
            << abort(FatalError);
    }

    gpupcg_perf perf[3];
    rc = gpupcg_solve_vector
    (
        entry.handle, nCmpt, x, b,
        tolerance_, relTolerance_, minIter_, maxIter_,
        perf, NULL, 0
    );
    if (rc != GPUPCG_OK)
    {
        FatalErrorIn("gpuPCG::solveVector(...)")
            << "gpupcg_solve_vector failed (" << rc << "): " << gpupcg_last_error(entry.handle)
            << abort(FatalError);
    }

    perfCmpt.clear();
    for (label c = 0; c < nCmpt; c++)
    {
        perfCmpt.push_back
        (
            lduMatrix::solverPerformance
            (
                preconName_ + typeName,
                fieldName_ + cmptNames[c],
                perf[c].initialResidual,
                perf[c].finalResidual,
                perf[c].nIterations,
                perf[c].converged != 0,
                perf[c].singular != 0
            )
        );
    }
}


// ************************************************************************* //
