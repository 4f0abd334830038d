/*---------------------------------------------------------------------------*\
    gpuPCG.C -- foam-extend-3.1 lduMatrix::solver adaptor over include/gpupcg.h
\*---------------------------------------------------------------------------*/

#include "gpuPCG.H"
#include "gpupcg.h"

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

// cheap fingerprint of upper(): size, address and <= 1024 strided samples
static scalar upperFingerprint(const scalarField& upper)
{
    const label n = upper.size();
    const label stride = (n > 1024) ? n/1024 : 1;
    scalar s = n;
    for (label i = 0; i < n; i += stride)
    {
        s += upper[i]*(1 + (i & 7));
    }
    return s;
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
    reuseUpper_(true),
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

    word reuse("yes");
    if (dict.readIfPresent("reuseUpper", reuse))
    {
        reuseUpper_ = (reuse == "yes" || reuse == "on" || reuse == "true");
    }
    dict.readIfPresent("device", device_);
}


// * * * * * * * * * * * * * * * Member Functions  * * * * * * * * * * * * * //

void Foam::gpuPCG::clearCache()
{
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
        // one decomposePar sub-domain per GPU of the box: rank r -> device r mod (visible devices)
        const int nDev = gpupcg_device_count();
        const int device = (device_ >= 0) ? device_ : (nDev > 0 ? Pstream::myProcNo() % nDev : 0);
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
    const scalar check = upperFingerprint(upper);
    const bool sameUpper =
        reuseUpper_ && entry.upperPtr == upper.begin() && entry.upperCheck == check;

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
    entry.upperPtr = upper.begin();
    entry.upperCheck = check;

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
    const scalar check = upperFingerprint(upper);
    const bool sameUpper =
        reuseUpper_ && entry.upperPtr == upper.begin() && entry.upperCheck == check;

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
    entry.upperPtr = upper.begin();
    entry.upperCheck = check;

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
