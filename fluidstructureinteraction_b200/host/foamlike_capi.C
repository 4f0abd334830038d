// C entry points so that the Python tests can drive the C++ host side (stand-in API + adaptor + segregated solve)
// exactly the way foam-extend would: dictionary text -> lduMatrix::solver::New -> solve.
#include "segregatedSolve.H"
#include "../foam/gpuPCG/gpuPCG.H"

#include <cstring>
#include <string>

using namespace Foam;

static std::string g_lastError;

struct foamlike_perf { double initialResidual, finalResidual; int nIterations, converged, singular, pad; char solverName[32]; char fieldName[32]; };

static void fill(foamlike_perf& o, const lduMatrix::solverPerformance& p)
{
    memset(&o, 0, sizeof(o));
    o.initialResidual = p.initialResidual(); o.finalResidual = p.finalResidual(); o.nIterations = p.nIterations();
    o.converged = p.converged(); o.singular = p.singular();
    strncpy(o.solverName, p.solverName().c_str(), 31);
    strncpy(o.fieldName, p.fieldName().c_str(), 31);
}

extern "C" const char* foamlike_last_error() { return g_lastError.c_str(); }

// One segregated solve of an assembled matrix.  patches: non-coupled, concatenated (patchStart[nPatches+1]).
// Returns 0, or -1 with the FatalError text in foamlike_last_error().
extern "C" int foamlike_solve(const char* psiName, int nCmpt, int nCells, int nFaces, const int* lower, const int* upper,
                              const double* diag, const double* upperCoeffs, const double* source,
                              int nPatches, const int* patchStart, const int* patchFaceCells,
                              const double* internalCoeffs, const double* boundaryCoeffs,
                              const char* solverDictText, double* psi, foamlike_perf* perCmpt, foamlike_perf* worst)
{
    try
    {
        lduMatrix::debug = 0;
        labelList l(nFaces), u(nFaces);
        for (int f = 0; f < nFaces; f++) { l[f] = lower[f]; u[f] = upper[f]; }
        lduAddressing addr(nCells, l, u, std::vector<labelList>());
        lduMatrix m(addr);
        for (int c = 0; c < nCells; c++) m.diag()[c] = diag[c];
        if (upperCoeffs) for (int f = 0; f < nFaces; f++) m.upper()[f] = upperCoeffs[f];
        fvMatrixLike eqn(psiName, nCmpt, m);
        eqn.source.assign(source, source + (size_t)nCells*nCmpt);
        for (int p = 0; p < nPatches; p++)
        {
            fvPatchCoeffs pc;
            const int n = patchStart[p + 1] - patchStart[p];
            pc.faceCells.setSize(n);
            for (int i = 0; i < n; i++) pc.faceCells[i] = patchFaceCells[patchStart[p] + i];
            pc.internalCoeffs.assign(internalCoeffs + (size_t)patchStart[p]*nCmpt, internalCoeffs + (size_t)patchStart[p + 1]*nCmpt);
            pc.boundaryCoeffs.assign(boundaryCoeffs + (size_t)patchStart[p]*nCmpt, boundaryCoeffs + (size_t)patchStart[p + 1]*nCmpt);
            eqn.patches.push_back(pc);
        }
        dictionary dict((std::string(solverDictText)));
        std::vector<scalar> psiV(psi, psi + (size_t)nCells*nCmpt);
        std::vector<lduMatrix::solverPerformance> per;
        lduMatrix::solverPerformance w = eqn.solve(psiV, dict, &per);
        memcpy(psi, psiV.data(), sizeof(double)*psiV.size());
        for (size_t k = 0; k < per.size() && perCmpt; k++) fill(perCmpt[k], per[k]);
        if (worst) fill(*worst, w);
        gpuPCG::clearCache();       // the lduAddressing above dies with this call
        return (int)per.size();
    }
    catch (const Foam::error& e)
    {
        g_lastError = e.what();
        gpuPCG::clearCache();
        return -1;
    }
}

// dictionary parser probe for the tests: value of `key` ("" if missing), sub-dictionaries with '/'.
extern "C" int foamlike_dict_lookup(const char* text, const char* key, char* out, int outLen)
{
    try
    {
        dictionary d((std::string(text)));
        const dictionary* cur = &d;
        std::string k(key);
        size_t p;
        while ((p = k.find('/')) != std::string::npos) { cur = &cur->subDict(k.substr(0, p)); k = k.substr(p + 1); }
        std::string v = cur->lookup(k);
        strncpy(out, v.c_str(), outLen - 1); out[outLen - 1] = 0;
        return 0;
    }
    catch (const Foam::error& e) { g_lastError = e.what(); return -1; }
}
