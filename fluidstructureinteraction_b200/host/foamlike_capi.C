// C entry points so that the Python tests can drive the C++ host side (stand-in API + adaptor + segregated solve)
// exactly the way foam-extend would: dictionary text -> lduMatrix::solver::New -> solve.
#include "segregatedSolve.H"
#include "stressModelEvolve.H"
#include "../foam/gpuPCG/gpuPCG.H"

#include <algorithm>
#include <cstring>
#include <sstream>
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

// Two scalar solves on ONE lduMatrix object (same upper() address, as when the allocator hands a fresh fvMatrix the
// storage of the previous one): between them upper[changeFace] += delta and diag is patched accordingly by the caller's
// second diag array.  With `reuseUpper yes` the adaptor must notice the changed coefficient.  psi1 / psi2: the solutions.
extern "C" int foamlike_solve_twice(int nCells, int nFaces, const int* lower, const int* upper, const double* diag1,
                                    const double* diag2, const double* upperCoeffs, int changeFace, double delta,
                                    const double* source, const char* solverDictText, double* psi1, double* psi2)
{
    try
    {
        lduMatrix::debug = 0;
        labelList l(nFaces), u(nFaces);
        for (int f = 0; f < nFaces; f++) { l[f] = lower[f]; u[f] = upper[f]; }
        lduAddressing addr(nCells, l, u, std::vector<labelList>());
        lduMatrix m(addr);
        for (int f = 0; f < nFaces; f++) m.upper()[f] = upperCoeffs[f];
        dictionary dict((std::string(solverDictText)));
        for (int pass = 0; pass < 2; pass++)
        {
            const double* diag = pass ? diag2 : diag1;
            for (int c = 0; c < nCells; c++) m.diag()[c] = diag[c];
            if (pass) m.upper()[changeFace] += delta;
            fvMatrixLike eqn("T", 1, m);
            eqn.source.assign(source, source + nCells);
            std::vector<scalar> psiV(nCells, 0.0);
            eqn.solve(psiV, dict, 0);
            memcpy(pass ? psi2 : psi1, psiV.data(), sizeof(double)*nCells);
        }
        gpuPCG::clearCache();
        return 0;
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


// One time step of the solid model (unsTotalLagrangianStress::evolve(), linear elastic, steadyState) on the GPU.
// ints:  [nCells, nFaces, nInternalFaces, nPatches, nPoints, device];  meshInts: owner, neighbour, patchStart, patchSize,
// patchType, faceStart, facePoints;  meshReals: Sf, magSf, Cf, C, V, weights, deltaCoeffs, points;
// fields: mu, lambda, rho, g[3], traction, pressure, fixedValue;  D/Dold in-out start values (D.oldTime() = Dold).
// out: D, pointD, gradD, gradDf, epsilon, sigma; result = {converged, nOuter, initialResidual, lastSolverInitialResidual,
// maxRes, res, totalPcgIterations}; history[historyCap] = relative residual of every outer iteration.
extern "C" int foamlike_evolve(const int* ints, const int* const* meshInts, const double* const* meshReals,
                               const double* const* fields, double limitCoeff, const char* stressPropertiesText,
                               const char* solverDictText, double relaxD, Foam::volToPointFn interpolate, void* ctx,
                               double* D, const double* Dold, double* pointD, double* gradD, double* gradDf,
                               double* epsilon, double* sigma, double* result, double* history, int historyCap)
{
    try
    {
        lduMatrix::debug = 0;
        solidMeshData M;
        M.nCells = ints[0]; M.nFaces = ints[1]; M.nInternalFaces = ints[2]; M.nPatches = ints[3]; M.nPoints = ints[4];
        M.owner = meshInts[0]; M.neighbour = meshInts[1]; M.patchStart = meshInts[2]; M.patchSize = meshInts[3];
        M.patchType = meshInts[4]; M.faceStart = meshInts[5]; M.facePoints = meshInts[6];
        M.Sf = meshReals[0]; M.magSf = meshReals[1]; M.Cf = meshReals[2]; M.C = meshReals[3]; M.V = meshReals[4];
        M.weights = meshReals[5]; M.deltaCoeffs = meshReals[6]; M.points = meshReals[7];
        unsTotalLagrangianStressGpu model(M, ints[5], fields[0], fields[1], fields[2], fields[3], fields[4], fields[5],
                                          fields[6], limitCoeff);
        model.D.assign(D, D + 3*(size_t)M.nCells);
        model.Dold.assign(Dold, Dold + 3*(size_t)M.nCells);
        model.gradD.assign(gradD, gradD + 9*(size_t)M.nCells);
        model.gradDf.assign(gradDf, gradDf + 9*(size_t)M.nFaces);
        dictionary sp((std::string(stressPropertiesText))), sd((std::string(solverDictText)));
        evolveResult R = model.evolve(sp, sd, relaxD, interpolate, ctx);
        memcpy(D, model.D.data(), sizeof(double)*model.D.size());
        memcpy(pointD, model.pointD.data(), sizeof(double)*model.pointD.size());
        memcpy(gradD, model.gradD.data(), sizeof(double)*model.gradD.size());
        memcpy(gradDf, model.gradDf.data(), sizeof(double)*model.gradDf.size());
        memcpy(epsilon, model.epsilon.data(), sizeof(double)*model.epsilon.size());
        memcpy(sigma, model.sigma.data(), sizeof(double)*model.sigma.size());
        result[0] = R.converged; result[1] = R.nOuterIterations; result[2] = R.initialResidual;
        result[3] = R.lastSolverInitialResidual; result[4] = R.maxRes; result[5] = R.res; result[6] = R.totalPcgIterations;
        const int nh = std::min<int>(historyCap, (int)model.residualHistory.size());
        for (int i = 0; i < nh; i++) history[i] = model.residualHistory[i];
        return nh;
    }
    catch (const Foam::error& e)
    {
        g_lastError = e.what();
        gpuPCG::clearCache();
        return -1;
    }
}


// ---- the run-time selectable GPU stress model (gpuStressModel.{H,C}) driven the way a foam-extend solver drives a
// stressModel: stressModel::New(mesh) from `stressModel ...;` of stressProperties, setTraction / setPressure, evolve() ------
#include "gpuStressModel.H"

namespace
{
struct StressCase
{
    fvMeshLike mesh;
    std::vector<std::vector<int> > ints;
    std::vector<std::vector<double> > reals;
    std::vector<unsigned char> mirror;
    stressModel* model;
    StressCase() : model(0) {}
    ~StressCase() { delete model; }
};

std::vector<word> splitWords(const char* s)
{
    std::vector<word> out; std::istringstream is(s); word w;
    while (is >> w) out.push_back(w);
    return out;
}
}

// ints: nCells, nFaces, nInternalFaces, nPatches, nPoints, device, nPointConstraints, solutionD[3]
// meshInts (with lengths in intLens): owner, neighbour, patchStart, patchSize, faceStart, facePoints, lsqStart, lsqEntry, pcPoint, pcStart, pcKind
// meshReals (realLens): Sf, magSf, Cf, C, V, weights, deltaCoeffs, points, traction, pressure, fixedValue, lsqInv, lsqOrigin, pcVec
extern "C" void* foamlike_stress_new(const int* ints, const int* const* meshInts, const long long* intLens,
                                     const double* const* meshReals, const long long* realLens, const unsigned char* lsqMirror,
                                     const char* patchNames, const char* patchFieldTypes, const char* stressProperties,
                                     const char* rheologyProperties, const char* fvSolution, const char* fvSchemes, double deltaT)
{
    StressCase* S = new StressCase();
    try
    {
        S->ints.resize(11); S->reals.resize(14);
        for (int k = 0; k < 11; k++) S->ints[k].assign(meshInts[k], meshInts[k] + intLens[k]);
        for (int k = 0; k < 14; k++) S->reals[k].assign(meshReals[k], meshReals[k] + realLens[k]);
        fvMeshLike& M = S->mesh;
        M.nCells = ints[0]; M.nFaces = ints[1]; M.nInternalFaces = ints[2]; M.nPatches = ints[3]; M.nPoints = ints[4]; M.device = ints[5];
        M.nPointConstraints = ints[6];
        for (int d = 0; d < 3; d++) M.solutionD[d] = ints[7 + d];
        S->mirror.assign(lsqMirror, lsqMirror + M.nPoints);
        M.owner = S->ints[0].data(); M.neighbour = S->ints[1].data(); M.patchStart = S->ints[2].data(); M.patchSize = S->ints[3].data();
        M.faceStart = S->ints[4].data(); M.facePoints = S->ints[5].data(); M.lsqStart = S->ints[6].data(); M.lsqEntry = S->ints[7].data();
        M.pcPoint = S->ints[8].data(); M.pcStart = S->ints[9].data(); M.pcKind = S->ints[10].data();
        M.Sf = S->reals[0].data(); M.magSf = S->reals[1].data(); M.Cf = S->reals[2].data(); M.C = S->reals[3].data(); M.V = S->reals[4].data();
        M.weights = S->reals[5].data(); M.deltaCoeffs = S->reals[6].data(); M.points = S->reals[7].data();
        M.traction = S->reals[8].data(); M.pressure = S->reals[9].data(); M.fixedValue = S->reals[10].data();
        M.lsqInv = S->reals[11].data(); M.lsqOrigin = S->reals[12].data(); M.pcVec = S->reals[13].data();
        M.lsqMirror = S->mirror.data();
        M.patchNames = splitWords(patchNames); M.patchFieldTypes = splitWords(patchFieldTypes);
        M.stressProperties = dictionary(std::string(stressProperties));
        M.rheologyProperties = dictionary(std::string(rheologyProperties));
        M.fvSolution = dictionary(std::string(fvSolution));
        M.fvSchemes = dictionary(std::string(fvSchemes));
        M.deltaT = deltaT;
        S->model = stressModel::New(M);
        return S;
    }
    catch (const Foam::error& e)
    {
        g_lastError = e.what();
        delete S;
        return 0;
    }
}

extern "C" void foamlike_stress_delete(void* p) { delete (StressCase*)p; }

extern "C" int foamlike_stress_set(void* p, int what /*0 traction, 1 pressure*/, int patchID, const double* values)
{
    try
    {
        StressCase* S = (StressCase*)p;
        if (what == 0) S->model->setTraction(patchID, values); else S->model->setPressure(patchID, values);
        return 0;
    }
    catch (const Foam::error& e) { g_lastError = e.what(); return -1; }
}

extern "C" int foamlike_stress_new_time_step(void* p)
{
    try { ((StressCase*)p)->model->newTimeStep(); return 0; }
    catch (const Foam::error& e) { g_lastError = e.what(); return -1; }
}

// result[9]: converged, nOuter, totalPcg, enforceLinear, initialResidual, lastSolverInitialResidual, maxRes, res, nValidCmpts;
// returns the number of outer iterations recorded in hist / pcg (pcg: 3 ints per outer iteration, the first nValidCmpts in use)
extern "C" int foamlike_stress_evolve(void* p, double* result, double* hist, int* pcg, int cap)
{
    try
    {
        StressCase* S = (StressCase*)p;
        const bool ok = S->model->evolve();
        const gpuEvolveInfo& I = static_cast<gpuUnsTotalLagrangianStress*>(S->model)->info();
        result[0] = ok; result[1] = I.nOuterIterations; result[2] = I.totalPcgIterations; result[3] = I.enforceLinear;
        result[4] = I.initialResidual; result[5] = I.lastSolverInitialResidual; result[6] = I.maxRes; result[7] = I.res;
        result[8] = I.nValidCmpts;
        const int n = std::min<int>(cap, (int)I.residualHistory.size());
        for (int i = 0; i < n; i++) hist[i] = I.residualHistory[i];
        for (int i = 0; i < 3*n; i++) pcg[i] = I.pcgIterations[i];
        return n;
    }
    catch (const Foam::error& e) { g_lastError = e.what(); return -1; }
}

extern "C" long long foamlike_stress_field(void* p, const char* name, double* out, long long cap)
{
    try
    {
        StressCase* S = (StressCase*)p;
        const long long n = S->model->fieldSize(name);
        if (out && cap >= n) S->model->field(name, out);
        return n;
    }
    catch (const Foam::error& e) { g_lastError = e.what(); return -1; }
}
