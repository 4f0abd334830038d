#include "stressModelEvolve.H"
#include "../foam/gpuPCG/gpuPCG.H"

#include <algorithm>
#include <chrono>
#include <cstdio>

#include <cstdlib>
#include <sstream>

namespace Foam
{

static void meshCheck(gpupcg_mesh m, int rc, const char* what)
{
    if (rc != 0)
    {
        std::ostringstream os;
        os << what << " failed (" << rc << "): " << gpupcg_mesh_last_error(m);
        FatalErrorIn("unsTotalLagrangianStressGpu") << os.str() << abort(FatalError);
    }
}

unsTotalLagrangianStressGpu::unsTotalLagrangianStressGpu
(
    const solidMeshData& mesh, int device, const scalar* mu, const scalar* lambda, const scalar* rho, const scalar* g,
    const scalar* traction, const scalar* pressure, const scalar* fixedValue, scalar limitCoeff
)
:
    mesh_(mesh), gm_(0), solver_(0), device_(device), mu_(mu, mu + mesh.nCells), lambda_(lambda, lambda + mesh.nCells), rho_(rho, rho + mesh.nCells),
    limitCoeff_(limitCoeff)
{
    const int nBF = mesh.nFaces - mesh.nInternalFaces;
    traction_.assign(traction, traction + 3*(size_t)nBF);
    pressure_.assign(pressure, pressure + nBF);
    fixedValue_.assign(fixedValue, fixedValue + 3*(size_t)nBF);
    g_[0] = g[0]; g_[1] = g[1]; g_[2] = g[2];
    int rc = gpupcg_mesh_create(&gm_, device, mesh.nCells, mesh.nFaces, mesh.nInternalFaces, mesh.owner, mesh.neighbour,
                                mesh.nPatches, mesh.patchStart, mesh.patchSize, mesh.patchType, mesh.Sf, mesh.magSf, mesh.Cf,
                                mesh.C, mesh.V, mesh.weights, mesh.deltaCoeffs);
    meshCheck(gm_, rc, "gpupcg_mesh_create");
    meshCheck(gm_, gpupcg_mesh_set_points(gm_, mesh.nPoints, mesh.points, mesh.faceStart, mesh.facePoints), "gpupcg_mesh_set_points");
    D.assign(3*(size_t)mesh.nCells, 0.0);
    Dold = D;
    pointD.assign(3*(size_t)mesh.nPoints, 0.0);
    gradD.assign(9*(size_t)mesh.nCells, 0.0);
    gradDf.assign(9*(size_t)mesh.nFaces, 0.0);
    epsilon.assign(6*(size_t)mesh.nCells, 0.0);
    sigma = epsilon;
}

unsTotalLagrangianStressGpu::~unsTotalLagrangianStressGpu()
{
    if (solver_) gpupcg_destroy(solver_);
    if (gm_) gpupcg_mesh_destroy(gm_);
}

evolveResult unsTotalLagrangianStressGpu::evolve
(
    const dictionary& stressProperties, const dictionary& solverDict, scalar relaxD, volToPointFn interpolate, void* ctx
)
{
    // This driver is the round-1, half-resident form of the loop (linear elastic, steadyState d2dt2, vol->point interpolation as
    // a host callback).  Settings it does not implement are refused, not ignored; the complete loop -- every branch, device
    // resident -- is gpuUnsTotalLagrangianStress (gpuStressModel.{H,C}).
    if (stressProperties.found("nonLinear"))
    {
        const std::string v = stressProperties.lookup("nonLinear");
        if (v == "yes" || v == "on" || v == "true")
            FatalErrorIn("unsTotalLagrangianStressGpu::evolve") << "nonLinear yes: use gpuUnsTotalLagrangianStress (the resident loop); "
                << "this driver solves the linear model only" << abort(FatalError);
    }
    if (stressProperties.found("K") || stressProperties.found("thermalStress"))
        FatalErrorIn("unsTotalLagrangianStressGpu::evolve") << "damping (K) / thermal stress: use gpuUnsTotalLagrangianStress"
            << abort(FatalError);
    const int nCorr = atoi(std::string(stressProperties.lookup("nCorrectors")).c_str());
    const scalar convergenceTolerance = atof(std::string(stressProperties.lookup("convergenceTolerance")).c_str());
    scalar relConvergenceTolerance = 0;
    if (stressProperties.found("relConvergenceTolerance"))
        relConvergenceTolerance = atof(std::string(stressProperties.lookup("relConvergenceTolerance")).c_str());

    const int nC = mesh_.nCells, nIF = mesh_.nInternalFaces, nBF = mesh_.nFaces - nIF;
    labelList l(nIF), u(nIF);
    for (int f = 0; f < nIF; f++) { l[f] = mesh_.owner[f]; u[f] = mesh_.neighbour[f]; }
    lduAddressing addr(nC, l, u, std::vector<labelList>());
    lduMatrix matrix(addr);
    std::vector<scalar> upper(nIF), diag(nC), source(3*(size_t)nC), ic(3*(size_t)std::max(1, nBF)), bc(3*(size_t)std::max(1, nBF));
    std::vector<scalar> Dprev(D.size());

    // `solver gpuPCG` with batchComponents (default): DEqn.solve() runs on the matrix where the assembly left it, on the
    // device (keyword residentMatrix no: the fvMatrix is rebuilt on the host and solved through lduMatrix::solver::New)
    const bool residentSolve = word(solverDict.lookup("solver")) == gpuPCG::typeName
                            && solverDict.lookupOrDefault<word>("batchComponents", "yes") != "no"
                            && solverDict.lookupOrDefault<word>("residentMatrix", "yes") != "no";
    if (residentSolve && solverDict.found("preconditioner") && word(solverDict.lookup("preconditioner")) != "DIC")
        FatalErrorIn("unsTotalLagrangianStressGpu::evolve") << "gpuPCG implements the DIC preconditioner only" << abort(FatalError);
    scalar tolerance = 1e-6, relTol = 0;
    int minIter = 0, maxIter = 1000;                                    // lduMatrix::solver::readControls defaults
    solverDict.readIfPresent("tolerance", tolerance); solverDict.readIfPresent("relTol", relTol);
    solverDict.readIfPresent("minIter", minIter); solverDict.readIfPresent("maxIter", maxIter);

    evolveResult R;
    R.converged = 1; R.nOuterIterations = 0; R.initialResidual = 0; R.lastSolverInitialResidual = 0;
    R.maxRes = 0; R.res = 1; R.totalPcgIterations = 0;
    residualHistory.clear();

    // GPUPCG_EVOLVE_TIMING=1: wall time per phase of the outer loop on stderr
    const bool timing = getenv("GPUPCG_EVOLVE_TIMING") != 0;
    typedef std::chrono::steady_clock clk;
    double tAsm = 0, tSolve = 0, tRelax = 0, tInterp = 0, tGrad = 0;
    clk::time_point t0 = clk::now();
    auto lap = [&](double& acc) { clk::time_point t1 = clk::now(); acc += std::chrono::duration<double>(t1 - t0).count(); t0 = t1; };

    int iCorr = 0;
    scalar res = 1, maxRes = 0, curConvergenceTolerance = convergenceTolerance;
    do
    {
        Dprev = D;                                                      // D_.storePrevIter()

        // the gradients cross PCIe once per time step: uploaded with the first assembly, then resident on the device
        // (gpupcg_gradients leaves them there), fetched once after the loop
        const scalar* gD = (iCorr == 0) ? &gradD[0] : 0;
        const scalar* gDf = (iCorr == 0) ? &gradDf[0] : 0;
        lduMatrix::solverPerformance solverPerf;
        if (residentSolve)
        {
            // the assembled equation stays on the device: addBoundaryDiag / addBoundarySource and the three component
            // solves happen there (gpupcg_solve_assembled_D); only D travels
            meshCheck(gm_, gpupcg_assemble_D(gm_, &mu_[0], &lambda_[0], &rho_[0], g_, &D[0], gD, gDf, &traction_[0], &pressure_[0],
                                             &fixedValue_[0], limitCoeff_, 1, 0, 0, 0, 0, 0, 0), "gpupcg_assemble_D");
            lap(tAsm);
            if (!solver_)
            {
                int rc = gpupcg_create(&solver_, device_, nC, nIF, mesh_.owner, mesh_.neighbour, 0, 0, 0, 0);
                if (rc != 0)
                    FatalErrorIn("unsTotalLagrangianStressGpu::evolve") << "gpupcg_create failed (" << rc << "): "
                                                                       << gpupcg_last_error(0) << abort(FatalError);
            }
            gpupcg_perf perf[3];
            int rc = gpupcg_solve_assembled_D(solver_, gm_, &D[0], tolerance, relTol, minIter, maxIter, perf);
            if (rc != 0)
                FatalErrorIn("unsTotalLagrangianStressGpu::evolve") << "gpupcg_solve_assembled_D failed (" << rc << "): "
                                                                   << gpupcg_last_error(solver_) << abort(FatalError);
            static const char* cmptNames[3] = {"x", "y", "z"};
            for (int k = 0; k < 3; k++)
            {
                lduMatrix::solverPerformance p("DICgpuPCG", word("D") + cmptNames[k], perf[k].initialResidual, perf[k].finalResidual,
                                               perf[k].nIterations, perf[k].converged != 0, perf[k].singular != 0);
                p.print();
                R.totalPcgIterations += p.nIterations();
                if (k == 0 || (p.initialResidual() > solverPerf.initialResidual() && !p.singular())) solverPerf = p;
            }
        }
        else
        {
        meshCheck(gm_, gpupcg_assemble_D(gm_, &mu_[0], &lambda_[0], &rho_[0], g_, &D[0], gD, gDf, &traction_[0],
                                         &pressure_[0], &fixedValue_[0], limitCoeff_, 1, &upper[0], &diag[0], &source[0],
                                         &ic[0], &bc[0], 0), "gpupcg_assemble_D");
        lap(tAsm);
        for (int c = 0; c < nC; c++) matrix.diag()[c] = diag[c];
        for (int f = 0; f < nIF; f++) matrix.upper()[f] = upper[f];
        fvMatrixLike DEqn("D", 3, matrix);
        DEqn.source = source;
        for (int p = 0; p < mesh_.nPatches; p++)
        {
            fvPatchCoeffs pc;
            const int n = mesh_.patchSize[p], b0 = mesh_.patchStart[p] - nIF;
            pc.faceCells.setSize(n);
            for (int i = 0; i < n; i++) pc.faceCells[i] = mesh_.owner[mesh_.patchStart[p] + i];
            pc.internalCoeffs.assign(ic.begin() + 3*(size_t)b0, ic.begin() + 3*(size_t)(b0 + n));
            pc.boundaryCoeffs.assign(bc.begin() + 3*(size_t)b0, bc.begin() + 3*(size_t)(b0 + n));
            DEqn.patches.push_back(pc);
        }
        std::vector<lduMatrix::solverPerformance> per;
        solverPerf = DEqn.solve(D, solverDict, &per);                                   // DEqn.solve()
        for (size_t k = 0; k < per.size(); k++) R.totalPcgIterations += per[k].nIterations();
        }

        lap(tSolve);
        if (iCorr == 0) R.initialResidual = solverPerf.initialResidual();
        R.lastSolverInitialResidual = solverPerf.initialResidual();

        // D_.relax();  ...  res = residual()   (the residual only reads D, prevIter and oldTime: evaluated with the relax)
        // relaxD <= 0: no relaxation factor for D in fvSolution -> D.relax() is a no-op (a factor of 0 would freeze D at prevIter)
        meshCheck(gm_, gpupcg_relax_residual(gm_, relaxD, relaxD > 0 ? 1 : 0, &D[0], &Dprev[0], &Dold[0], &res),
                  "gpupcg_relax_residual");
        lap(tRelax);

        interpolate(ctx, &D[0], &pointD[0]);                            // volToPoint_.interpolate(D_, pointD_)
        lap(tInterp);
        // gradDold = NULL: the device's current grad(D) is the registered gradient; patchGradient = NULL: the
        // tractionDisplacement gradient() the assembly above left on the device; outputs stay on the device
        meshCheck(gm_, gpupcg_gradients(gm_, &D[0], &pointD[0], 0, &fixedValue_[0], 0, limitCoeff_, 1, 0, 0, 0, 0), "gpupcg_gradients");
        lap(tGrad);

        if (res > maxRes) maxRes = res;
        curConvergenceTolerance = maxRes*relConvergenceTolerance;
        if (curConvergenceTolerance < convergenceTolerance) curConvergenceTolerance = convergenceTolerance;
        residualHistory.push_back(res);
    }
    while (res > curConvergenceTolerance && ++iCorr < nCorr);

    meshCheck(gm_, gpupcg_sigma(gm_, &mu_[0], &lambda_[0], 0, &epsilon[0], &sigma[0]), "gpupcg_sigma");
    meshCheck(gm_, gpupcg_mesh_fetch(gm_, 1, &gradD[0]), "gpupcg_mesh_fetch(gradD)");
    meshCheck(gm_, gpupcg_mesh_fetch(gm_, 2, &gradDf[0]), "gpupcg_mesh_fetch(gradDf)");
    double tEnd = 0; lap(tEnd);
    if (timing)
        fprintf(stderr, "evolve timing [s]: assemble %.3f  solve %.3f  relax/residual %.3f  volToPoint %.3f  gradients %.3f  sigma+fetch %.3f  (%s)\n",
                tAsm, tSolve, tRelax, tInterp, tGrad, tEnd, residentSolve ? "resident" : "host matrix");
    gpuPCG::clearCache(&addr);  // `addr` dies with this call: only its handle, other meshes keep theirs
    R.nOuterIterations = iCorr; R.maxRes = maxRes; R.res = res;
    return R;
}

} // namespace Foam
