"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the outer loop of the reference's solid model,
unsTotalLagrangianStress::evolve() (stressModels/unsTotalLagrangianStress/unsTotalLagrangianStress.C:769-1009; linear elastic,
steadyState d2dt2), chaining the oracle's own pieces: orc_assemble_D -> fvMatrix<vector>::solve ([EXT] fvMatrixSolve.C:
addBoundarySource, per component addBoundaryDiag + PCG/DIC) -> D.relax() -> vol-to-point interpolation (callback) ->
fvc::grad / fvc::fGrad -> residual() (unsTotalLagrangianStressSolve.C:647-672) -> convergence logic (:949-971) -> sigma.
PARITY UNPINNED (see oracle/ldu_oracle.c): restated from the reference's source, no foam-extend run to compare with."""
import numpy as np

from . import binding as orc


def segregated_solve(l, u, diag, upper, source, patches, D, tolerance, relTol, maxIter=1000):
    """[EXT] fvMatrix<vector>::solve: patches = [(faceCells, internalCoeffs (n,3), boundaryCoeffs (n,3))].  D in place."""
    total = np.array(source, dtype=np.float64, copy=True)
    for fc, ic, bc in patches:
        for k in range(3):
            np.add.at(total[:, k], fc, bc[:, k])
    perfs = []
    for k in range(3):
        dk = np.array(diag, dtype=np.float64, copy=True)
        for fc, ic, bc in patches:
            np.add.at(dk, fc, ic[:, k])
        x = np.ascontiguousarray(D[:, k])
        perf, hist = orc.pcg_solve(l, u, dk, upper, x, np.ascontiguousarray(total[:, k]), tolerance, relTol, 0, maxIter)
        D[:, k] = x
        perfs.append(perf)
    worst = perfs[0]
    for p in perfs[1:]:
        if p["initialResidual"] > worst["initialResidual"] and not p["singular"]:
            worst = p
    return worst, perfs


def evolve(mesh, geom, patchTypes, mu, lam, rho, g, traction, pressure, fixedValue, volToPoint, nCorrectors,
           convergenceTolerance, relConvergenceTolerance=0.0, tolerance=1e-9, relTol=0.01, D=None, Dold=None, gradD=None,
           gradDf=None, limitCoeff=1.0, relaxD=None):
    nC, nF, nIF = mesh.nCells, mesh.nFaces, mesh.nInternalFaces
    D = np.zeros((nC, 3)) if D is None else np.array(D, dtype=np.float64, copy=True)
    Dold = np.zeros((nC, 3)) if Dold is None else np.asarray(Dold, dtype=np.float64)
    gradD = np.zeros((nC, 9)) if gradD is None else np.array(gradD, dtype=np.float64, copy=True)
    gradDf = np.zeros((nF, 9)) if gradDf is None else np.array(gradDf, dtype=np.float64, copy=True)
    l = np.ascontiguousarray(mesh.owner[:nIF], dtype=np.int32); u = np.ascontiguousarray(mesh.neighbour, dtype=np.int32)
    iCorr, res, maxRes, cur = 0, 1.0, 0.0, convergenceTolerance
    hist, initialResidual, lastInit, totalIts = [], 0.0, 0.0, 0
    pointD = None
    while True:
        Dprev = D.copy()
        a = orc.assemble_D(mesh, geom, patchTypes, mu, lam, rho, g, D, gradD, gradDf, traction, pressure, fixedValue, limitCoeff)
        patches = []
        for p in mesh.patches:
            b0 = p["start"] - nIF
            patches.append((np.ascontiguousarray(mesh.owner[p["start"]:p["start"] + p["size"]], dtype=np.int64),
                            a["internalCoeffs"][b0:b0 + p["size"]], a["boundaryCoeffs"][b0:b0 + p["size"]]))
        worst, perfs = segregated_solve(l, u, a["diag"], a["upper"], a["source"], patches, D, tolerance, relTol)
        totalIts += sum(p["nIterations"] for p in perfs)
        if iCorr == 0:
            initialResidual = worst["initialResidual"]
        lastInit = worst["initialResidual"]
        res = orc.relax_residual(D, Dprev, Dold, relaxD)
        pointD = np.ascontiguousarray(volToPoint(D))
        gradOld = gradD
        gradD, gradDb = orc.grad(mesh, geom, patchTypes, D, pointD, gradOld, fixedValue, a["patchGradient"])
        gradDf = orc.fgrad(mesh, geom, patchTypes, D, pointD, gradD, fixedValue, a["patchGradient"], limitCoeff)
        maxRes = max(maxRes, res)
        cur = max(maxRes*relConvergenceTolerance, convergenceTolerance)
        hist.append(res)
        if not (res > cur):
            break
        iCorr += 1
        if not (iCorr < nCorrectors):
            break
    eps, sig = orc.sigma(mu, lam, gradD)
    return dict(D=D, pointD=pointD, gradD=gradD, gradDf=gradDf, epsilon=eps, sigma=sig, nOuterIterations=iCorr,
                initialResidual=initialResidual, lastSolverInitialResidual=lastInit, maxRes=maxRes, res=res,
                totalPcgIterations=totalIts, residualHistory=np.array(hist))
