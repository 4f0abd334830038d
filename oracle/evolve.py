"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the outer loop of the reference's solid model,
unsTotalLagrangianStress::evolve() (stressModels/unsTotalLagrangianStress/unsTotalLagrangianStress.C:769-1009; linear elastic,
steadyState d2dt2), chaining the oracle's own pieces: orc_assemble_D -> fvMatrix<vector>::solve ([EXT] fvMatrixSolve.C:
addBoundarySource, per component addBoundaryDiag + PCG/DIC) -> D.relax() -> vol-to-point interpolation (callback) ->
fvc::grad / fvc::fGrad -> residual() (unsTotalLagrangianStressSolve.C:647-672) -> convergence logic (:949-971) -> sigma.
PARITY UNPINNED (see oracle/ldu_oracle.c): restated from the reference's source, no foam-extend run to compare with."""
import numpy as np

from . import binding as orc


GREAT = 1.0e15


class OldTimes:
    """The old-time levels of D as [EXT] GeometricField keeps them -- oldTime() creates a level on first use (a copy of the
    level above, carrying its timeIndex), storeOldTimes() shifts them on the first non-const access of a new time step
    (deepest first; every level takes the values and the timeIndex of the one above), nOldTimes() counts them -- because the
    `backward` schemes decide from exactly this state whether a use sees the real deltaT0 or GREAT (first-order start-up):
      backwardD2dt2Scheme::deltaT0_(vf)   backwardD2dt2Scheme.C:58-76: GREAT while the 2nd and 3rd old level carry the same timeIndex
      [EXT] backwardDdtScheme::deltaT0_(vf): GREAT while vf.nOldTimes() < 2
    (oracle/ASSUMPTIONS.md #13: the order in which evolve() touches D and creates the levels)."""

    def __init__(self):
        self.lev = []           # [values, timeIndex] of D.oldTime(), D.oldTime().oldTime(), ...
        self.ti = 0             # D.timeIndex()
        self.now = 0            # time().timeIndex()

    def advance(self):
        self.now += 1

    def touch(self, D):
        """A non-const access of D (the fvMatrix constructor's psi.boundaryField().updateCoeffs()): storeOldTimes()."""
        if self.lev and self.ti != self.now:
            for j in range(len(self.lev) - 1, 0, -1):
                self.lev[j] = [self.lev[j - 1][0], self.lev[j - 1][1]]
            self.lev[0] = [D.copy(), self.ti]
        self.ti = self.now

    def ensure(self, k, D):
        while len(self.lev) < k:
            src, ti = (D, self.ti) if not self.lev else self.lev[-1]
            self.lev.append([src.copy(), ti])

    def n_old(self, j=0):
        """nOldTimes() of the j-th old level (j = 0: D itself)."""
        return max(len(self.lev) - j, 0)

    def level(self, j, like):
        return self.lev[j][0] if j < len(self.lev) else np.zeros_like(like)


def backward_controls(old, D, d2dt2, ddt, damping, deltaT0):
    """The effective old time step of every use of a `backward` scheme in one assembly of the D equation, in the order the
    reference evaluates them (unsTotalLagrangianStress.C:846-865), creating the old levels the calls create."""
    rule = lambda n: GREAT if n < 2 else deltaT0
    e = dict(eD2=deltaT0, eDdtD=deltaT0, eDdtD0=deltaT0, eDdtD00=deltaT0, eDamp=deltaT0)
    old.touch(D)
    if d2dt2 == "backward":
        old.ensure(3, D)                                        # deltaT0_(vf) looks at vf.oldTime().oldTime().oldTime()
        e["eD2"] = GREAT if old.lev[1][1] == old.lev[2][1] else deltaT0
        e["eDdtD"] = rule(old.n_old(0))                         # backwardDdtScheme.fvmDdt(vf)
        e["eDdtD0"] = rule(old.n_old(1))                        # fvcDdt(vf.oldTime())
        e["eDdtD00"] = rule(old.n_old(2))                       # fvcDdt(vf.oldTime().oldTime()), which then creates the 4th level
        old.ensure(4, D)
    elif d2dt2 == "Euler":
        old.ensure(2, D)
    if damping:
        if ddt == "backward":
            e["eDamp"] = rule(old.n_old(0))
            old.ensure(2, D)
        elif ddt == "Euler":
            old.ensure(1, D)
    return e


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


# ---------------------------------------------------------------------------------------------------------------------
# Round 2: the complete outer loop -- every branch of evolve() (:769-1009), D.boundaryField() tracked, the reference's
# least-squares vol->point interpolation (:924) instead of a callback, 2-D (empty) cases, time stepping.
# ---------------------------------------------------------------------------------------------------------------------
def segregated_solve_ex(l, u, diag, upper, source, patches, D, tolerance, relTol, validCmpts=(1, 1, 1), maxIter=1000, minIter=0):
    """[EXT] fvMatrix<vector>::solve: components with solutionD == -1 (the empty direction) are skipped."""
    total = np.array(source, dtype=np.float64, copy=True)
    for fc, ic, bc in patches:
        for k in range(3):
            np.add.at(total[:, k], fc, bc[:, k])
    perfs, hists = {}, {}
    for k in range(3):
        if not validCmpts[k]:
            continue
        dk = np.array(diag, dtype=np.float64, copy=True)
        for fc, ic, bc in patches:
            np.add.at(dk, fc, ic[:, k])
        x = np.ascontiguousarray(D[:, k])
        perf, hist = orc.pcg_solve(l, u, dk, upper, x, np.ascontiguousarray(total[:, k]), tolerance, relTol, minIter, maxIter)
        D[:, k] = x
        perfs[k] = perf; hists[k] = hist
    ks = sorted(perfs)
    worst = perfs[ks[0]]
    for k in ks[1:]:
        if perfs[k]["initialResidual"] > worst["initialResidual"] and not perfs[k]["singular"]:
            worst = perfs[k]
    return worst, perfs, hists


class SolidCase(object):
    """State of one unsTotalLagrangianStress object: fields that persist across evolve() calls and time steps."""

    def __init__(self, mesh, geom, patchTypes, mu, lam, rho, traction=None, pressure=None, fixedValue=None, g=(0.0, 0.0, 0.0),
                 threeK=None, alpha=None, validCmpts=(1, 1, 1), pointPatchTypes=None, pointPatchValues=None):
        nC, nF, nIF = mesh.nCells, mesh.nFaces, mesh.nInternalFaces
        nBF = nF - nIF
        self.mesh, self.geom, self.patchTypes = mesh, geom, patchTypes
        # point patch fields of pointD (0/pointD): applied at the end of every interpolate() (:301 of ...Interpolate.C)
        self.pointConstraints = orc.point_constraints_build(mesh, geom, pointPatchTypes, pointPatchValues) if pointPatchTypes else None
        self.mu = np.full(nC, mu) if np.isscalar(mu) else np.asarray(mu, dtype=np.float64)
        self.lam = np.full(nC, lam) if np.isscalar(lam) else np.asarray(lam, dtype=np.float64)
        self.rho = np.full(nC, rho) if np.isscalar(rho) else np.asarray(rho, dtype=np.float64)
        self.g = np.asarray(g, dtype=np.float64)
        self.traction = np.zeros((nBF, 3)) if traction is None else np.asarray(traction, dtype=np.float64)
        self.pressure = np.zeros(nBF) if pressure is None else np.asarray(pressure, dtype=np.float64)
        self.fixedValue = np.zeros((nBF, 3)) if fixedValue is None else np.asarray(fixedValue, dtype=np.float64)
        self.threeK = None if threeK is None else (np.full(nC, threeK) if np.isscalar(threeK) else np.asarray(threeK, dtype=np.float64))
        self.alpha = None if alpha is None else (np.full(nC, alpha) if np.isscalar(alpha) else np.asarray(alpha, dtype=np.float64))
        self.DT = np.zeros(nC); self.DTb = np.zeros(nBF)
        self.validCmpts = tuple(validCmpts)
        self.D = np.zeros((nC, 3)); self.Dold = np.zeros((nC, 3)); self.Doldold = np.zeros((nC, 3))
        self.old = None             # OldTimes, once a `backward` scheme is used
        self.gradD = np.zeros((nC, 9)); self.gradDb = np.zeros((nBF, 9)); self.gradDf = np.zeros((nF, 9))
        self.pointD = np.zeros((mesh.nPoints, 3))
        self.U = np.zeros((nC, 3))
        self.lsq = orc.Lsq(mesh, geom, patchTypes)
        self.patchGradient = np.zeros((nBF, 3))
        self.Db = orc.patch_evaluate(mesh, geom, patchTypes, self.D, self.gradD, self.fixedValue, self.patchGradient)
        self.enforceLinear = False
        self.l = np.ascontiguousarray(mesh.owner[:nIF], dtype=np.int32)
        self.u = np.ascontiguousarray(mesh.neighbour, dtype=np.int32)

    @staticmethod
    def _tie(perf, hist, tolerance, relTol):
        """How close the stopping test of a PCG solve came to going the other way: the smallest relative distance of the residuals
        of its last two iterations from the bound they were compared with (max(tolerance, relTol*initialResidual)).  A solver
        whose sums are associated differently can stop one iteration earlier or later only where this is ~1e-12."""
        bound = max(tolerance, relTol*perf["initialResidual"])
        h = np.asarray(hist)[-2:] if len(hist) else np.array([perf["initialResidual"]])
        return float(np.min(np.abs(h - bound))/bound) if bound > 0 else 1.0

    def new_time_step(self):
        """runTime++ : storeOldTimes of D."""
        self.Doldold = self.Dold
        self.Dold = self.D.copy()
        if self.old is not None:
            self.old.advance()

    def _backward(self, d2dt2, ddt, damping, deltaT0):
        """`backward` schemes: the levels live in self.old (created and shifted as GeometricField does); Dold / Doldold follow."""
        if self.old is None:
            self.old = OldTimes()
            self.old.now = self.old.ti = 1          # the first time step of a run (startTime index 0, runTime++)
        e = backward_controls(self.old, self.D, d2dt2, ddt, damping, deltaT0)
        z = self.D
        self.Dold, self.Doldold = self.old.level(0, z), self.old.level(1, z)
        e["Dooo"], e["Doooo"] = self.old.level(2, z), self.old.level(3, z)
        return e

    def evolve(self, nCorrectors, convergenceTolerance, relConvergenceTolerance=0.0, tolerance=1e-9, relTol=0.01,
               d2dt2="steadyState", ddt="steadyState", deltaT=1.0, deltaT0=None, K=0.0, nonLinear=False, thermal=False,
               limitCoeff=1.0, relaxD=None, maxIter=1000, incremental=False):
        """incremental: the loop of unsIncrTotalLagrangianStress::evolve() (unsIncrTotalLagrangianStress.C:980-1159) -- this
        object's D is the increment DD (IncrSolidCase below): residual() = max|DD - DD.prevIter()|/(max|DD| + SMALL)
        (unsIncrTotalLagrangianStressSolve.C:695-716), loop while res > convergenceTolerance (:1155-1159)."""
        mesh, geom, pt = self.mesh, self.geom, self.patchTypes
        nIF = mesh.nInternalFaces
        self.enforceLinear = False                      # evolve() resets it (:812-813)
        iCorr, res, maxRes = 0, 1.0, 0.0
        hist, initialResidual, lastInit, totalIts, pcgIts, pcgTies = [], 0.0, 0.0, 0, [], []
        while True:
            Dprev = self.D.copy(); DbPrev = self.Db.copy()
            if incremental:
                self.gradDfBStart = self.gradDf[nIF:].copy()        # what DSigmaf_ (:1013-1018) is formed from in this iteration
            bw = None
            if "backward" in (d2dt2, ddt):
                bw = self._backward(d2dt2, ddt, K > 1e-15, deltaT if deltaT0 is None else deltaT0)
                self.lastBackward = {k: v for k, v in bw.items() if k.startswith("e")}
            ctl = orc.controls(d2dt2, ddt, deltaT, deltaT0, K, nonLinear and not self.enforceLinear, thermal, limitCoeff, bw)
            a = orc.assemble_D_ex(mesh, geom, pt, ctl, self.mu, self.lam, self.rho, self.g, self.D, self.Dold, self.Doldold,
                                  self.gradD, self.gradDf, self.traction, self.pressure, self.fixedValue, self.threeK,
                                  self.alpha, self.DT, self.DTb)
            self.lastAssembly = a
            patches = []
            for p in mesh.patches:
                if pt[p["name"]] == "empty":
                    continue
                b0 = p["start"] - nIF
                patches.append((np.ascontiguousarray(mesh.owner[p["start"]:p["start"] + p["size"]], dtype=np.int64),
                                a["internalCoeffs"][b0:b0 + p["size"]], a["boundaryCoeffs"][b0:b0 + p["size"]]))
            worst, perfs, hists = segregated_solve_ex(self.l, self.u, a["diag"], a["upper"], a["source"], patches, self.D,
                                                      tolerance, relTol, self.validCmpts, maxIter)
            totalIts += sum(p["nIterations"] for p in perfs.values())
            pcgIts.append([perfs[k]["nIterations"] for k in sorted(perfs)])
            pcgTies.append([self._tie(perfs[k], hists[k], tolerance, relTol) for k in sorted(perfs)])
            if iCorr == 0:
                initialResidual = worst["initialResidual"]
            lastInit = worst["initialResidual"]
            self.patchGradient = a["patchGradient"]
            self.Db = orc.patch_evaluate(mesh, geom, pt, self.D, self.gradD, self.fixedValue, self.patchGradient, self.Db)
            if relaxD is not None:
                orc.relax_boundary(self.Db, DbPrev, relaxD)
            res = orc.relax_residual(self.D, Dprev, np.zeros_like(self.D) if incremental else self.Dold, relaxD)
            self.pointD = self.lsq.interpolate(self.D, self.Db)
            if self.pointConstraints is not None:
                orc.point_constraints_apply(self.pointConstraints, self.pointD)
            gradOld = self.gradD
            self.gradD, self.gradDb = orc.grad_ex(mesh, geom, pt, self.D, self.pointD, gradOld, self.fixedValue,
                                                  self.patchGradient, self.gradDb)
            self.gradDf = orc.fgrad_ex(mesh, geom, pt, self.D, self.pointD, self.gradD, self.fixedValue, self.patchGradient,
                                       limitCoeff, self.gradDf)
            if nonLinear and not self.enforceLinear:
                mn, mx = orc.det_minmax(mesh, geom, pt, self.gradDf)
                if mn < 0.01 or mx > 100:
                    self.enforceLinear = True
            maxRes = max(maxRes, res)
            cur = max(maxRes*relConvergenceTolerance, convergenceTolerance)
            hist.append(res)
            if not (res > (convergenceTolerance if incremental else cur)):
                break
            iCorr += 1
            if not (iCorr < nCorrectors):
                break
        if ddt == "backward":
            # U = fvc::ddt(D): [EXT] backwardDdtScheme::fvcDdt with its own deltaT0_(D), then D.oldTime().oldTime() exists
            if self.old is None:
                self._backward(d2dt2, ddt, False, deltaT if deltaT0 is None else deltaT0)
            self.old.ensure(1, self.D)              # residual() has used D.oldTime()
            eU = GREAT if self.old.n_old(0) < 2 else (deltaT if deltaT0 is None else deltaT0)
            self.old.ensure(2, self.D)
            self.U = orc.ddt_backward(self.D, self.old.level(0, self.D), self.old.level(1, self.D), deltaT, eU)
            self.lastBackward = dict(getattr(self, "lastBackward", {}), eU=eU)
        else:
            self.U = orc.ddt(self.D, self.Dold, ddt, deltaT)
        self.epsilon, self.sigma = orc.sigma_ex(self.mu, self.lam, self.gradD, nonLinear, thermal, self.threeK, self.alpha, self.DT)
        return dict(nOuterIterations=iCorr, initialResidual=initialResidual, lastSolverInitialResidual=lastInit, maxRes=maxRes,
                    res=res, totalPcgIterations=totalIts, pcgIterations=pcgIts, pcgStopMargin=pcgTies, residualHistory=np.array(hist),
                    enforceLinear=self.enforceLinear, converged=not (nonLinear and self.enforceLinear))


def boundary_normals(mesh, geom):
    """patch().nf() [EXT fvPatch::nf]: Sf/magSf of the boundary faces."""
    nIF = mesh.nInternalFaces
    return geom.Sf[nIF:]/geom.magSf[nIF:, None]


def vector_dot_symm(n, S):
    """(n & S) for a vector and a symmTensor (xx xy xz yy yz zz) [EXT SymmTensorI.H operator&]: terms added x, y, z."""
    return np.stack([n[:, 0]*S[:, 0] + n[:, 1]*S[:, 1] + n[:, 2]*S[:, 2],
                     n[:, 0]*S[:, 1] + n[:, 1]*S[:, 3] + n[:, 2]*S[:, 4],
                     n[:, 0]*S[:, 2] + n[:, 1]*S[:, 4] + n[:, 2]*S[:, 5]], axis=1)


def face_stress_increment(muf, lambdaf, gradDDf):
    """DEpsilonf_ = symm(gradDDf_); DSigmaf_ = 2*muf_*DEpsilonf_ + I*(lambdaf_*tr(DEpsilonf_))   (unsIncrTotalLagrangianStress.C:1013-1022,
    linear branch).  gradDDf (n, 9) row-major; returns (DEpsilonf, DSigmaf) as symmTensors (n, 6)."""
    g = gradDDf
    eps = np.stack([g[:, 0], 0.5*(g[:, 1] + g[:, 3]), 0.5*(g[:, 2] + g[:, 6]), g[:, 4], 0.5*(g[:, 5] + g[:, 7]), g[:, 8]], axis=1)
    tr = eps[:, 0] + eps[:, 3] + eps[:, 5]
    lt = lambdaf*tr
    two = (2*muf)[:, None]*eps
    sig = two.copy()
    for k, one in ((0, 1.0), (1, 0.0), (2, 0.0), (3, 1.0), (4, 0.0), (5, 1.0)):
        sig[:, k] = two[:, k] + one*lt
    return eps, sig


class IncrSolidCase(object):
    """unsIncrTotalLagrangianStress, linear branch without plasticity (stressModels/unsIncrTotalLagrangianStress/
    unsIncrTotalLagrangianStress.C:946-1187, updateTotalFields :1212-1233; BC fvPatchFields/tractionDisplacementIncrement/
    tractionDisplacementIncrementFvPatchVectorField.C:147-307): the increment DD is solved for with the equation of the total
    model -- same laplacian, same explicit divergence, in gradDDf -- under the traction increment
        DTraction = (traction - pressure*n) - (n & sigmaf)          (:203-233)
    which makes tractionDisplacementIncrement's gradient() (:235-262) the total model's with DTraction as the traction and no
    pressure; the loop differs in residual() and in its exit test (SolidCase.evolve(incremental=True)); updateTotalFields()
    accumulates.  `inc` holds DD, pointDD, gradDD, gradDDf; the totals live here."""

    def __init__(self, mesh, geom, patchTypes, mu, lam, rho, traction=None, pressure=None, fixedValue=None, **kw):
        nC, nF, nIF = mesh.nCells, mesh.nFaces, mesh.nInternalFaces
        nBF = nF - nIF
        self.mesh, self.geom = mesh, geom
        self.inc = SolidCase(mesh, geom, patchTypes, mu, lam, rho, fixedValue=fixedValue, **kw)
        self.traction = np.zeros((nBF, 3)) if traction is None else np.asarray(traction, dtype=np.float64)
        self.pressure = np.zeros(nBF) if pressure is None else np.asarray(pressure, dtype=np.float64)
        own = mesh.owner[nIF:]
        self.mufB, self.lambdafB = self.inc.mu[own].copy(), self.inc.lam[own].copy()        # boundary value of a linear interpolate
        self.D = np.zeros((nC, 3)); self.pointD = np.zeros((mesh.nPoints, 3)); self.gradD = np.zeros((nC, 9))
        self.gradDf = np.zeros((nF, 9)); self.sigma = np.zeros((nC, 6)); self.epsilon = np.zeros((nC, 6))
        self.sigmafB = np.zeros((nBF, 6)); self.epsilonfB = np.zeros((nBF, 6))

    def new_time_step(self):
        self.inc.new_time_step()

    def evolve(self, *a, **kw):
        n = boundary_normals(self.mesh, self.geom)
        t = self.traction - self.pressure[:, None]*n
        self.inc.traction = t - vector_dot_symm(n, self.sigmafB)
        self.inc.pressure = np.zeros_like(self.pressure)
        r = self.inc.evolve(*a, incremental=True, **kw)
        self.DEpsilonfB, self.DSigmafB = face_stress_increment(self.mufB, self.lambdafB, self.inc.gradDfBStart)
        return r

    def update_total_fields(self):
        """updateTotalFields() (:1212-1233)."""
        i = self.inc
        self.D = self.D + i.D; self.pointD = self.pointD + i.pointD; self.gradD = self.gradD + i.gradD
        self.gradDf = self.gradDf + i.gradDf; self.sigma = self.sigma + i.sigma; self.epsilon = self.epsilon + i.epsilon
        self.sigmafB = self.sigmafB + self.DSigmafB; self.epsilonfB = self.epsilonfB + self.DEpsilonfB
