"""Harness over the device-resident solid model: the state an `unsTotalLagrangianStress` object holds
(stressModels/unsTotalLagrangianStress/unsTotalLagrangianStress.C:59-183: D, pointD, gradD, gradDf, mu/lambda/rho, the
volToPoint_ interpolation) lives on the GPU across outer iterations and time steps; evolve() is one call into
gpupcg_evolve_D.  Fields cross PCIe only when the caller asks for them (get)."""
import numpy as np

from . import capi, lsq as lsqmod, timelevels


class GpuSolidCase(object):
    def __init__(self, mesh, geom, patchTypes, mu, lam, rho, traction=None, pressure=None, fixedValue=None, g=(0.0, 0.0, 0.0),
                 threeK=None, alpha=None, validCmpts=(1, 1, 1), device=0, lsq=None,
                 pointPatchTypes=None, pointPatchValues=None):
        nC, nF, nIF = mesh.nCells, mesh.nFaces, mesh.nInternalFaces
        nBF = nF - nIF
        self.mesh, self.geom, self.patchTypes = mesh, geom, patchTypes
        self.validCmpts = tuple(validCmpts)
        self.g = tuple(float(v) for v in g)
        full = lambda v, n: np.full(n, float(v)) if np.isscalar(v) else np.ascontiguousarray(v, dtype=np.float64)
        self.gm = capi.GpuMesh(mesh, geom, patchTypes, device=device)
        l, u = mesh.lower_upper()
        self.pcg = capi.GpuPCG(l, u, nC, device=device)
        gm = self.gm
        gm.set_field("mu", full(mu, nC)); gm.set_field("lambda_", full(lam, nC)); gm.set_field("rho", full(rho, nC))
        gm.set_field("traction", np.zeros((nBF, 3)) if traction is None else traction)
        gm.set_field("pressure", np.zeros(nBF) if pressure is None else pressure)
        gm.set_field("fixedValue", np.zeros((nBF, 3)) if fixedValue is None else fixedValue)
        for name, n in (("D", (nC, 3)), ("Dold", (nC, 3)), ("Doldold", (nC, 3)), ("gradD", (nC, 9)), ("gradDf", (nF, 9)),
                        ("patchGradient", (nBF, 3))):
            gm.set_field(name, np.zeros(n))
        self.old = None                 # timelevels.OldTimes once a `backward` scheme is used
        self.thermal = threeK is not None
        if self.thermal:
            gm.set_field("threeK", full(threeK, nC)); gm.set_field("alpha", full(alpha, nC))
            gm.set_field("DT", np.zeros(nC)); gm.set_field("DTb", np.zeros(nBF))
        gm.set_lsq(lsq if lsq is not None else lsqmod.build(mesh, geom, patchTypes))
        if pointPatchTypes:             # 0/pointD: symmetryPlane / fixedValue point patches act after every interpolation
            gm.set_point_constraints(lsqmod.point_constraints(mesh, geom, pointPatchTypes, pointPatchValues))
        gm.set_field("pointD", np.zeros((mesh.nPoints, 3)))
        gm.set_field("Db", np.zeros((nBF, 3)))
        gm.patch_evaluate()             # D.boundaryField() of the initial field

    def set(self, name, a):
        self.gm.set_field(name, a)

    def get(self, name):
        return self.gm.get_field(name)

    def new_time_step(self):
        if self.old is not None:        # `backward` schemes: the levels shift when GeometricField would shift them (evolve)
            self.old.advance()
        else:
            self.gm.new_time_step()

    def evolve(self, nCorrectors, convergenceTolerance, relConvergenceTolerance=0.0, tolerance=1e-9, relTol=0.01,
               d2dt2="steadyState", ddt="steadyState", deltaT=1.0, deltaT0=None, K=0.0, nonLinear=False, thermal=False,
               limitCoeff=1.0, relaxD=None, maxIter=1000, incremental=False):
        bw = None
        if "backward" in (d2dt2, ddt):
            dT0 = deltaT if deltaT0 is None else deltaT0
            if self.old is None:
                self.old = timelevels.OldTimes(self.gm)
            # the effective old time steps of this evolve(): those of its first assembly (they can differ from the later
            # outer iterations' only in the very first evolve of a run, where every old level still equals the initial field)
            bw = self.old.assembly(d2dt2, ddt, K > 1e-15, dT0)
            bw["eU"] = dT0
        eqn = capi.eqn_controls(d2dt2, ddt, deltaT, deltaT0, K, nonLinear, thermal, limitCoeff, bw)
        r = self.pcg.evolve_D(self.gm, nCorrectors, convergenceTolerance, relConvergenceTolerance, tolerance, relTol, 0, maxIter,
                              relaxD, self.validCmpts, self.g, eqn, incremental=incremental)
        # U, epsilon, sigma (:973-990): `nonLinear` here is the dictionary switch, whatever enforceLinear became
        if ddt == "backward":
            eqn.eU = self.old.velocity(deltaT if deltaT0 is None else deltaT0)
        self.gm.sigma_resident(eqn)
        return r

    def close(self):
        self.pcg.close(); self.gm.close()


class GpuIncrSolidCase(object):
    """unsIncrTotalLagrangianStress, linear branch (stressModels/unsIncrTotalLagrangianStress/unsIncrTotalLagrangianStress.C:946-1187,
    updateTotalFields :1212-1233): the increment DD is the device-resident field of a GpuSolidCase -- same kernels, the loop of
    gpupcg_evolve_D with `incremental` set -- under the traction increment of tractionDisplacementIncrementFvPatchVectorField.C:203-233,
    DTraction = (traction - pressure*n) - (n & sigmaf).  What happens once per TIME STEP is host arithmetic here: the traction
    increment going in, DSigmaf of the boundary faces and updateTotalFields() coming out (the totals D, pointD, gradD, gradDf,
    sigma, epsilon are numpy arrays of this object)."""

    def __init__(self, mesh, geom, patchTypes, mu, lam, rho, traction=None, pressure=None, fixedValue=None, **kw):
        nC, nF, nIF = mesh.nCells, mesh.nFaces, mesh.nInternalFaces
        nBF = nF - nIF
        self.mesh, self.geom = mesh, geom
        self.inc = GpuSolidCase(mesh, geom, patchTypes, mu, lam, rho, fixedValue=fixedValue, **kw)
        self.traction = np.zeros((nBF, 3)) if traction is None else np.asarray(traction, dtype=np.float64)
        self.pressure = np.zeros(nBF) if pressure is None else np.asarray(pressure, dtype=np.float64)
        full = lambda v: np.full(nC, float(v)) if np.isscalar(v) else np.asarray(v, dtype=np.float64)
        own = mesh.owner[nIF:]
        self.mufB, self.lambdafB = full(mu)[own].copy(), full(lam)[own].copy()
        self.n = geom.Sf[nIF:]/geom.magSf[nIF:, None]                       # patch().nf()
        self.D = np.zeros((nC, 3)); self.pointD = np.zeros((mesh.nPoints, 3)); self.gradD = np.zeros((nC, 9))
        self.gradDf = np.zeros((nF, 9)); self.sigma = np.zeros((nC, 6)); self.epsilon = np.zeros((nC, 6))
        self.sigmafB = np.zeros((nBF, 6)); self.epsilonfB = np.zeros((nBF, 6))

    def new_time_step(self):
        self.inc.new_time_step()

    def evolve(self, *a, **kw):
        n, S = self.n, self.sigmafB
        nS = np.stack([n[:, 0]*S[:, 0] + n[:, 1]*S[:, 1] + n[:, 2]*S[:, 2],            # (n & sigmaf), terms added x, y, z
                       n[:, 0]*S[:, 1] + n[:, 1]*S[:, 3] + n[:, 2]*S[:, 4],
                       n[:, 0]*S[:, 2] + n[:, 1]*S[:, 4] + n[:, 2]*S[:, 5]], axis=1)
        self.inc.set("traction", (self.traction - self.pressure[:, None]*n) - nS)
        self.inc.set("pressure", np.zeros_like(self.pressure))
        r = self.inc.evolve(*a, incremental=True, **kw)
        # DEpsilonf_ = symm(gradDDf_), DSigmaf_ = 2*muf_*DEpsilonf_ + I*(lambdaf_*tr(DEpsilonf_)) of the boundary faces, from the
        # gradDDf the LAST outer iteration started with (:1013-1022 -- not re-formed after its solve)
        g = self.inc.get("gradDfBStart").reshape(-1, 9)
        eps = np.stack([g[:, 0], 0.5*(g[:, 1] + g[:, 3]), 0.5*(g[:, 2] + g[:, 6]), g[:, 4], 0.5*(g[:, 5] + g[:, 7]), g[:, 8]], axis=1)
        lt = self.lambdafB*(eps[:, 0] + eps[:, 3] + eps[:, 5])
        two = (2*self.mufB)[:, None]*eps
        sig = two.copy()
        for k, one in ((0, 1.0), (1, 0.0), (2, 0.0), (3, 1.0), (4, 0.0), (5, 1.0)):
            sig[:, k] = two[:, k] + one*lt
        self.DEpsilonfB, self.DSigmafB = eps, sig
        return r

    def update_total_fields(self):
        """updateTotalFields() (:1212-1233): one read of the increment fields per time step."""
        i = self.inc
        self.D = self.D + i.get("D").reshape(-1, 3); self.pointD = self.pointD + i.get("pointD").reshape(-1, 3)
        self.gradD = self.gradD + i.get("gradD").reshape(-1, 9); self.gradDf = self.gradDf + i.get("gradDf").reshape(-1, 9)
        self.sigma = self.sigma + i.get("sigma").reshape(-1, 6); self.epsilon = self.epsilon + i.get("epsilon").reshape(-1, 6)
        self.sigmafB = self.sigmafB + self.DSigmafB; self.epsilonfB = self.epsilonfB + self.DEpsilonfB

    def close(self):
        self.inc.close()
