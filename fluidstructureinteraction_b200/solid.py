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
               limitCoeff=1.0, relaxD=None, maxIter=1000):
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
                              relaxD, self.validCmpts, self.g, eqn)
        # U, epsilon, sigma (:973-990): `nonLinear` here is the dictionary switch, whatever enforceLinear became
        if ddt == "backward":
            eqn.eU = self.old.velocity(deltaT if deltaT0 is None else deltaT0)
        self.gm.sigma_resident(eqn)
        return r

    def close(self):
        self.pcg.close(); self.gm.close()
