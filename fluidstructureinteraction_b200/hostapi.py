"""ctypes binding of host/libfoamlike.so: the C++ host side (foam-extend API stand-in + the gpuPCG adaptor source that
wmake compiles against foam-extend + the segregated fvMatrix solve glue), driven the way foam-extend drives it:
a solver sub-dictionary in fvSolution syntax -> lduMatrix::solver::New -> solve."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "host", "libfoamlike.so")

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


class Perf(C.Structure):
    _fields_ = [("initialResidual", C.c_double), ("finalResidual", C.c_double), ("nIterations", C.c_int),
                ("converged", C.c_int), ("singular", C.c_int), ("pad", C.c_int),
                ("solverName", C.c_char*32), ("fieldName", C.c_char*32)]

    def as_dict(self):
        return dict(initialResidual=self.initialResidual, finalResidual=self.finalResidual, nIterations=self.nIterations,
                    converged=self.converged, singular=self.singular, solverName=self.solverName.decode(),
                    fieldName=self.fieldName.decode())


class FoamFatalError(RuntimeError):
    pass


_lib = None


def load():
    global _lib
    if _lib is None:
        from . import capi
        capi.load_library()         # libgpupcg_cuda.so first (RTLD_GLOBAL)
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError("%s is not built (make -C %s)" % (LIB_PATH, os.path.dirname(LIB_PATH)))
        lib = C.CDLL(LIB_PATH)
        lib.foamlike_last_error.restype = C.c_char_p
        lib.foamlike_solve.restype = C.c_int
        lib.foamlike_solve.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, ip, ip, dp, dp, dp, C.c_int, ip, ip, dp, dp,
                                       C.c_char_p, dp, C.POINTER(Perf), C.POINTER(Perf)]
        lib.foamlike_dict_lookup.restype = C.c_int
        lib.foamlike_dict_lookup.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
        lib.foamlike_evolve.restype = C.c_int
        _lib = lib
    return _lib


def dict_lookup(text, key):
    lib = load()
    buf = C.create_string_buffer(256)
    if lib.foamlike_dict_lookup(text.encode(), key.encode(), buf, 256) != 0:
        raise FoamFatalError(lib.foamlike_last_error().decode())
    return buf.value.decode()


def solve(psiName, l, u, diag, upper, source, patches, solverDict, psi):
    """fvMatrix<Type>::solve(solverDict).  source/psi: (nCells,) or (nCells,3); patches: list of
    (faceCells, internalCoeffs, boundaryCoeffs) with coefficient arrays shaped like source rows.
    psi is updated in place.  Returns (worst perf dict, [per-component perf dicts])."""
    lib = load()
    source = np.ascontiguousarray(source, dtype=np.float64)
    nC = source.shape[0]
    nCmpt = 1 if source.ndim == 1 else source.shape[1]
    l = np.ascontiguousarray(l, dtype=np.int32); u = np.ascontiguousarray(u, dtype=np.int32)
    diag = np.ascontiguousarray(diag, dtype=np.float64)
    upper = np.ascontiguousarray(upper, dtype=np.float64) if upper is not None else None
    sizes = [len(p[0]) for p in patches]
    start = np.ascontiguousarray(np.concatenate([[0], np.cumsum(sizes)]), dtype=np.int32)
    if patches:
        fc = np.ascontiguousarray(np.concatenate([p[0] for p in patches]), dtype=np.int32)
        ic = np.ascontiguousarray(np.concatenate([np.reshape(p[1], (len(p[0]), nCmpt)) for p in patches]), dtype=np.float64)
        bc = np.ascontiguousarray(np.concatenate([np.reshape(p[2], (len(p[0]), nCmpt)) for p in patches]), dtype=np.float64)
    else:
        fc = np.zeros(1, dtype=np.int32); ic = np.zeros(1); bc = np.zeros(1)
    assert psi.dtype == np.float64 and psi.flags.c_contiguous and psi.shape == source.shape
    per = (Perf*3)()
    worst = Perf()
    n = lib.foamlike_solve(psiName.encode(), nCmpt, nC, l.shape[0], l.ctypes.data_as(ip), u.ctypes.data_as(ip),
                           diag.ctypes.data_as(dp), upper.ctypes.data_as(dp) if upper is not None else None,
                           source.ctypes.data_as(dp), len(patches), start.ctypes.data_as(ip), fc.ctypes.data_as(ip),
                           ic.ctypes.data_as(dp), bc.ctypes.data_as(dp), solverDict.encode(), psi.ctypes.data_as(dp),
                           per, C.byref(worst))
    if n < 0:
        raise FoamFatalError(lib.foamlike_last_error().decode())
    return worst.as_dict(), [per[k].as_dict() for k in range(n)]


def solve_twice(l, u, diag1, diag2, upper, changeFace, delta, source, solverDict):
    """Two scalar solves on one lduMatrix object (same upper() address); before the second, upper[changeFace] += delta and
    the diagonal becomes diag2.  Returns (psi1, psi2)."""
    lib = load()
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    l, u, diag1, diag2, upper, source = i32(l), i32(u), f64(diag1), f64(diag2), f64(upper), f64(source)
    n = diag1.shape[0]
    p1, p2 = np.zeros(n), np.zeros(n)
    lib.foamlike_solve_twice.restype = C.c_int
    lib.foamlike_solve_twice.argtypes = [C.c_int, C.c_int, ip, ip, dp, dp, dp, C.c_int, C.c_double, dp, C.c_char_p, dp, dp]
    rc = lib.foamlike_solve_twice(n, l.shape[0], l.ctypes.data_as(ip), u.ctypes.data_as(ip), diag1.ctypes.data_as(dp),
                                  diag2.ctypes.data_as(dp), upper.ctypes.data_as(dp), int(changeFace), float(delta),
                                  source.ctypes.data_as(dp), solverDict.encode(), p1.ctypes.data_as(dp), p2.ctypes.data_as(dp))
    if rc < 0:
        raise FoamFatalError(lib.foamlike_last_error().decode())
    return p1, p2


VOLTOPOINT = C.CFUNCTYPE(None, C.c_void_p, dp, dp)


def evolve(mesh, geom, patchTypes, mu, lam, rho, g, traction, pressure, fixedValue, volToPoint, stressProperties, solverDict,
           D=None, Dold=None, gradD=None, gradDf=None, limitCoeff=1.0, relaxD=None, device=0):
    """One time step of the solid model: unsTotalLagrangianStress::evolve() (unsTotalLagrangianStress.C:769-1009) driven by
    the C++ host side (host/stressModelEvolve.C) with assembly, PCG, relax/residual, gradients and stress on the GPU.
    volToPoint(D (nC,3)) -> pointD (nP,3) stands in for leastSquaresVolPointInterpolation.
    stressProperties / solverDict: dictionary text in foam syntax.  Returns a dict of fields and loop statistics."""
    from .capi import PATCH_TYPES
    lib = load()
    nC, nF, nIF, nP = mesh.nCells, mesh.nFaces, mesh.nInternalFaces, mesh.nPoints
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    ps = i32([p["start"] for p in mesh.patches]); pz = i32([p["size"] for p in mesh.patches])
    pt = i32([PATCH_TYPES[patchTypes[p["name"]]] for p in mesh.patches])
    ints = i32([nC, nF, nIF, len(mesh.patches), nP, device])
    mi = [i32(mesh.owner), i32(mesh.neighbour), ps, pz, pt, i32(mesh.faceStart), i32(mesh.facePoints)]
    mr = [f64(geom.Sf), f64(geom.magSf), f64(geom.Cf), f64(geom.C), f64(geom.V), f64(geom.weights), f64(geom.deltaCoeffs),
          f64(mesh.points)]
    fl = [f64(mu), f64(lam), f64(rho), f64(g), f64(traction), f64(pressure), f64(fixedValue)]
    MI = (ip*len(mi))(*[a.ctypes.data_as(ip) for a in mi])
    MR = (dp*len(mr))(*[a.ctypes.data_as(dp) for a in mr])
    FL = (dp*len(fl))(*[a.ctypes.data_as(dp) for a in fl])
    Dv = f64(np.zeros((nC, 3)) if D is None else D).copy(); Do = f64(np.zeros((nC, 3)) if Dold is None else Dold)
    pD = np.zeros((nP, 3)); G = f64(np.zeros((nC, 9)) if gradD is None else gradD).copy()
    Gf = f64(np.zeros((nF, 9)) if gradDf is None else gradDf).copy()
    eps = np.zeros((nC, 6)); sig = np.zeros((nC, 6)); res = np.zeros(8); hist = np.zeros(1024)

    def cb(ctx, dptr, pptr):
        Din = np.ctypeslib.as_array(dptr, shape=(nC, 3))
        out = np.ctypeslib.as_array(pptr, shape=(nP, 3))
        out[...] = volToPoint(Din)

    cfn = VOLTOPOINT(cb)
    n = lib.foamlike_evolve(ints.ctypes.data_as(ip), MI, MR, FL, C.c_double(limitCoeff), stressProperties.encode(),
                            solverDict.encode(), C.c_double(-1.0 if relaxD is None else relaxD), cfn, None,
                            Dv.ctypes.data_as(dp), Do.ctypes.data_as(dp), pD.ctypes.data_as(dp), G.ctypes.data_as(dp),
                            Gf.ctypes.data_as(dp), eps.ctypes.data_as(dp), sig.ctypes.data_as(dp), res.ctypes.data_as(dp),
                            hist.ctypes.data_as(dp), C.c_int(hist.shape[0]))
    if n < 0:
        raise FoamFatalError(lib.foamlike_last_error().decode())
    return dict(D=Dv, pointD=pD, gradD=G, gradDf=Gf, epsilon=eps, sigma=sig, converged=bool(res[0]), nOuterIterations=int(res[1]),
                initialResidual=res[2], lastSolverInitialResidual=res[3], maxRes=res[4], res=res[5],
                totalPcgIterations=int(res[6]), residualHistory=hist[:n].copy())


class StressModel(object):
    """`stressModel::New(mesh)` of the C++ host side (host/gpuStressModel.{H,C}): the model named by `stressModel ...;` in the
    stressProperties text is selected from the run-time table -- `gpuUnsTotalLagrangianStress` keeps the solid on the GPU and
    runs evolve() as one gpupcg_evolve_D call.  Dictionaries are foam-syntax text, the boundary fields of 0/D per boundary
    face; `lsq` / `pointConstraints` as lsq.build / lsq.point_constraints return them."""

    def __init__(self, mesh, geom, patchFieldTypes, stressProperties, rheologyProperties, fvSolution, fvSchemes, deltaT=1.0,
                 traction=None, pressure=None, fixedValue=None, lsq=None, pointConstraints=None, solutionD=(1, 1, 1), device=0):
        from . import lsq as lsqmod
        lib = load()
        nC, nF, nIF, nP = mesh.nCells, mesh.nFaces, mesh.nInternalFaces, mesh.nPoints
        nBF = nF - nIF
        self.nCells, self.nPoints = nC, nP
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        if lsq is None:
            lsq = lsqmod.build(mesh, geom, patchFieldTypes)
        pc = pointConstraints or dict(point=np.zeros(0, np.int32), start=np.zeros(1, np.int32), kind=np.zeros(0, np.int32), vec=np.zeros((0, 3)))
        mi = [i32(mesh.owner), i32(mesh.neighbour), i32([p["start"] for p in mesh.patches]), i32([p["size"] for p in mesh.patches]),
              i32(mesh.faceStart), i32(mesh.facePoints), i32(lsq["start"]), i32(lsq["entry"]), i32(pc["point"]), i32(pc["start"]),
              i32(pc["kind"])]
        mr = [f64(geom.Sf), f64(geom.magSf), f64(geom.Cf), f64(geom.C), f64(geom.V), f64(geom.weights), f64(geom.deltaCoeffs),
              f64(mesh.points), f64(np.zeros((nBF, 3)) if traction is None else traction),
              f64(np.zeros(nBF) if pressure is None else pressure), f64(np.zeros((nBF, 3)) if fixedValue is None else fixedValue),
              f64(lsq["inv"]), f64(lsq["origin"]), f64(pc["vec"])]
        mirror = np.ascontiguousarray(lsq["mirror"], dtype=np.uint8)
        ints = i32([nC, nF, nIF, len(mesh.patches), nP, device, int(pc["point"].shape[0])] + [1 if s else -1 for s in solutionD])
        MI = (ip*len(mi))(*[a.ctypes.data_as(ip) for a in mi])
        MR = (dp*len(mr))(*[a.ctypes.data_as(dp) for a in mr])
        IL = (C.c_longlong*len(mi))(*[a.size for a in mi])
        RL = (C.c_longlong*len(mr))(*[a.size for a in mr])
        lib.foamlike_stress_new.restype = C.c_void_p
        lib.foamlike_stress_delete.argtypes = [C.c_void_p]
        lib.foamlike_stress_set.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
        lib.foamlike_stress_new_time_step.argtypes = [C.c_void_p]
        lib.foamlike_stress_evolve.argtypes = [C.c_void_p, dp, dp, ip, C.c_int]
        lib.foamlike_stress_field.restype = C.c_longlong
        lib.foamlike_stress_field.argtypes = [C.c_void_p, C.c_char_p, dp, C.c_longlong]
        self.lib = lib
        self.h = lib.foamlike_stress_new(ints.ctypes.data_as(ip), MI, IL, MR, RL, mirror.ctypes.data_as(C.POINTER(C.c_ubyte)),
                                         " ".join(p["name"] for p in mesh.patches).encode(),
                                         " ".join(patchFieldTypes[p["name"]] for p in mesh.patches).encode(),
                                         stressProperties.encode(), rheologyProperties.encode(), fvSolution.encode(),
                                         fvSchemes.encode(), C.c_double(deltaT))
        if not self.h:
            raise FoamFatalError(lib.foamlike_last_error().decode())
        self.h = C.c_void_p(self.h)
        self.patchIndex = {p["name"]: i for i, p in enumerate(mesh.patches)}

    def _check(self, rc):
        if rc < 0:
            raise FoamFatalError(self.lib.foamlike_last_error().decode())
        return rc

    def set_traction(self, patch, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        self._check(self.lib.foamlike_stress_set(self.h, 0, self.patchIndex[patch], v.ctypes.data_as(dp)))

    def set_pressure(self, patch, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        self._check(self.lib.foamlike_stress_set(self.h, 1, self.patchIndex[patch], v.ctypes.data_as(dp)))

    def new_time_step(self):
        self._check(self.lib.foamlike_stress_new_time_step(self.h))

    def evolve(self, cap=4096):
        res = np.zeros(9); hist = np.zeros(cap); pcg = np.zeros(3*cap, dtype=np.int32)
        n = self._check(self.lib.foamlike_stress_evolve(self.h, res.ctypes.data_as(dp), hist.ctypes.data_as(dp), pcg.ctypes.data_as(ip), cap))
        return dict(converged=bool(res[0]), nOuterIterations=int(res[1]), totalPcgIterations=int(res[2]), enforceLinear=int(res[3]),
                    initialResidual=res[4], lastSolverInitialResidual=res[5], maxRes=res[6], res=res[7], residualHistory=hist[:n].copy(),
                    pcgIterations=pcg[:3*n].reshape(n, 3)[:, :int(res[8])].copy())

    def get(self, name):
        n = self._check(self.lib.foamlike_stress_field(self.h, name.encode(), None, 0))
        a = np.zeros(n)
        self._check(self.lib.foamlike_stress_field(self.h, name.encode(), a.ctypes.data_as(dp), n))
        return a

    def close(self):
        if getattr(self, "h", None):
            self.lib.foamlike_stress_delete(self.h)
            self.h = None
