"""ctypes binding of oracle/liboracle.so (gcc -O3 -ffp-contract=off; see oracle/Makefile)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


class Domain(C.Structure):
    _fields_ = [
        ("nCells", C.c_int), ("nFaces", C.c_int),
        ("lowerAddr", ip), ("upperAddr", ip),
        ("diag", dp), ("upper", dp),
        ("nPatches", C.c_int),
        ("patchStart", ip), ("patchFaceCells", ip), ("patchBouCoeffs", dp),
        ("patchNbrDomain", ip), ("patchNbrPatch", ip),
        ("x", dp), ("b", dp),
    ]


class Mesh(C.Structure):
    _fields_ = [("nCells", C.c_int), ("nFaces", C.c_int), ("nInternalFaces", C.c_int), ("nPatches", C.c_int),
                ("owner", ip), ("neighbour", ip), ("patchStart", ip), ("patchSize", ip), ("patchType", ip),
                ("Sf", dp), ("magSf", dp), ("Cf", dp), ("C", dp), ("V", dp), ("weights", dp), ("deltaCoeffs", dp),
                ("nPoints", C.c_int), ("faceStart", ip), ("facePoints", ip), ("points", dp)]


class Perf(C.Structure):
    _fields_ = [
        ("initialResidual", C.c_double), ("finalResidual", C.c_double), ("normFactor", C.c_double),
        ("nIterations", C.c_int), ("converged", C.c_int), ("singular", C.c_int),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def load(build=True):
    global _lib
    if _lib is not None:
        return _lib
    if build and not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    lib = C.CDLL(LIB_PATH)
    lib.orc_amul.restype = None
    lib.orc_amul.argtypes = [C.c_int, C.c_int, ip, ip, dp, dp, dp, dp, C.c_int, ip, ip, dp, dp]
    lib.orc_sumA.restype = None
    lib.orc_sumA.argtypes = [C.c_int, C.c_int, ip, ip, dp, dp, dp, C.c_int, ip, ip, dp, dp]
    lib.orc_dic_calc_reciprocal_d.restype = None
    lib.orc_dic_calc_reciprocal_d.argtypes = [C.c_int, C.c_int, ip, ip, dp, dp, dp]
    lib.orc_dic_precondition.restype = None
    lib.orc_dic_precondition.argtypes = [C.c_int, C.c_int, ip, ip, dp, dp, dp, dp]
    lib.orc_pcg_solve.restype = C.c_int
    lib.orc_pcg_solve.argtypes = [C.c_int, C.POINTER(Domain), C.c_double, C.c_double, C.c_int, C.c_int,
                                  C.POINTER(Perf), dp, C.c_int]
    lib.orc_bench_amul.restype = C.c_double
    lib.orc_bench_amul.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip, dp, dp, dp, dp]
    lib.orc_bench_dic.restype = C.c_double
    lib.orc_bench_dic.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip, dp, dp, dp, dp]
    lib.orc_set_sum_mode.argtypes = [C.c_int]
    lib.orc_lame.argtypes = [C.c_double, C.c_double, C.c_int, dp, dp]
    lib.orc_assemble_D.restype = None
    lib.orc_assemble_D.argtypes = [C.POINTER(Mesh)] + [dp]*10 + [C.c_double] + [dp]*5
    lib.orc_grad.restype = None
    lib.orc_grad.argtypes = [C.POINTER(Mesh)] + [dp]*7
    lib.orc_fgrad.restype = None
    lib.orc_fgrad.argtypes = [C.POINTER(Mesh)] + [dp]*5 + [C.c_double, dp]
    lib.orc_sigma.restype = None
    lib.orc_sigma.argtypes = [C.c_int, dp, dp, dp, dp, dp]
    lib.orc_num_threads.restype = C.c_int
    lib.orc_set_num_threads.argtypes = [C.c_int]
    _lib = lib
    return lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _dp(a):
    return a.ctypes.data_as(dp)


def _ip(a):
    return a.ctypes.data_as(ip)


def _patch_arrays(patchFaceCells, patchBouCoeffs):
    sizes = [len(fc) for fc in patchFaceCells]
    start = _i(np.concatenate([[0], np.cumsum(sizes)]))
    fc = _i(np.concatenate(patchFaceCells)) if sizes and sum(sizes) else _i(np.zeros(1))
    bc = _f(np.concatenate(patchBouCoeffs)) if sizes and sum(sizes) else _f(np.zeros(1))
    return start, fc, bc


def amul(l, u, diag, upper, x, patchFaceCells=(), patchBouCoeffs=(), xNbr=None):
    lib = load()
    l, u, diag, upper, x = _i(l), _i(u), _f(diag), _f(upper), _f(x)
    Ax = np.empty_like(x)
    start, fc, bc = _patch_arrays(list(patchFaceCells), list(patchBouCoeffs))
    xn = _f(xNbr) if xNbr is not None else _f(np.zeros(max(1, int(start[-1]))))
    lib.orc_amul(x.shape[0], l.shape[0], _ip(l), _ip(u), _dp(diag), _dp(upper), _dp(x), _dp(Ax),
                 len(patchFaceCells), _ip(start), _ip(fc), _dp(bc), _dp(xn))
    return Ax


def dic_factor(l, u, diag, upper):
    lib = load()
    l, u, diag, upper = _i(l), _i(u), _f(diag), _f(upper)
    rD = np.empty_like(diag)
    lib.orc_dic_calc_reciprocal_d(diag.shape[0], l.shape[0], _ip(l), _ip(u), _dp(diag), _dp(upper), _dp(rD))
    return rD


def precondition(l, u, upper, rD, rA):
    lib = load()
    l, u, upper, rD, rA = _i(l), _i(u), _f(upper), _f(rD), _f(rA)
    wA = np.empty_like(rA)
    lib.orc_dic_precondition(rA.shape[0], l.shape[0], _ip(l), _ip(u), _dp(upper), _dp(rD), _dp(rA), _dp(wA))
    return wA


def pcg_solve_multi(domains, tolerance=1e-6, relTol=0.0, minIter=0, maxIter=1000):
    """domains: list of dict(l, u, diag, upper, x, b, patchFaceCells=[], patchBouCoeffs=[], patchNbrDomain=[],
    patchNbrPatch=[]).  x arrays are updated in place.  Returns (perf dict, residual history)."""
    lib = load()
    nD = len(domains)
    arr = (Domain * nD)()
    keep = []
    for d, dom in enumerate(domains):
        l, u, diag, upper = _i(dom["l"]), _i(dom["u"]), _f(dom["diag"]), _f(dom["upper"])
        x = dom["x"]
        assert isinstance(x, np.ndarray) and x.dtype == np.float64 and x.flags.c_contiguous
        b = _f(dom["b"])
        pfc = list(dom.get("patchFaceCells", []))
        pbc = list(dom.get("patchBouCoeffs", []))
        start, fc, bc = _patch_arrays(pfc, pbc)
        nbrD = _i(dom.get("patchNbrDomain", [])) if pfc else _i(np.zeros(1))
        nbrP = _i(dom.get("patchNbrPatch", [])) if pfc else _i(np.zeros(1))
        keep.append((l, u, diag, upper, b, start, fc, bc, nbrD, nbrP))
        D = arr[d]
        D.nCells, D.nFaces = x.shape[0], l.shape[0]
        D.lowerAddr, D.upperAddr, D.diag, D.upper = _ip(l), _ip(u), _dp(diag), _dp(upper)
        D.nPatches = len(pfc)
        D.patchStart, D.patchFaceCells, D.patchBouCoeffs = _ip(start), _ip(fc), _dp(bc)
        D.patchNbrDomain, D.patchNbrPatch = _ip(nbrD), _ip(nbrP)
        D.x, D.b = _dp(x), _dp(b)
    perf = Perf()
    hist = np.zeros(max(1, maxIter), dtype=np.float64)
    rc = lib.orc_pcg_solve(nD, arr, tolerance, relTol, minIter, maxIter, C.byref(perf), _dp(hist), hist.shape[0])
    if rc != 0:
        raise RuntimeError("orc_pcg_solve failed: %d" % rc)
    p = perf.as_dict()
    return p, hist[:p["nIterations"]].copy()


def pcg_solve(l, u, diag, upper, x, b, tolerance=1e-6, relTol=0.0, minIter=0, maxIter=1000):
    return pcg_solve_multi([dict(l=l, u=u, diag=diag, upper=upper, x=x, b=b)], tolerance, relTol, minIter, maxIter)


def bench_amul(reps, l, u, diag, upper, x):
    lib = load()
    l, u, diag, upper, x = _i(l), _i(u), _f(diag), _f(upper), _f(x)
    Ax = np.empty_like(x)
    return lib.orc_bench_amul(reps, x.shape[0], l.shape[0], _ip(l), _ip(u), _dp(diag), _dp(upper), _dp(x), _dp(Ax))


def bench_dic(reps, l, u, upper, rD, rA):
    lib = load()
    l, u, upper, rD, rA = _i(l), _i(u), _f(upper), _f(rD), _f(rA)
    wA = np.empty_like(rA)
    return lib.orc_bench_dic(reps, rA.shape[0], l.shape[0], _ip(l), _ip(u), _dp(upper), _dp(rD), _dp(rA), _dp(wA))


def num_threads():
    return load().orc_num_threads()


def set_num_threads(n):
    load().orc_set_num_threads(int(n))


def set_sum_mode(mode):
    """0: the reference's sequential sums (default).  1: pairwise summation of the same terms -- test-only, to measure
    the sensitivity of a residual history to re-association of the reductions."""
    load().orc_set_sum_mode(int(mode))


PATCH_TYPES = {"tractionDisplacement": 0, "fixedDisplacement": 1, "symmetryDisplacement": 2, "symmetryPlane": 3, "empty": 4,
               "fixedValue": 5}


def lame(E, nu, planeStress=False):
    lib = load()
    mu, lam = C.c_double(), C.c_double()
    lib.orc_lame(E, nu, 1 if planeStress else 0, C.byref(mu), C.byref(lam))
    return mu.value, lam.value


def _mesh_struct(mesh, geom, patchTypes):
    keep = dict(owner=_i(mesh.owner), neighbour=_i(mesh.neighbour), ps=_i([p["start"] for p in mesh.patches]),
                pz=_i([p["size"] for p in mesh.patches]), pt=_i([PATCH_TYPES[patchTypes[p["name"]]] for p in mesh.patches]),
                Sf=_f(geom.Sf), magSf=_f(geom.magSf), Cf=_f(geom.Cf), C=_f(geom.C), V=_f(geom.V), w=_f(geom.weights),
                dc=_f(geom.deltaCoeffs), fs=_i(mesh.faceStart), fp=_i(mesh.facePoints), pts=_f(mesh.points))
    M = Mesh()
    M.nCells, M.nFaces, M.nInternalFaces, M.nPatches = mesh.nCells, mesh.nFaces, mesh.nInternalFaces, len(mesh.patches)
    M.owner, M.neighbour = _ip(keep["owner"]), _ip(keep["neighbour"])
    M.patchStart, M.patchSize, M.patchType = _ip(keep["ps"]), _ip(keep["pz"]), _ip(keep["pt"])
    M.Sf, M.magSf, M.Cf, M.C, M.V = _dp(keep["Sf"]), _dp(keep["magSf"]), _dp(keep["Cf"]), _dp(keep["C"]), _dp(keep["V"])
    M.weights, M.deltaCoeffs = _dp(keep["w"]), _dp(keep["dc"])
    M.nPoints, M.faceStart, M.facePoints, M.points = mesh.nPoints, _ip(keep["fs"]), _ip(keep["fp"]), _dp(keep["pts"])
    return M, keep


def assemble_D(mesh, geom, patchTypes, mu, lam, rho, g, D, gradD, gradDf, traction, pressure, fixedValue, limitCoeff=1.0):
    lib = load()
    M, keep = _mesh_struct(mesh, geom, patchTypes)
    a = [_f(v) for v in (mu, lam, rho, g, D, gradD, gradDf, traction, pressure, fixedValue)]
    nC, nIF, nBF = mesh.nCells, mesh.nInternalFaces, mesh.nFaces - mesh.nInternalFaces
    upper = np.empty(nIF); diag = np.empty(nC); source = np.empty((nC, 3)); ic = np.empty((nBF, 3)); bc = np.empty((nBF, 3))
    pg = np.zeros((nBF, 3))
    lib.orc_assemble_D(C.byref(M), *[_dp(v) for v in a], float(limitCoeff), _dp(upper), _dp(diag), _dp(source), _dp(ic), _dp(bc), _dp(pg))
    return dict(upper=upper, diag=diag, source=source, internalCoeffs=ic, boundaryCoeffs=bc, patchGradient=pg)


def relax_residual(D, Dprev, Dold, alpha=None):
    """D.relax(alpha) (skipped when alpha is None) then unsTotalLagrangianStress::residual(); D is updated in place."""
    lib = load()
    assert D.dtype == np.float64 and D.flags.c_contiguous
    Dprev, Dold = _f(Dprev), _f(Dold)
    lib.orc_relax_residual.restype = C.c_double
    return float(lib.orc_relax_residual(C.c_int(D.shape[0]), C.c_double(alpha if alpha is not None else 1.0),
                                        C.c_int(0 if alpha is None else 1), _dp(D), _dp(Dprev), _dp(Dold)))


def sigma(mu, lam, gradD):
    lib = load()
    mu, lam, gradD = _f(mu), _f(lam), _f(gradD)
    n = mu.shape[0]
    eps = np.empty((n, 6)); sig = np.empty((n, 6))
    lib.orc_sigma(n, _dp(mu), _dp(lam), _dp(gradD), _dp(eps), _dp(sig))
    return eps, sig


def grad(mesh, geom, patchTypes, D, pointD, gradDold, fixedValue, patchGradient):
    """fvc::grad(D, pointD): returns (gradD (nC,9), boundary gradient (nBF,9))."""
    lib = load()
    M, keep = _mesh_struct(mesh, geom, patchTypes)
    a = [_f(v) for v in (D, pointD, gradDold, fixedValue, patchGradient)]
    nBF = mesh.nFaces - mesh.nInternalFaces
    G = np.empty((mesh.nCells, 9)); Gb = np.empty((nBF, 9))
    lib.orc_grad(C.byref(M), *[_dp(v) for v in a], _dp(G), _dp(Gb))
    return G, Gb


def fgrad(mesh, geom, patchTypes, D, pointD, gradD, fixedValue, patchGradient, limitCoeff=1.0):
    """fvc::fGrad(D, pointD): face gradient on every face (nF,9)."""
    lib = load()
    M, keep = _mesh_struct(mesh, geom, patchTypes)
    a = [_f(v) for v in (D, pointD, gradD, fixedValue, patchGradient)]
    Gf = np.empty((mesh.nFaces, 9))
    lib.orc_fgrad(C.byref(M), *[_dp(v) for v in a], float(limitCoeff), _dp(Gf))
    return Gf


# ---- round 2: transient / damping / non-linear / thermal branches, boundary evaluate, least-squares vol->point -------
class Controls(C.Structure):
    _fields_ = [("d2dt2Scheme", C.c_int), ("ddtScheme", C.c_int), ("deltaT", C.c_double), ("deltaT0", C.c_double),
                ("K", C.c_double), ("nonLinear", C.c_int), ("thermal", C.c_int), ("limitCoeff", C.c_double),
                ("eD2", C.c_double), ("eDdtD", C.c_double), ("eDdtD0", C.c_double), ("eDdtD00", C.c_double),
                ("eDamp", C.c_double), ("Dooo", C.c_void_p), ("Doooo", C.c_void_p)]


SCHEMES = {"steadyState": 0, "Euler": 1, "backward": 2}


def controls(d2dt2="steadyState", ddt="steadyState", deltaT=1.0, deltaT0=None, K=0.0, nonLinear=False, thermal=False,
             limitCoeff=1.0, backward=None):
    """backward: dict(eD2, eDdtD, eDdtD0, eDdtD00, eDamp, Dooo, Doooo) when a scheme is `backward` (fvm_oracle.c: orc_controls)."""
    ctl = Controls(SCHEMES[d2dt2], SCHEMES[ddt], float(deltaT), float(deltaT if deltaT0 is None else deltaT0), float(K),
                   int(bool(nonLinear)), int(bool(thermal)), float(limitCoeff))
    if backward is not None:
        for k in ("eD2", "eDdtD", "eDdtD0", "eDdtD00", "eDamp"):
            setattr(ctl, k, float(backward.get(k, 0.0)))
        ctl._keep = [_f(backward["Dooo"]), _f(backward["Doooo"])] if backward.get("Dooo") is not None else []
        if ctl._keep:
            ctl.Dooo = ctl._keep[0].ctypes.data; ctl.Doooo = ctl._keep[1].ctypes.data
    elif 2 in (ctl.d2dt2Scheme, ctl.ddtScheme):
        raise ValueError("backward schemes need the effective old time steps and the old-time levels (backward=...)")
    return ctl


def _opt(a):
    return None if a is None else _f(a)


def _odp(a):
    return None if a is None else _dp(a)


def assemble_D_ex(mesh, geom, patchTypes, ctl, mu, lam, rho, g, D, Dold, Doldold, gradD, gradDf, traction, pressure,
                  fixedValue, threeK=None, alpha=None, DT=None, DTb=None):
    lib = load()
    M, keep = _mesh_struct(mesh, geom, patchTypes)
    a = [_opt(v) for v in (mu, lam, rho, g, D, Dold, Doldold, gradD, gradDf, traction, pressure, fixedValue, threeK, alpha, DT, DTb)]
    nC, nIF, nBF = mesh.nCells, mesh.nInternalFaces, mesh.nFaces - mesh.nInternalFaces
    upper = np.empty(nIF); diag = np.empty(nC); source = np.empty((nC, 3)); ic = np.empty((nBF, 3)); bc = np.empty((nBF, 3))
    pg = np.zeros((nBF, 3))
    lib.orc_assemble_D_ex.restype = None
    lib.orc_assemble_D_ex.argtypes = [C.POINTER(Mesh), C.POINTER(Controls)] + [dp]*22
    lib.orc_assemble_D_ex(C.byref(M), C.byref(ctl), *[_odp(v) for v in a], _dp(upper), _dp(diag), _dp(source), _dp(ic),
                          _dp(bc), _dp(pg))
    return dict(upper=upper, diag=diag, source=source, internalCoeffs=ic, boundaryCoeffs=bc, patchGradient=pg)


def patch_evaluate(mesh, geom, patchTypes, D, gradD, fixedValue, patchGradient, Db=None):
    """D.correctBoundaryConditions(): boundary values (nBF,3); slots of empty faces keep what Db held (zeros if None)."""
    lib = load()
    M, keep = _mesh_struct(mesh, geom, patchTypes)
    nBF = mesh.nFaces - mesh.nInternalFaces
    out = np.zeros((nBF, 3)) if Db is None else np.array(Db, dtype=np.float64, copy=True)
    a = [_f(v) for v in (D, gradD, fixedValue, patchGradient)]
    lib.orc_patch_evaluate.restype = None
    lib.orc_patch_evaluate.argtypes = [C.POINTER(Mesh)] + [dp]*5
    lib.orc_patch_evaluate(C.byref(M), *[_dp(v) for v in a], _dp(out))
    return out


def grad_ex(mesh, geom, patchTypes, D, pointD, gradDold, fixedValue, patchGradient, gradDb=None):
    lib = load()
    M, keep = _mesh_struct(mesh, geom, patchTypes)
    a = [_f(v) for v in (D, pointD, gradDold, fixedValue, patchGradient)]
    nBF = mesh.nFaces - mesh.nInternalFaces
    G = np.empty((mesh.nCells, 9)); Gb = np.zeros((nBF, 9)) if gradDb is None else np.array(gradDb, dtype=np.float64, copy=True)
    lib.orc_grad_ex.restype = None
    lib.orc_grad_ex.argtypes = [C.POINTER(Mesh)] + [dp]*7
    lib.orc_grad_ex(C.byref(M), *[_dp(v) for v in a], _dp(G), _dp(Gb))
    return G, Gb


def fgrad_ex(mesh, geom, patchTypes, D, pointD, gradD, fixedValue, patchGradient, limitCoeff=1.0, gradDf=None):
    lib = load()
    M, keep = _mesh_struct(mesh, geom, patchTypes)
    a = [_f(v) for v in (D, pointD, gradD, fixedValue, patchGradient)]
    Gf = np.zeros((mesh.nFaces, 9)) if gradDf is None else np.array(gradDf, dtype=np.float64, copy=True)
    lib.orc_fgrad_ex.restype = None
    lib.orc_fgrad_ex.argtypes = [C.POINTER(Mesh)] + [dp]*5 + [C.c_double, dp]
    lib.orc_fgrad_ex(C.byref(M), *[_dp(v) for v in a], float(limitCoeff), _dp(Gf))
    return Gf


def sigma_ex(mu, lam, gradD, nonLinear=False, thermal=False, threeK=None, alpha=None, DT=None):
    lib = load()
    mu, lam, gradD = _f(mu), _f(lam), _f(gradD)
    n = mu.shape[0]
    eps = np.empty((n, 6)); sig = np.empty((n, 6))
    t = [_opt(v) for v in (threeK, alpha, DT)]
    lib.orc_sigma_ex.restype = None
    lib.orc_sigma_ex.argtypes = [C.c_int, C.c_int, C.c_int] + [dp]*8
    lib.orc_sigma_ex(n, int(bool(nonLinear)), int(bool(thermal)), _dp(mu), _dp(lam), _dp(gradD), *[_odp(v) for v in t], _dp(eps), _dp(sig))
    return eps, sig


def det_minmax(mesh, geom, patchTypes, gradDf):
    lib = load()
    M, keep = _mesh_struct(mesh, geom, patchTypes)
    g = _f(gradDf)
    mn, mx = C.c_double(), C.c_double()
    lib.orc_det_minmax.restype = None
    lib.orc_det_minmax.argtypes = [C.POINTER(Mesh), dp, dp, dp]
    lib.orc_det_minmax(C.byref(M), _dp(g), C.byref(mn), C.byref(mx))
    return mn.value, mx.value


def ddt(D, Dold, scheme, deltaT):
    lib = load()
    D, Dold = _f(D), _f(Dold)
    U = np.empty_like(D)
    lib.orc_ddt.restype = None
    lib.orc_ddt.argtypes = [C.c_int, C.c_int, C.c_double, dp, dp, dp]
    lib.orc_ddt(D.shape[0], {"steadyState": 0, "Euler": 1}[scheme], float(deltaT), _dp(D), _dp(Dold), _dp(U))
    return U


def ddt_backward(D, Dold, Doldold, deltaT, e):
    lib = load()
    D, Dold, Doldold = _f(D), _f(Dold), _f(Doldold)
    U = np.empty_like(D)
    lib.orc_ddt_backward.restype = None
    lib.orc_ddt_backward.argtypes = [C.c_int, C.c_double, C.c_double, dp, dp, dp, dp]
    lib.orc_ddt_backward(D.shape[0], float(deltaT), float(e), _dp(D), _dp(Dold), _dp(Doldold), _dp(U))
    return U


def relax_boundary(Db, DbPrev, alpha):
    lib = load()
    assert Db.dtype == np.float64 and Db.flags.c_contiguous
    DbPrev = _f(DbPrev)
    lib.orc_relax_boundary.restype = None
    lib.orc_relax_boundary.argtypes = [C.c_int, C.c_double, dp, dp]
    lib.orc_relax_boundary(Db.size, float(alpha), _dp(Db), _dp(DbPrev))


class Lsq(object):
    """leastSquaresVolPointInterpolation of one mesh (stencils, origins, inverse least-squares matrices)."""

    def __init__(self, mesh, geom, patchTypes):
        lib = load()
        self._M, self._keep = _mesh_struct(mesh, geom, patchTypes)
        lib.orc_lsq_build.restype = C.c_void_p
        lib.orc_lsq_build.argtypes = [C.POINTER(Mesh)]
        self._h = C.c_void_p(lib.orc_lsq_build(C.byref(self._M)))
        self.nPoints = mesh.nPoints

    def export(self):
        lib = load()
        lib.orc_lsq_rows.restype = C.c_int
        lib.orc_lsq_rows.argtypes = [C.c_void_p]
        rows = lib.orc_lsq_rows(self._h)
        start = np.empty(self.nPoints + 1, dtype=np.int32); entry = np.empty(max(rows, 1), dtype=np.int32)
        inv = np.empty((max(rows, 1), 3)); origin = np.empty((self.nPoints, 3)); mirror = np.empty(self.nPoints, dtype=np.uint8)
        lib.orc_lsq_export.restype = None
        lib.orc_lsq_export.argtypes = [C.c_void_p, ip, ip, dp, dp, C.POINTER(C.c_ubyte)]
        lib.orc_lsq_export(self._h, _ip(start), _ip(entry), _dp(inv), _dp(origin), mirror.ctypes.data_as(C.POINTER(C.c_ubyte)))
        return dict(start=start, entry=entry[:rows], inv=inv[:rows], origin=origin, mirror=mirror)

    def interpolate(self, D, Db):
        lib = load()
        D, Db = _f(D), _f(Db)
        out = np.empty((self.nPoints, 3))
        lib.orc_lsq_interpolate.restype = None
        lib.orc_lsq_interpolate.argtypes = [C.c_void_p, C.POINTER(Mesh), dp, dp, dp]
        lib.orc_lsq_interpolate(self._h, C.byref(self._M), _dp(D), _dp(Db), _dp(out))
        return out

    def __del__(self):
        try:
            lib = load()
            lib.orc_lsq_free.argtypes = [C.c_void_p]
            lib.orc_lsq_free(self._h)
        except Exception:
            pass


def point_constraints_build(mesh, geom, pointPatchTypes, pointPatchValues=None):
    """Plain-loop restatement of which points pf.correctBoundaryConditions() touches and with what (see orc_point_constraints):
    per patch in boundary order, meshPoints() in order of first appearance over the patch faces; symmetryPlane: pointNormals()."""
    rows = {}
    nIF = mesh.nInternalFaces
    for p in mesh.patches:
        t = pointPatchTypes.get(p["name"], "calculated")
        if t not in ("symmetryPlane", "fixedValue"):
            continue
        order, acc = [], {}
        for f in range(p["start"], p["start"] + p["size"]):
            n = geom.Sf[f]/(geom.magSf[f] + 1.0e-300)
            for q in mesh.facePoints[mesh.faceStart[f]:mesh.faceStart[f + 1]]:
                q = int(q)
                if q not in acc:
                    acc[q] = np.zeros(3); order.append(q)
                acc[q] = acc[q] + n
        for q in order:
            if t == "fixedValue":
                v = np.zeros(3) if not pointPatchValues or p["name"] not in pointPatchValues else np.asarray(pointPatchValues[p["name"]], dtype=np.float64)
                rows.setdefault(q, []).append((2, v))
            else:
                a = acc[q]
                rows.setdefault(q, []).append((1, a/(np.sqrt(a[0]*a[0] + a[1]*a[1] + a[2]*a[2]) + 1.0e-300)))
    pts = sorted(rows)
    start = [0]; kind = []; vec = []
    for q in pts:
        for k, v in rows[q]:
            kind.append(k); vec.append(v)
        start.append(len(kind))
    return dict(point=np.array(pts, dtype=np.int32), start=np.array(start, dtype=np.int32), kind=np.array(kind, dtype=np.int32),
                vec=np.array(vec, dtype=np.float64).reshape(-1, 3))


def point_constraints_apply(pc, pointD):
    """pf.correctBoundaryConditions() on pointD (n,3), in place."""
    lib = load()
    assert pointD.dtype == np.float64 and pointD.flags.c_contiguous
    point, start, kind, vec = _i(pc["point"]), _i(pc["start"]), _i(pc["kind"]), _f(pc["vec"])
    lib.orc_point_constraints.restype = None
    lib.orc_point_constraints.argtypes = [C.c_int, ip, ip, ip, dp, dp]
    lib.orc_point_constraints(point.shape[0], _ip(point), _ip(start), _ip(kind), _dp(vec), _dp(pointD))
    return pointD


def assemble_scalar(mesh, geom, patchTypes, sign, bfKind, bfValue, gammaCells=None, gammaFaces=None, ddt="steadyState", deltaT=1.0,
                    rate=None, psiOld=None, phi=None, refCell=-1, refValue=0.0, psi=None, ls=None, limitCoeff=1.0):
    """orc_assemble_scalar: TEqn / pEqn forms (see fvm_oracle.c).  Returns dict(upper, diag, source, internalCoeffs, boundaryCoeffs).
    psi + ls = (lsP, lsN, lsB): with the explicit skewCorrected correction (least-squares gradient of psi)."""
    if psi is not None:
        return _assemble_scalar_corrected(mesh, geom, patchTypes, sign, bfKind, bfValue, gammaCells, gammaFaces, ddt, deltaT, rate,
                                          psiOld, phi, refCell, refValue, psi, ls, limitCoeff)
    lib = load()
    M, keep = _mesh_struct(mesh, geom, patchTypes)
    nC, nIF, nBF = mesh.nCells, mesh.nInternalFaces, mesh.nFaces - mesh.nInternalFaces
    upper = np.empty(nIF); diag = np.empty(nC); source = np.empty(nC); ic = np.empty(max(nBF, 1)); bc = np.empty(max(nBF, 1))
    kind = _i(bfKind); val = _f(bfValue)
    a = [_opt(v) for v in (gammaCells, gammaFaces)]
    b = [_opt(v) for v in (rate, psiOld, phi)]
    lib.orc_assemble_scalar.restype = None
    lib.orc_assemble_scalar.argtypes = [C.POINTER(Mesh), C.c_int, dp, dp, ip, dp, C.c_int, C.c_double, dp, dp, dp, C.c_int, C.c_double,
                                        dp, dp, dp, dp, dp]
    lib.orc_assemble_scalar(C.byref(M), int(sign), _odp(a[0]), _odp(a[1]), _ip(kind), _dp(val), 1 if ddt == "Euler" else 0, float(deltaT),
                            _odp(b[0]), _odp(b[1]), _odp(b[2]), int(refCell), float(refValue), _dp(upper), _dp(diag), _dp(source), _dp(ic), _dp(bc))
    return dict(upper=upper, diag=diag, source=source, internalCoeffs=ic[:nBF], boundaryCoeffs=bc[:nBF])


def _assemble_scalar_corrected(mesh, geom, patchTypes, sign, bfKind, bfValue, gammaCells, gammaFaces, ddt, deltaT, rate, psiOld, phi,
                               refCell, refValue, psi, ls, limitCoeff):
    lib = load()
    M, keep = _mesh_struct(mesh, geom, patchTypes)
    nC, nIF, nBF = mesh.nCells, mesh.nInternalFaces, mesh.nFaces - mesh.nInternalFaces
    upper = np.empty(nIF); diag = np.empty(nC); source = np.empty(nC); ic = np.empty(max(nBF, 1)); bc = np.empty(max(nBF, 1))
    kind = _i(bfKind); val = _f(bfValue)
    a = [_opt(v) for v in (gammaCells, gammaFaces)]
    b = [_opt(v) for v in (rate, psiOld, phi)]
    c = [_f(psi), _f(ls[0]), _f(ls[1]), _f(ls[2])]
    lib.orc_assemble_scalar_corrected.restype = None
    lib.orc_assemble_scalar_corrected.argtypes = [C.POINTER(Mesh), C.c_int, dp, dp, ip, dp, C.c_int, C.c_double, dp, dp, dp, C.c_int,
                                                  C.c_double, dp, dp, dp, dp, C.c_double, dp, dp, dp, dp, dp]
    lib.orc_assemble_scalar_corrected(C.byref(M), int(sign), _odp(a[0]), _odp(a[1]), _ip(kind), _dp(val), 1 if ddt == "Euler" else 0,
                                      float(deltaT), _odp(b[0]), _odp(b[1]), _odp(b[2]), int(refCell), float(refValue),
                                      _dp(c[0]), _dp(c[1]), _dp(c[2]), _dp(c[3]), float(limitCoeff),
                                      _dp(upper), _dp(diag), _dp(source), _dp(ic), _dp(bc))
    return dict(upper=upper, diag=diag, source=source, internalCoeffs=ic[:nBF], boundaryCoeffs=bc[:nBF])
