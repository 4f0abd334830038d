"""ctypes binding of include/gpupcg.h (libgpupcg_cuda.so).

This is a thin test/benchmark driver over the C ABI -- the product is the shared library, the
foam-extend adaptor (fluidstructureinteraction_b200/foam) is its production caller.  There is no
fallback of any kind here: if the CUDA library is missing or no B200 is visible, calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GPUPCG_LIBRARY") or os.path.join(_HERE, "libgpupcg_cuda.so")     # GPUPCG_LIBRARY: A/B builds (tools/)

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class GpuPcgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("gpupcg error %d: %s" % (code, msg))
        self.code = code


class Perf(C.Structure):
    _fields_ = [
        ("initialResidual", C.c_double),
        ("finalResidual", C.c_double),
        ("normFactor", C.c_double),
        ("nIterations", C.c_int),
        ("converged", C.c_int),
        ("singular", C.c_int),
        ("reserved", C.c_int),
        ("solve_ms", C.c_double),
        ("h2d_ms", C.c_double),
        ("d2h_ms", C.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class EqnControls(C.Structure):
    """gpupcg_eqn_controls"""
    _fields_ = [("d2dt2Scheme", C.c_int), ("ddtScheme", C.c_int), ("deltaT", C.c_double), ("deltaT0", C.c_double),
                ("K", C.c_double), ("nonLinear", C.c_int), ("thermal", C.c_int), ("limitCoeff", C.c_double),
                ("eD2", C.c_double), ("eDdtD", C.c_double), ("eDdtD0", C.c_double), ("eDdtD00", C.c_double),
                ("eDamp", C.c_double), ("eU", C.c_double)]


class EvolveControls(C.Structure):
    """gpupcg_evolve_controls"""
    _fields_ = [("nCorrectors", C.c_int), ("convergenceTolerance", C.c_double), ("relConvergenceTolerance", C.c_double),
                ("tolerance", C.c_double), ("relTol", C.c_double), ("minIter", C.c_int), ("maxIter", C.c_int),
                ("relaxD", C.c_double), ("doRelax", C.c_int), ("validCmpt", C.c_int*3), ("g", C.c_double*3),
                ("eqn", EqnControls), ("incremental", C.c_int)]


class EvolveResult(C.Structure):
    """gpupcg_evolve_result"""
    _fields_ = [("nOuterIterations", C.c_int), ("totalPcgIterations", C.c_int), ("enforceLinear", C.c_int),
                ("converged", C.c_int), ("initialResidual", C.c_double), ("lastSolverInitialResidual", C.c_double),
                ("maxRes", C.c_double), ("res", C.c_double), ("assemble_ms", C.c_double), ("solve_ms", C.c_double),
                ("update_ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


SCHEMES = {"steadyState": 0, "Euler": 1, "backward": 2}


def eqn_controls(d2dt2="steadyState", ddt="steadyState", deltaT=1.0, deltaT0=None, K=0.0, nonLinear=False, thermal=False,
                 limitCoeff=1.0, backward=None):
    """backward: dict(eD2, eDdtD, eDdtD0, eDdtD00, eDamp, eU) -- the effective old time step of every use of a `backward`
    scheme (include/gpupcg.h: gpupcg_eqn_controls; timelevels.OldTimes works them out)."""
    c = EqnControls(SCHEMES[d2dt2], SCHEMES[ddt], float(deltaT), float(deltaT if deltaT0 is None else deltaT0), float(K),
                    int(bool(nonLinear)), int(bool(thermal)), float(limitCoeff))
    for k, v in (backward or {}).items():
        setattr(c, k, float(v))
    return c


class PlanInfo(C.Structure):
    _fields_ = [
        ("nCells", C.c_int), ("nFaces", C.c_int), ("nRows", C.c_int), ("nSlices", C.c_int),
        ("nLevels", C.c_int), ("maxLevelRows", C.c_int),
        ("entriesUpper", C.c_longlong), ("entriesLower", C.c_longlong),
        ("nInterfaceFaces", C.c_int), ("nInterfaces", C.c_int),
        ("sweepGrid", C.c_int), ("streamGrid", C.c_int), ("numSMs", C.c_int),
        ("nTiles", C.c_int), ("tileRowsMax", C.c_int), ("maxRounds", C.c_int), ("sweepSmemBytes", C.c_int),
        ("extForward", C.c_int), ("extBackward", C.c_int), ("roundEntries", C.c_int),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


# every symbol include/gpupcg.h declares: (restype, argtypes)
SYMBOLS = {
    "gpupcg_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, c_int_p, c_int_p,
                                C.c_int, c_int_p, C.POINTER(c_int_p), c_int_p]),
    "gpupcg_destroy": (C.c_int, [C.c_void_p]),
    "gpupcg_last_error": (C.c_char_p, [C.c_void_p]),
    "gpupcg_get_plan_info": (C.c_int, [C.c_void_p, C.POINTER(PlanInfo)]),
    "gpupcg_debug_tile_trace": (C.c_int, [C.c_void_p, C.POINTER(C.c_ulonglong)]),
    "gpupcg_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_longlong)]),
    "gpupcg_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gpupcg_set_matrix": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, C.POINTER(c_double_p)]),
    "gpupcg_set_matrix_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpupcg_solve": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_double, C.c_double, C.c_int, C.c_int,
                               C.POINTER(Perf), c_double_p, C.c_int]),
    "gpupcg_solve_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int,
                                      C.c_int, C.POINTER(Perf), c_double_p, C.c_int]),
    "gpupcg_amul": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p]),
    "gpupcg_dic_factor": (C.c_int, [C.c_void_p, c_double_p]),
    "gpupcg_precondition": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    "gpupcg_time_kernel": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p]),
    "gpupcg_vector_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "gpupcg_set_matrix_vector": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, C.POINTER(c_double_p)]),
    "gpupcg_set_matrix_vector_device": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpupcg_solve_vector": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, C.c_double, C.c_double, C.c_int, C.c_int,
                                      C.POINTER(Perf), c_double_p, C.c_int]),
    "gpupcg_solve_vector_device": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int,
                                             C.c_int, C.POINTER(Perf), c_double_p, C.c_int]),
    "gpupcg_time_kernel_vector": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p]),
    "gpupcg_comm_unique_id": (C.c_int, [C.c_void_p]),
    "gpupcg_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "gpupcg_mesh_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, c_int_p, c_int_p, C.c_int,
                                     c_int_p, c_int_p, c_int_p] + [c_double_p]*7),
    "gpupcg_mesh_destroy": (C.c_int, [C.c_void_p]),
    "gpupcg_mesh_last_error": (C.c_char_p, [C.c_void_p]),
    "gpupcg_assemble_D": (C.c_int, [C.c_void_p] + [c_double_p]*10 + [C.c_double, C.c_int] + [c_double_p]*6),
    "gpupcg_sigma": (C.c_int, [C.c_void_p] + [c_double_p]*5),
    "gpupcg_patch_gradient": (C.c_int, [C.c_void_p, c_double_p]),
    "gpupcg_relax_residual": (C.c_int, [C.c_void_p, C.c_double, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]),
    "gpupcg_mesh_set_points": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_int_p, c_int_p]),
    "gpupcg_mesh_fetch": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "gpupcg_solve_assembled_D": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, C.c_double, C.c_double, C.c_int, C.c_int,
                                           C.POINTER(Perf)]),
    "gpupcg_gradients": (C.c_int, [C.c_void_p] + [c_double_p]*5 + [C.c_double, C.c_int] + [c_double_p]*4),
    "gpupcg_mesh_field_size": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_longlong)]),
    "gpupcg_mesh_set_field": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "gpupcg_mesh_get_field": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "gpupcg_assemble_scalar": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                        C.POINTER(C.c_double)]),
    "gpupcg_mesh_set_lsq": (C.c_int, [C.c_void_p, c_int_p, c_int_p, c_double_p, c_double_p, C.POINTER(C.c_ubyte)]),
    "gpupcg_mesh_set_point_constraints": (C.c_int, [C.c_void_p, C.c_int, c_int_p, c_int_p, c_int_p, c_double_p]),
    "gpupcg_vol_to_point": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "gpupcg_patch_evaluate": (C.c_int, [C.c_void_p]),
    "gpupcg_assemble_D_resident": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p]),
    "gpupcg_gradients_resident": (C.c_int, [C.c_void_p, C.c_double]),
    "gpupcg_sigma_resident": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gpupcg_mesh_new_time_step": (C.c_int, [C.c_void_p]),
    "gpupcg_mesh_copy_field": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "gpupcg_evolve_D": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_double_p, c_int_p, C.c_int]),
    "gpupcg_host_register": (C.c_int, [C.c_void_p, C.c_ulonglong]),
    "gpupcg_host_unregister": (C.c_int, [C.c_void_p]),
    "gpupcg_plan_export": (C.c_int, [C.c_int, C.c_int, c_int_p, c_int_p, C.c_int, C.c_int, C.POINTER(PlanInfo),
                                     C.POINTER(c_int_p)]),
    "gpupcg_device_count": (C.c_int, []),
    "gpupcg_device_report": (C.c_int, [C.c_int, C.c_char_p, C.c_int]),
}

_lib = None


def load_library(path=None):
    """dlopen libgpupcg_cuda.so and type every exported symbol.  Raises if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise FileNotFoundError(
            "%s is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback." % p)
    lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)     # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _dp(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(c_int_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class GpuPCG(object):
    """One handle = one lduAddressing on one GPU (mirror of the lifetime foam-extend gives a mesh)."""

    def __init__(self, lowerAddr, upperAddr, nCells, patchFaceCells=(), patchNbrRank=(), device=0):
        self.lib = load_library()
        self.l = _i32(lowerAddr)
        self.u = _i32(upperAddr)
        self.nCells = int(nCells)
        self.nFaces = int(self.l.shape[0])
        self.patchFaceCells = [_i32(fc) for fc in patchFaceCells]
        nP = len(self.patchFaceCells)
        sizes = _i32([fc.shape[0] for fc in self.patchFaceCells]) if nP else None
        ptrs = (c_int_p * max(nP, 1))(*[_ip(fc) for fc in self.patchFaceCells]) if nP else None
        nbr = _i32(list(patchNbrRank)) if nP else None
        self.h = C.c_void_p()
        rc = self.lib.gpupcg_create(C.byref(self.h), device, self.nCells, self.nFaces, _ip(self.l), _ip(self.u),
                                    nP, _ip(sizes), ptrs, _ip(nbr))
        if rc != 0:
            msg = self.lib.gpupcg_last_error(None)
            self.h = None
            raise GpuPcgError(rc, msg.decode() if msg else "?")

    def _check(self, rc):
        if rc != 0:
            msg = self.lib.gpupcg_last_error(self.h)
            raise GpuPcgError(rc, msg.decode() if msg else "?")

    def close(self):
        if getattr(self, "h", None):
            self.lib.gpupcg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def plan_info(self):
        info = PlanInfo()
        self._check(self.lib.gpupcg_get_plan_info(self.h, C.byref(info)))
        return info.as_dict()

    def tile_trace(self):
        n = self.plan_info()["nTiles"]
        a = np.zeros(3*n, dtype=np.uint64)
        self._check(self.lib.gpupcg_debug_tile_trace(self.h, a.ctypes.data_as(C.POINTER(C.c_ulonglong))))
        return a.reshape(n, 3)

    def launch_count(self):
        n = C.c_longlong(0)
        self._check(self.lib.gpupcg_launch_count(self.h, C.byref(n)))
        return n.value

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.gpupcg_set_stream(self.h, C.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))

    def set_matrix(self, diag, upper=None, lower=None, patchBouCoeffs=()):
        diag = _f64(diag)
        upper = _f64(upper) if upper is not None else None
        lower_p = None
        if lower is not None:
            lower = _f64(lower)
            lower_p = _dp(upper) if lower is upper else _dp(lower)
        bc = [_f64(b) for b in patchBouCoeffs]
        ptrs = (c_double_p * max(len(bc), 1))(*[_dp(b) for b in bc]) if bc else None
        self._check(self.lib.gpupcg_set_matrix(self.h, _dp(diag), _dp(upper), lower_p, ptrs))

    def set_matrix_device(self, d_diag, d_upper=None, d_bou=None):
        self._check(self.lib.gpupcg_set_matrix_device(self.h, C.c_void_p(d_diag),
                                                      C.c_void_p(d_upper) if d_upper else None,
                                                      C.c_void_p(d_bou) if d_bou else None))

    def solve(self, x, b, tolerance=1e-6, relTol=0.0, minIter=0, maxIter=1000, history=True):
        """x: float64 array, updated in place (must be C-contiguous float64).  Returns (perf dict, history)."""
        if not (isinstance(x, np.ndarray) and x.dtype == np.float64 and x.flags.c_contiguous):
            raise TypeError("x must be a C-contiguous float64 ndarray (it is the in/out psi component)")
        b = _f64(b)
        perf = Perf()
        hist = np.zeros(max(maxIter, 1), dtype=np.float64) if history else None
        self._check(self.lib.gpupcg_solve(self.h, _dp(x), _dp(b), tolerance, relTol, minIter, maxIter,
                                          C.byref(perf), _dp(hist), hist.shape[0] if history else 0))
        d = perf.as_dict()
        return d, (hist[:min(d["nIterations"], hist.shape[0])].copy() if history else None)

    def solve_device(self, d_x, d_b, tolerance=1e-6, relTol=0.0, minIter=0, maxIter=1000, history=False):
        perf = Perf()
        hist = np.zeros(max(maxIter, 1), dtype=np.float64) if history else None
        self._check(self.lib.gpupcg_solve_device(self.h, C.c_void_p(d_x), C.c_void_p(d_b), tolerance, relTol,
                                                 minIter, maxIter, C.byref(perf), _dp(hist),
                                                 hist.shape[0] if history else 0))
        d = perf.as_dict()
        return d, (hist[:min(d["nIterations"], hist.shape[0])].copy() if history else None)

    def amul(self, x, xNbr=None):
        x = _f64(x)
        Ax = np.empty(self.nCells, dtype=np.float64)
        xn = _f64(xNbr) if xNbr is not None else None
        self._check(self.lib.gpupcg_amul(self.h, _dp(x), _dp(Ax), _dp(xn)))
        return Ax

    def dic_factor(self):
        rD = np.empty(self.nCells, dtype=np.float64)
        self._check(self.lib.gpupcg_dic_factor(self.h, _dp(rD)))
        return rD

    def precondition(self, rA):
        rA = _f64(rA)
        wA = np.empty(self.nCells, dtype=np.float64)
        self._check(self.lib.gpupcg_precondition(self.h, _dp(rA), _dp(wA)))
        return wA

    KERNELS = {"amul": 0, "dic": 1, "p_update": 2, "xr_update": 3, "dic_factor": 4}

    def time_kernel(self, which, reps=10, flushL2=True):
        ms = C.c_double(0.0)
        self._check(self.lib.gpupcg_time_kernel(self.h, self.KERNELS[which], reps, 1 if flushL2 else 0, C.byref(ms)))
        return ms.value

    # ---- component-batched solve (fvMatrix<vector>::solve's component loop in one launch sequence)
    def vector_mode(self, nCmpt=3):
        """1: the batched kernels run; 0: the components are solved one after the other on the scalar kernels; 2: as
        concurrent scalar solves (one stream per component inside the library)."""
        rc = self.lib.gpupcg_vector_mode(self.h, nCmpt)
        if rc < 0:
            self._check(rc)
        return rc

    def set_matrix_vector(self, diagCmpt, upper=None, patchBouCoeffsCmpt=()):
        """diagCmpt: (nCells, nCmpt) = diag + internalCoeffs per component; upper shared (None keeps the previous)."""
        diagCmpt = _f64(diagCmpt)
        assert diagCmpt.ndim == 2 and diagCmpt.shape[0] == self.nCells
        upper = _f64(upper) if upper is not None else None
        bc = [_f64(b) for b in patchBouCoeffsCmpt]
        ptrs = (c_double_p * max(len(bc), 1))(*[_dp(b) for b in bc]) if bc else None
        self._check(self.lib.gpupcg_set_matrix_vector(self.h, diagCmpt.shape[1], _dp(diagCmpt), _dp(upper), ptrs))

    def set_matrix_vector_device(self, nCmpt, d_diagCmpt, d_upper=None, d_bou=None):
        self._check(self.lib.gpupcg_set_matrix_vector_device(self.h, nCmpt, C.c_void_p(d_diagCmpt),
                                                             C.c_void_p(d_upper) if d_upper else None,
                                                             C.c_void_p(d_bou) if d_bou else None))

    def _vec_result(self, perf, hist, nCmpt, history):
        ds = [perf[k].as_dict() for k in range(nCmpt)]
        hs = [hist[k, :min(ds[k]["nIterations"], hist.shape[1])].copy() for k in range(nCmpt)] if history else None
        return ds, hs

    def solve_vector(self, x, b, tolerance=1e-6, relTol=0.0, minIter=0, maxIter=1000, history=True):
        """x: (nCells, nCmpt) float64 C-contiguous, in/out.  Returns ([perf dict per component], [history per component])."""
        if not (isinstance(x, np.ndarray) and x.dtype == np.float64 and x.flags.c_contiguous and x.ndim == 2):
            raise TypeError("x must be a C-contiguous float64 (nCells, nCmpt) ndarray (psi.internalField(), in/out)")
        nCmpt = x.shape[1]
        b = _f64(b)
        assert b.shape == x.shape
        perf = (Perf*nCmpt)()
        hist = np.zeros((nCmpt, max(maxIter, 1)), dtype=np.float64) if history else None
        self._check(self.lib.gpupcg_solve_vector(self.h, nCmpt, _dp(x), _dp(b), tolerance, relTol, minIter, maxIter, perf,
                                                 _dp(hist), hist.shape[1] if history else 0))
        return self._vec_result(perf, hist, nCmpt, history)

    def solve_vector_device(self, nCmpt, d_x, d_b, tolerance=1e-6, relTol=0.0, minIter=0, maxIter=1000, history=False):
        perf = (Perf*nCmpt)()
        hist = np.zeros((nCmpt, max(maxIter, 1)), dtype=np.float64) if history else None
        self._check(self.lib.gpupcg_solve_vector_device(self.h, nCmpt, C.c_void_p(d_x), C.c_void_p(d_b), tolerance, relTol,
                                                        minIter, maxIter, perf, _dp(hist), hist.shape[1] if history else 0))
        return self._vec_result(perf, hist, nCmpt, history)

    def solve_assembled_D(self, gpuMesh, D, tolerance=1e-6, relTol=0.0, minIter=0, maxIter=1000):
        """DEqn.solve() on the D equation gpuMesh.assemble_D left on the device (the matrix never crosses PCIe).
        D: (nCells, 3) float64 C-contiguous, in/out.  Returns [perf dict per component]."""
        assert D.dtype == np.float64 and D.flags.c_contiguous and D.shape == (self.nCells, 3)
        perf = (Perf*3)()
        self._check(self.lib.gpupcg_solve_assembled_D(self.h, gpuMesh.h, _dp(D), tolerance, relTol, minIter, maxIter, perf))
        return [perf[k].as_dict() for k in range(3)]

    def evolve_D(self, gpuMesh, nCorrectors, convergenceTolerance, relConvergenceTolerance=0.0, tolerance=1e-9, relTol=0.01,
                 minIter=0, maxIter=1000, relaxD=None, validCmpts=(1, 1, 1), g=(0.0, 0.0, 0.0), eqn=None, historyCap=None,
                 incremental=False):
        """unsTotalLagrangianStress::evolve() on the fields resident in gpuMesh (gpupcg_evolve_D); incremental: the loop of
        unsIncrTotalLagrangianStress::evolve() (the resident D is the increment DD)."""
        ec = EvolveControls()
        ec.incremental = 1 if incremental else 0
        ec.nCorrectors = int(nCorrectors); ec.convergenceTolerance = convergenceTolerance
        ec.relConvergenceTolerance = relConvergenceTolerance; ec.tolerance = tolerance; ec.relTol = relTol
        ec.minIter = minIter; ec.maxIter = maxIter
        ec.relaxD = 1.0 if relaxD is None else float(relaxD); ec.doRelax = 0 if relaxD is None else 1
        for k in range(3):
            ec.validCmpt[k] = int(validCmpts[k]); ec.g[k] = float(g[k])
        ec.eqn = eqn if eqn is not None else eqn_controls()
        cap = int(historyCap if historyCap is not None else nCorrectors)
        hist = np.zeros(max(cap, 1)); its = np.zeros((max(cap, 1), 3), dtype=np.int32)
        res = EvolveResult()
        self._check(self.lib.gpupcg_evolve_D(self.h, gpuMesh.h, C.byref(ec), C.byref(res), _dp(hist), _ip(its), cap))
        r = res.as_dict()
        n = min(cap, r["nOuterIterations"] + 1)
        r["residualHistory"] = hist[:n].copy()
        nv = int(sum(1 for v in validCmpts if v))
        r["pcgIterations"] = its[:n, :nv].copy()
        return r

    def time_kernel_vector(self, nCmpt, which, reps=10, flushL2=True):
        ms = C.c_double(0.0)
        self._check(self.lib.gpupcg_time_kernel_vector(self.h, nCmpt, self.KERNELS[which], reps, 1 if flushL2 else 0, C.byref(ms)))
        return ms.value

    def comm_init(self, rank, nRanks, unique_id_bytes):
        buf = C.create_string_buffer(bytes(unique_id_bytes), 128)
        self._check(self.lib.gpupcg_comm_init(self.h, rank, nRanks, buf))


PATCH_TYPES = {"tractionDisplacement": 0, "fixedDisplacement": 1, "symmetryDisplacement": 2, "symmetryPlane": 3, "empty": 4,
               "fixedValue": 5}

FIELD_IDS = dict(mu=0, lambda_=1, rho=2, D=3, Dold=4, Doldold=5, gradD=6, gradDf=7, traction=8, pressure=9, fixedValue=10,
                 threeK=11, alpha=12, DT=13, DTb=14, pointD=15, Db=16, gradDb=17, U=18, epsilon=19, sigma=20, upper=21,
                 diag=22, source=23, internalCoeffs=24, boundaryCoeffs=25, patchGradient=26, Dprev=27, DbPrev=28, Dooo=29,
                 Doooo=30, gradDfBStart=31)
_FIELD_COLS = dict(mu=1, lambda_=1, rho=1, D=3, Dold=3, Doldold=3, gradD=9, gradDf=9, traction=3, pressure=1, fixedValue=3,
                   threeK=1, alpha=1, DT=1, DTb=1, pointD=3, Db=3, gradDb=9, U=3, epsilon=6, sigma=6, upper=1, diag=1,
                   source=3, internalCoeffs=3, boundaryCoeffs=3, patchGradient=3, Dprev=3, DbPrev=3, Dooo=3, Doooo=3,
                   gradDfBStart=9)


class GpuMesh(object):
    """Device copy of what foam-extend's fvMesh provides (addressing + geometry) for the assembly kernels."""

    def __init__(self, mesh, geom, patchTypes, device=0):
        self.lib = load_library()
        self.mesh, self.geom = mesh, geom
        self.nC, self.nF, self.nIF = mesh.nCells, mesh.nFaces, mesh.nInternalFaces
        self.nBF = self.nF - self.nIF
        ps = _i32([p["start"] for p in mesh.patches]); pz = _i32([p["size"] for p in mesh.patches])
        pt = _i32([PATCH_TYPES[patchTypes[p["name"]]] for p in mesh.patches])
        self._keep = [_f64(geom.Sf), _f64(geom.magSf), _f64(geom.Cf), _f64(geom.C), _f64(geom.V), _f64(geom.weights),
                      _f64(geom.deltaCoeffs)]
        self.h = C.c_void_p()
        rc = self.lib.gpupcg_mesh_create(C.byref(self.h), device, self.nC, self.nF, self.nIF, _ip(_i32(mesh.owner)),
                                         _ip(_i32(mesh.neighbour)), len(mesh.patches), _ip(ps), _ip(pz), _ip(pt),
                                         *[_dp(a) for a in self._keep])
        if rc != 0:
            msg = self.lib.gpupcg_mesh_last_error(None)
            self.h = None
            raise GpuPcgError(rc, msg.decode() if msg else "?")

    def _check(self, rc):
        if rc != 0:
            raise GpuPcgError(rc, self.lib.gpupcg_mesh_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.gpupcg_mesh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    FIELDS = {"D": 0, "gradD": 1, "gradDf": 2, "gradDb": 3}

    def fetch(self, which):
        """A device-resident field on demand (gpupcg_mesh_fetch)."""
        shape = {"D": (self.nC, 3), "gradD": (self.nC, 9), "gradDf": (self.nF, 9), "gradDb": (self.nBF, 9)}[which]
        out = np.empty(shape)
        self._check(self.lib.gpupcg_mesh_fetch(self.h, self.FIELDS[which], _dp(out)))
        return out

    def assemble_D(self, mu, lam, rho, g, D, gradD, gradDf, traction, pressure, fixedValue, limitCoeff=1.0, reps=1, copy_back=True):
        """gradD / gradDf = None: the gradients the device holds (left there by gradients())."""
        a = [_f64(v) if v is not None else None for v in (mu, lam, rho, g, D, gradD, gradDf, traction, pressure, fixedValue)]
        ms = C.c_double(0.0)
        if not copy_back:       # the equation stays on the device (solve_assembled_D)
            self._check(self.lib.gpupcg_assemble_D(self.h, *[_dp(v) for v in a], float(limitCoeff), int(reps), None, None, None,
                                                   None, None, C.byref(ms)))
            return dict(kernel_ms=ms.value)
        upper = np.empty(self.nIF); diag = np.empty(self.nC); source = np.empty((self.nC, 3))
        ic = np.empty((self.nBF, 3)); bc = np.empty((self.nBF, 3))
        self._check(self.lib.gpupcg_assemble_D(self.h, *[_dp(v) for v in a], float(limitCoeff), int(reps), _dp(upper),
                                               _dp(diag), _dp(source), _dp(ic), _dp(bc), C.byref(ms)))
        return dict(upper=upper, diag=diag, source=source, internalCoeffs=ic, boundaryCoeffs=bc, kernel_ms=ms.value)

    def gradients_resident(self, D, pointD, fixedValue, patchGradient=None, limitCoeff=1.0):
        """fvc::grad + fvc::fGrad with the device's current grad(D) as the registered gradient; results stay on the
        device (fetch() them when needed)."""
        self._ensure_points()
        a = [_f64(v) if v is not None else None for v in (D, pointD, None, fixedValue, patchGradient)]
        self._check(self.lib.gpupcg_gradients(self.h, *[_dp(v) for v in a], float(limitCoeff), 1, None, None, None, None))

    def _ensure_points(self):
        if not getattr(self, "_points", None):
            self._points = (_f64(self.mesh.points), _i32(self.mesh.faceStart), _i32(self.mesh.facePoints))
            self._check(self.lib.gpupcg_mesh_set_points(self.h, self.mesh.nPoints, _dp(self._points[0]),
                                                        _ip(self._points[1]), _ip(self._points[2])))

    def gradients(self, D, pointD, gradDold, fixedValue, patchGradient, limitCoeff=1.0, reps=1):
        self._ensure_points()
        a = [_f64(v) if v is not None else None for v in (D, pointD, gradDold, fixedValue, patchGradient)]
        G = np.empty((self.nC, 9)); Gb = np.empty((self.nBF, 9)); Gf = np.empty((self.nF, 9))
        ms = C.c_double(0.0)
        self._check(self.lib.gpupcg_gradients(self.h, *[_dp(v) for v in a], float(limitCoeff), int(reps), _dp(G), _dp(Gb),
                                              _dp(Gf), C.byref(ms)))
        return G, Gb, Gf, ms.value

    def patch_gradient(self):
        """gradient() of the tractionDisplacement patches as the last assemble_D left it on the device."""
        pg = np.empty((self.nBF, 3))
        self._check(self.lib.gpupcg_patch_gradient(self.h, _dp(pg)))
        return pg

    def relax_residual(self, D, Dprev, Dold, alpha=None):
        """D.relax(alpha) in place (skipped when alpha is None), then unsTotalLagrangianStress::residual()."""
        assert D.dtype == np.float64 and D.flags.c_contiguous
        res = C.c_double(0.0)
        self._check(self.lib.gpupcg_relax_residual(self.h, float(alpha if alpha is not None else 1.0), 0 if alpha is None else 1,
                                                   _dp(D), _dp(_f64(Dprev)), _dp(_f64(Dold)), C.byref(res)))
        return res.value

    # ---- resident fields (round 2) ----
    def _field_shape(self, name):
        n = C.c_longlong(0)
        self._check(self.lib.gpupcg_mesh_field_size(self.h, FIELD_IDS[name], C.byref(n)))
        k = _FIELD_COLS[name]
        return (n.value // k, k) if k > 1 else (n.value,)

    def set_field(self, name, a):
        if name == "pointD":
            self._ensure_points()
        a = _f64(a)
        shape = self._field_shape(name)
        assert a.size == int(np.prod(shape)), (name, a.shape, shape)
        self._check(self.lib.gpupcg_mesh_set_field(self.h, FIELD_IDS[name], _dp(a)))

    def get_field(self, name):
        out = np.empty(self._field_shape(name))
        self._check(self.lib.gpupcg_mesh_get_field(self.h, FIELD_IDS[name], _dp(out)))
        return out

    def assemble_scalar(self, sign, bfKind, bfValue, gammaCells=None, gammaFaces=None, ddt="steadyState", deltaT=1.0, rate=None,
                        psiOld=None, phi=None, refCell=-1, refValue=0.0, psi=None, ls=None, limitCoeff=1.0):
        """gpupcg_assemble_scalar: TEqn (sign +1, Euler ddt, gammaCells) / pEqn (sign -1, gammaFaces, phi, reference cell).
        psi + ls = (lsP, lsN, lsB) (geometry.least_squares_vectors): with the explicit skewCorrected correction."""
        class Eqn(C.Structure):
            _fields_ = [("sign", C.c_int), ("gammaCells", c_double_p), ("gammaFaces", c_double_p), ("bfKind", c_int_p),
                        ("bfValue", c_double_p), ("ddtScheme", C.c_int), ("deltaT", C.c_double), ("rate", c_double_p),
                        ("psiOld", c_double_p), ("phi", c_double_p), ("refCell", C.c_int), ("refValue", C.c_double),
                        ("psi", c_double_p), ("lsP", c_double_p), ("lsN", c_double_p), ("lsB", c_double_p), ("limitCoeff", C.c_double)]
        lsv = (None, None, None) if psi is None else ls
        keep = [None if a is None else _f64(a) for a in (gammaCells, gammaFaces, bfValue, rate, psiOld, phi, psi, lsv[0], lsv[1], lsv[2])]
        kind = _i32(bfKind)
        ptr = lambda a: None if a is None else _dp(a)
        e = Eqn(int(sign), ptr(keep[0]), ptr(keep[1]), _ip(kind), ptr(keep[2]), 1 if ddt == "Euler" else 0, float(deltaT),
                ptr(keep[3]), ptr(keep[4]), ptr(keep[5]), int(refCell), float(refValue),
                ptr(keep[6]), ptr(keep[7]), ptr(keep[8]), ptr(keep[9]), float(limitCoeff))
        nC, nIF, nBF = self.nC, self.nIF, self.nBF
        out = dict(upper=np.empty(nIF), diag=np.empty(nC), source=np.empty(nC), internalCoeffs=np.empty(max(nBF, 1)),
                   boundaryCoeffs=np.empty(max(nBF, 1)))
        ms = C.c_double(0.0)
        self.lib.gpupcg_assemble_scalar.restype = C.c_int
        self.lib.gpupcg_assemble_scalar.argtypes = [C.c_void_p, C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                                    C.POINTER(C.c_double)]
        self._check(self.lib.gpupcg_assemble_scalar(self.h, C.byref(e), _dp(out["upper"]), _dp(out["diag"]), _dp(out["source"]),
                                                     _dp(out["internalCoeffs"]), _dp(out["boundaryCoeffs"]), C.byref(ms)))
        out["internalCoeffs"] = out["internalCoeffs"][:nBF]; out["boundaryCoeffs"] = out["boundaryCoeffs"][:nBF]
        out["kernel_ms"] = ms.value
        return out

    def set_point_constraints(self, pc):
        """pc: dict(point, start, kind, vec) -- the point patch fields of pointD (lsq.point_constraints)."""
        pt, st, kd, vc = _i32(pc["point"]), _i32(pc["start"]), _i32(pc["kind"]), _f64(pc["vec"])
        self._check(self.lib.gpupcg_mesh_set_point_constraints(self.h, int(pt.shape[0]), _ip(pt), _ip(st), _ip(kd), _dp(vc)))

    def set_lsq(self, lsq):
        """lsq: dict(start, entry, inv, origin, mirror) -- what leastSquaresVolPointInterpolation holds (lsq.build)."""
        self._ensure_points()
        st, en, iv, og, mi = _i32(lsq["start"]), _i32(lsq["entry"]), _f64(lsq["inv"]), _f64(lsq["origin"]), \
            np.ascontiguousarray(lsq["mirror"], dtype=np.uint8)
        self._check(self.lib.gpupcg_mesh_set_lsq(self.h, _ip(st), _ip(en), _dp(iv), _dp(og), mi.ctypes.data_as(C.POINTER(C.c_ubyte))))

    def vol_to_point(self, reps=1):
        ms = C.c_double(0.0)
        self._check(self.lib.gpupcg_vol_to_point(self.h, int(reps), C.byref(ms)))
        return ms.value

    def patch_evaluate(self):
        self._check(self.lib.gpupcg_patch_evaluate(self.h))

    def assemble_resident(self, eqn, g=(0.0, 0.0, 0.0)):
        gv = _f64(g)
        self._check(self.lib.gpupcg_assemble_D_resident(self.h, C.byref(eqn) if eqn is not None else None, _dp(gv)))

    def gradients_on_device(self, limitCoeff=1.0):
        self._ensure_points()
        self._check(self.lib.gpupcg_gradients_resident(self.h, float(limitCoeff)))

    def sigma_resident(self, eqn=None):
        self._check(self.lib.gpupcg_sigma_resident(self.h, C.byref(eqn) if eqn is not None else None))

    def new_time_step(self):
        self._check(self.lib.gpupcg_mesh_new_time_step(self.h))

    def copy_field(self, dst, src):
        """dst = src on the device (names of FIELD_IDS)."""
        self._check(self.lib.gpupcg_mesh_copy_field(self.h, FIELD_IDS[dst], FIELD_IDS[src]))

    def sigma(self, mu, lam, gradD):
        """gradD = None: the device's grad(D)."""
        eps = np.empty((self.nC, 6)); sig = np.empty((self.nC, 6))
        self._check(self.lib.gpupcg_sigma(self.h, _dp(_f64(mu)), _dp(_f64(lam)), _dp(_f64(gradD)) if gradD is not None else None,
                                          _dp(eps), _dp(sig)))
        return eps, sig


def plan_export(lowerAddr, upperAddr, nCells, tileRows=1024, tileMode=1):
    """Host-only schedule export (no GPU): dict with info and every schedule array (see include/gpupcg.h)."""
    lib = load_library()
    l, u = _i32(lowerAddr), _i32(upperAddr)
    info = PlanInfo()
    rc = lib.gpupcg_plan_export(int(nCells), l.shape[0], _ip(l), _ip(u), int(tileRows), int(tileMode), C.byref(info), None)
    if rc != 0:
        raise GpuPcgError(rc, lib.gpupcg_last_error(None).decode())
    d = info.as_dict()
    sizes = [("rowOld", d["nRows"]), ("slices", 4*d["nSlices"]), ("Ucol", d["entriesUpper"]),
             ("UfaceOld", d["entriesUpper"]), ("Lidx", d["entriesLower"]), ("Lcol", d["entriesLower"]),
             ("tileStart", d["nTiles"] + 1), ("tileRoundPtr", d["nTiles"] + 1), ("roundStart", d["roundEntries"]),
             ("LcolTile", d["entriesLower"]), ("UcolTile", d["entriesUpper"]), ("extFPtr", d["nTiles"] + 1),
             ("extFCol", d["extForward"]), ("extBPtr", d["nTiles"] + 1), ("extBCol", d["extBackward"])]
    arrs = [np.zeros(max(1, n), dtype=np.int32) for _, n in sizes]
    ptrs = (c_int_p * len(arrs))(*[_ip(a) for a in arrs])
    rc = lib.gpupcg_plan_export(int(nCells), l.shape[0], _ip(l), _ip(u), int(tileRows), int(tileMode), C.byref(info), ptrs)
    if rc != 0:
        raise GpuPcgError(rc, lib.gpupcg_last_error(None).decode())
    out = dict(info=d)
    for (name, n), a in zip(sizes, arrs):
        out[name] = a[:n]
    out["slices"] = out["slices"].reshape(-1, 4)
    return out


def comm_unique_id():
    lib = load_library()
    buf = C.create_string_buffer(128)
    rc = lib.gpupcg_comm_unique_id(buf)
    if rc != 0:
        raise GpuPcgError(rc, "ncclGetUniqueId failed (is libnccl.so.2 loadable?)")
    return buf.raw


def device_report(device=0):
    lib = load_library()
    buf = C.create_string_buffer(256)
    rc = lib.gpupcg_device_report(device, buf, 256)
    if rc != 0:
        raise GpuPcgError(rc, "no CUDA device %d" % device)
    return buf.value.decode()
