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
