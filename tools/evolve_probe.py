"""Wall time of one evolve() time step (C++ driver host/stressModelEvolve.C, nCorrectors outer iterations) with the
equation and the gradients resident on the device (residentMatrix yes) and through host arrays (no).
usage: evolve_probe.py nx ny nz [nCorr]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import scipy.sparse as sp
from fluidstructureinteraction_b200 import cases, geometry, hostapi, meshes

nx, ny, nz = [int(v) for v in sys.argv[1:4]] if len(sys.argv) > 3 else (100, 25, 50)
nCorr = int(sys.argv[4]) if len(sys.argv) > 4 else 4
m = meshes.hex_box(nx, ny, nz); g = geometry.compute(m)
nC, nBF, nIF = m.nCells, m.nFaces - m.nInternalFaces, m.nInternalFaces
fs = np.asarray(m.faceStart); fp = np.asarray(m.facePoints); sz = np.diff(fs)
rows = np.concatenate([fp, fp[:fs[nIF]]]); cols = np.concatenate([np.repeat(np.asarray(m.owner), sz), np.repeat(np.asarray(m.neighbour[:nIF]), sz[:nIF])])
W = sp.coo_matrix((np.ones(rows.shape[0]), (rows, cols)), shape=(m.nPoints, nC)).tocsr(); W.data[:] = 1.0
W = sp.diags(1.0/np.asarray(W.sum(axis=1)).ravel()) @ W
v2p = lambda D: np.ascontiguousarray(W @ D)
mu, lam = cases.lame(cases.STEEL["E"], cases.STEEL["nu"])
trac = np.zeros((nBF, 3)); pres = np.zeros(nBF); fixed = np.zeros((nBF, 3))
p = m.patch("loaded"); b0 = p["start"] - nIF
trac[b0:b0 + p["size"], 2] = -1.0e4
PATCHES = {"fixed": "fixedDisplacement", "loaded": "tractionDisplacement", "free": "tractionDisplacement"}
args = (m, g, PATCHES, np.full(nC, mu), np.full(nC, lam), np.full(nC, cases.STEEL["rho"]), np.array([0.0, 0.0, -9.81]), trac, pres, fixed, v2p,
        "nCorrectors %d; convergenceTolerance 1e-30;" % nCorr)
out = {}
for rep in range(2):
    for res in ("yes", "no"):
        t0 = time.perf_counter()
        r = hostapi.evolve(*args, "solver gpuPCG; preconditioner DIC; tolerance 1e-09; relTol 0.01; residentMatrix %s;" % res)
        dt = time.perf_counter() - t0
        out[res] = r
        print("cells %d nCorr %d residentMatrix %s: %.3f s per time step (incl. mesh upload + schedule), %d PCG iterations" %
              (nC, nCorr, res, dt, r["totalPcgIterations"]), flush=True)
print("fields identical:", np.array_equal(out["yes"]["D"], out["no"]["D"]), np.array_equal(out["yes"]["sigma"], out["no"]["sigma"]))
