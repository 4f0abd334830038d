"""Kernel timings on the Kuhn-split tet cantilever (greedy tiles: the numbering is not a structured block).
usage: python tools/profile_tet.py [nx,ny,nz]   (cells = 6*nx*ny*nz)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fluidstructureinteraction_b200 as pkg
from fluidstructureinteraction_b200 import meshes
nx, ny, nz = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "40,20,30").split(",")]
t0 = time.time()
m = meshes.tet_box(nx, ny, nz, jitter=0.0)
l, u = m.lower_upper()
n = m.nCells
print("tet mesh: %d cells, %d internal faces (%.1f s)" % (n, l.shape[0], time.time() - t0))
rng = np.random.default_rng(1)
upper = -(0.5 + rng.random(l.shape[0]))
diag = np.zeros(n); np.subtract.at(diag, l, upper); np.subtract.at(diag, u, upper); diag += 0.01
b = rng.standard_normal(n)
g = pkg.GpuPCG(l, u, n)
info = g.plan_info()
print({k: info[k] for k in ("nTiles", "tileRowsMax", "maxRounds", "nLevels", "sweepSmemBytes", "extForward")})
g.set_matrix(diag, upper)
x = np.zeros(n)
perf, hist = g.solve(x, b, 1e-9, 0.01, 0, 40)
print("solve:", perf["nIterations"], "iterations, %.3f ms" % perf["solve_ms"])
for k in ("amul", "dic", "p_update", "xr_update"):
    print(k, "%.4f ms" % g.time_kernel(k, reps=3, flushL2=True))
g.close()
