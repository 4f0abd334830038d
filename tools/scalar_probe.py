"""Scalar kernels only: timings under the current environment.  usage: scalar_probe.py nx ny nz"""
import os, sys
import numpy as np
sys.path.insert(0, ".")
import fluidstructureinteraction_b200 as pkg
from fluidstructureinteraction_b200 import cases
nx, ny, nz = [int(v) for v in sys.argv[1:4]] if len(sys.argv) > 3 else (200, 50, 100)
s = cases.cantilever_system(nx, ny, nz)
g = pkg.GpuPCG(s["l"], s["u"], s["nCells"])
g.set_matrix(s["diag"], s["upper"])
x = np.zeros(s["nCells"])
p, _ = g.solve(x, s["b"]*-1e4, 1e-9, 0.01, history=False)
x[:] = 0
p, _ = g.solve(x, s["b"]*-1e4, 1e-9, 0.01, history=False)
print("scalar", (nx, ny, nz), "its", p["nIterations"], "solve_ms %.2f" % p["solve_ms"],
      " ".join("%s %.4f" % (k, g.time_kernel(k, 10, True)) for k in ("amul", "dic", "p_update")))
