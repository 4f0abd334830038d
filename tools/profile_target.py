"""Short, deterministic run for ncu: a few PCG iterations' worth of the solver kernels on resident data.
usage: python tools/profile_target.py [nx,ny,nz]   (default 200,50,100 = the bench workload)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fluidstructureinteraction_b200 as pkg
if os.environ.get("GPUPCG_LIB"):
    from fluidstructureinteraction_b200 import capi
    capi.LIB_PATH = os.environ["GPUPCG_LIB"]
nx, ny, nz = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "200,50,100").split(",")]
s = pkg.cases.cantilever_system(nx, ny, nz)
g = pkg.GpuPCG(s["l"], s["u"], s["nCells"])
g.set_matrix(s["diag"], s["upper"])
x = np.zeros(s["nCells"])
perf, hist = g.solve(x, s["b"]*-1e4, 1e-9, 0.01, 0, 12)
print("solve:", perf["nIterations"], "iterations, %.3f ms" % perf["solve_ms"])
for k in ("amul", "dic", "p_update", "xr_update"):
    print(k, "%.4f ms" % g.time_kernel(k, reps=3, flushL2=True))
g.close()
