"""Short, deterministic run for ncu: a few iterations of the component-batched solve on resident data.
usage: python tools/profile_target_vec.py [nx,ny,nz]   (default 200,50,100 = the bench workload)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fluidstructureinteraction_b200 as pkg
nx, ny, nz = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "200,50,100").split(",")]
s = pkg.cases.cantilever_system(nx, ny, nz)
n = s["nCells"]
g = pkg.GpuPCG(s["l"], s["u"], n)
g.set_matrix_vector(np.ascontiguousarray(np.stack([s["diag"]]*3, axis=1)), s["upper"])
x = np.zeros((n, 3))
bC = np.ascontiguousarray(np.stack([s["b"]*t for t in (0.5e4, 0.25e4, -1e4)], axis=1))
perf, _ = g.solve_vector(x, bC, 1e-9, 0.01, 0, 12, history=False)
print("solve_vector:", [p["nIterations"] for p in perf], "iterations, %.3f ms" % perf[0]["solve_ms"], "mode", g.vector_mode(3))
for k in ("amul", "dic", "p_update", "xr_update"):
    print(k, "%.4f ms" % g.time_kernel_vector(3, k, reps=3, flushL2=True))
g.close()
