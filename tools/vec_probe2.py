"""Batched kernels only: timings under the current environment.  usage: vec_probe2.py nx ny nz"""
import os, sys
import numpy as np
sys.path.insert(0, ".")
import fluidstructureinteraction_b200 as pkg
from fluidstructureinteraction_b200 import cases
nx, ny, nz = [int(v) for v in sys.argv[1:4]] if len(sys.argv) > 3 else (200, 50, 100)
s = cases.cantilever_system(nx, ny, nz)
l, u, diag, upper, b, n = s["l"], s["u"], s["diag"], s["upper"], s["b"], s["nCells"]
g = pkg.GpuPCG(l, u, n)
pi = g.plan_info()
dC = np.ascontiguousarray(np.stack([diag]*3, axis=1))
bC = np.ascontiguousarray(np.stack([b*0.5e4, b*0.25e4, b*-1e4], axis=1))
g.set_matrix_vector(dC, upper)
x = np.zeros((n, 3))
pv, _ = g.solve_vector(x, bC, 1e-9, 0.01, history=False)
x[:] = 0
pv, _ = g.solve_vector(x, bC, 1e-9, 0.01, history=False)
env = {k: v for k, v in os.environ.items() if k.startswith("GPUPCG_")}
print(env, "tiles", pi["nTiles"], "T", pi["tileRowsMax"], "mode", g.vector_mode(3), "its", [p["nIterations"] for p in pv],
      "solve_ms %.2f" % pv[0]["solve_ms"], " ".join("%s %.4f" % (k, g.time_kernel_vector(3, k, 10, True)) for k in ("amul", "dic", "p_update")))
