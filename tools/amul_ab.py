"""Amul (scalar and batched-3) at 1 M and 8 M cells -- A/B over libraries through GPUPCG_LIBRARY.  usage: python tools/amul_ab.py"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fluidstructureinteraction_b200 as pkg
out = []
for shape in [(200, 50, 100), (400, 100, 200)]:
    s = pkg.cases.cantilever_system(*shape)
    g = pkg.GpuPCG(s["l"], s["u"], s["nCells"], device=0)
    g.set_matrix(s["diag"], s["upper"])
    x = np.zeros(s["nCells"]); g.solve(x, s["b"], 1e-9, 0.01, 0, 4)
    a1 = g.time_kernel("amul", reps=20, flushL2=True)
    g.set_matrix_vector(np.ascontiguousarray(np.stack([s["diag"]]*3, axis=1)), s["upper"])
    xv = np.zeros((s["nCells"], 3)); g.solve_vector(xv, np.ascontiguousarray(np.stack([s["b"]]*3, axis=1)), 1e-9, 0.01, 0, 4, history=False)
    a3 = g.time_kernel_vector(3, "amul", reps=20, flushL2=True)
    out.append("%d: scalar %.1f us, batched3 %.1f us" % (shape[0], 1e3*a1, 1e3*a3))
    g.close()
print(" | ".join(out))
