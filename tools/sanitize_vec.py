import sys
sys.path.insert(0, ".")
import numpy as np
import fluidstructureinteraction_b200 as pkg
from fluidstructureinteraction_b200 import cases
s = cases.cantilever_system(24, 8, 12)
l, u, diag, upper, b, n = s["l"], s["u"], s["diag"], s["upper"], s["b"], s["nCells"]
g = pkg.GpuPCG(l, u, n)
dC = np.ascontiguousarray(np.stack([diag, diag*1.25, diag*1.5], axis=1))
bC = np.ascontiguousarray(np.stack([b, -0.5*b, np.cos(0.11*np.arange(n))*np.abs(b).max()], axis=1))
g.set_matrix_vector(dC, upper)
x = np.zeros((n, 3))
pv, hv = g.solve_vector(x, bC, 1e-9, 0.01)
print("its", [p["nIterations"] for p in pv], "mode", g.vector_mode(3))
dC2 = np.ascontiguousarray(dC[:, :2]); bC2 = np.ascontiguousarray(bC[:, :2])
g2 = pkg.GpuPCG(l, u, n)
g2.set_matrix_vector(dC2, upper)
x2 = np.zeros((n, 2))
pv2, _ = g2.solve_vector(x2, bC2, 1e-9, 0.01)
print("its2", [p["nIterations"] for p in pv2], np.max(np.abs(x2 - x[:, :2])))
