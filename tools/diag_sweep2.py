import sys, os
os.environ["GPUPCG_TILE_ROWS"] = "32"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, oracle
import fluidstructureinteraction_b200 as pkg
from conftest import synthetic_matrix
np.set_printoptions(linewidth=200, precision=5)
nx, ny, nz = 8, 4, 2
l, u = pkg.meshes.hex_ldu(nx, ny, nz); n = nx*ny*nz
diag, upper, b, x0 = synthetic_matrix(l, u, n, 11)
g = pkg.GpuPCG(l, u, n); g.set_matrix(diag, upper)
print(g.plan_info())
rD = oracle.dic_factor(l, u, diag, upper)
rg = g.dic_factor()
print("oracle", rD)
print("gpu   ", rg)
p = pkg.plan_export(l, u, n, 32)
print("tileStart", p["tileStart"], "extFPtr", p["extFPtr"], "extFCol", p["extFCol"][:40])
print("rowOld", p["rowOld"])
