import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, oracle
import fluidstructureinteraction_b200 as pkg
from conftest import synthetic_matrix
for dims in [(24, 8, 12), (40, 10, 20)]:
    nx, ny, nz = dims
    l, u = pkg.meshes.hex_ldu(nx, ny, nz); n = nx*ny*nz
    diag, upper, b, x0 = synthetic_matrix(l, u, n, 11)
    g = pkg.GpuPCG(l, u, n); g.set_matrix(diag, upper)
    info = g.plan_info()
    print(dims, {k: info[k] for k in ("nTiles", "tileRowsMax", "maxRounds", "sweepSmemBytes", "extForward")})
    rD = oracle.dic_factor(l, u, diag, upper)
    rg = g.dic_factor()
    bad = np.flatnonzero(rg != rD); print("factor mismatches", bad.size, bad[:10], "nan:", int(np.isnan(rg).sum()))
    w = oracle.precondition(l, u, upper, rD, b); wg = g.precondition(b)
    bad = np.flatnonzero(wg != w); print("precond mismatches", bad.size, bad[:10])
    for rep in range(2):
        wg2 = g.precondition(b); print(" repeat mismatches", np.flatnonzero(wg2 != w).size)
