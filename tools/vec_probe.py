"""Batched (vector) solve probe: kernel timings and a full 3-component solve vs three scalar solves.  usage: vec_probe.py nx ny nz"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import fluidstructureinteraction_b200 as pkg
from fluidstructureinteraction_b200 import cases

nx, ny, nz = [int(v) for v in sys.argv[1:4]] if len(sys.argv) > 3 else (200, 50, 100)
s = cases.cantilever_system(nx, ny, nz)
l, u, diag, upper, b, n = s["l"], s["u"], s["diag"], s["upper"], s["b"], s["nCells"]
g = pkg.GpuPCG(l, u, n)
print("plan", g.plan_info())
dC = np.ascontiguousarray(np.stack([diag]*3, axis=1))
bC = np.ascontiguousarray(np.stack([b*0.5e4, b*0.25e4, b*-1e4], axis=1))
g.set_matrix_vector(dC, upper)
print("vector_mode", g.vector_mode(3))
x = np.zeros((n, 3))
for rep in range(3):
    x[:] = 0
    t0 = time.perf_counter()
    pv, _ = g.solve_vector(x, bC, 1e-9, 0.01, history=False)
    dt = time.perf_counter() - t0
    print("vector solve: its", [p["nIterations"] for p in pv], "solve_ms %.2f wall %.2f ms" % (pv[0]["solve_ms"], dt*1e3))
g.set_matrix(diag, upper)
for rep in range(2):
    tot = 0.0
    for k in range(3):
        xs = np.zeros(n)
        ps, _ = g.solve(xs, np.ascontiguousarray(bC[:, k]), 1e-9, 0.01, history=False)
        tot += ps["solve_ms"]
    print("3 scalar solves: its", ps["nIterations"], "solve_ms %.2f" % tot, "max diff", np.max(np.abs(xs - x[:, 2]))/np.max(np.abs(xs)))
for k in ("amul", "dic", "p_update", "xr_update", "dic_factor"):
    print("%-10s scalar %.4f ms   vector(3) %.4f ms" % (k, g.time_kernel(k, 10, True), g.time_kernel_vector(3, k, 10, True)))
