"""p_update / xr_update / Amul of the batched solve over GPUPCG_STREAM_BPS (CTAs per SM of the streaming kernels).
usage: python tools/stream_grid.py"""
import sys, os, subprocess
HERE = os.path.dirname(os.path.abspath(__file__))
code = r'''
import sys, os
import numpy as np
sys.path.insert(0, %r)
import fluidstructureinteraction_b200 as pkg
for shape in [(200, 50, 100), (400, 100, 200)]:
    s = pkg.cases.cantilever_system(*shape)
    g = pkg.GpuPCG(s["l"], s["u"], s["nCells"], device=0)
    dC = np.ascontiguousarray(np.stack([s["diag"]]*3, axis=1))
    g.set_matrix_vector(dC, s["upper"])
    xv = np.zeros((s["nCells"], 3))
    g.solve_vector(xv, np.ascontiguousarray(np.stack([s["b"]]*3, axis=1)), 1e-9, 0.01, 0, 4, history=False)
    print(shape[0], " ".join("%%s %%.1f" %% (k, 1e3*g.time_kernel_vector(3, k, reps=20, flushL2=True)) for k in ("p_update", "xr_update", "amul")), end=" | ")
    g.close()
print()
''' % os.path.dirname(HERE)
for bps in ("2", "3", "4", "6", "8", "12", "16"):
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, GPUPCG_STREAM_BPS=bps), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    print("BPS", bps, r.stdout.strip()[-300:], flush=True)
