"""Per-tile timeline of one forward sweep of the batched kernels (GPUPCG_TILE_TRACE): staging / sweep / lifetime.
usage: tile_trace_vec.py nx ny nz [scalar]"""
import os, sys
os.environ["GPUPCG_TILE_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fluidstructureinteraction_b200 as pkg
nx, ny, nz = [int(v) for v in sys.argv[1:4]]
scalar = len(sys.argv) > 4
s = pkg.cases.cantilever_system(nx, ny, nz)
g = pkg.GpuPCG(s["l"], s["u"], s["nCells"])
info = g.plan_info()
if scalar:
    g.set_matrix(s["diag"], s["upper"])
    t = g.time_kernel("dic", reps=2, flushL2=False)
else:
    g.set_matrix_vector(np.ascontiguousarray(np.stack([s["diag"]]*3, axis=1)), s["upper"])
    t = g.time_kernel_vector(3, "dic", reps=2, flushL2=False)
tr = g.tile_trace().astype(np.int64)
tr -= tr[:, 0].min()
tr = tr/1e3
q = lambda a: "p10 %.2f p50 %.2f p90 %.2f max %.2f" % tuple(np.percentile(a, [10, 50, 90, 100]))
print("scalar" if scalar else "vector(3)", (nx, ny, nz), "tiles", info["nTiles"], "T", info["tileRowsMax"], "rounds/tile", info["maxRounds"],
      "dic ms %.4f" % t, "forward span %.1f us" % tr[:, 2].max())
print("  staging (CTA start -> staged) us:", q(tr[:, 1] - tr[:, 0]))
print("  sweep   (staged -> done)      us:", q(tr[:, 2] - tr[:, 1]))
print("  lifetime                      us:", q(tr[:, 2] - tr[:, 0]))
order = np.argsort(tr[:, 0])
st = tr[order, 0]
n = len(st)
print("  CTA starts: first %.1f, 25%% %.1f, 50%% %.1f, 75%% %.1f, last %.1f us" % (st[0], st[n//4], st[n//2], st[3*n//4], st[-1]))
busy = np.sum(tr[:, 2] - tr[:, 1])
print("  sum of sweep time / span = %.1f sweep warps busy on average (of %d SMs)" % (busy/tr[:, 2].max(), info["numSMs"]))
