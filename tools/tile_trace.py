"""Scratch probe: per-tile timeline of one forward DIC sweep (run on the GPU box)."""
import os, sys
os.environ["GPUPCG_TILE_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fluidstructureinteraction_b200 as pkg
nx, ny, nz = [int(v) for v in os.environ.get("CELLS", "200,50,100").split(",")]
s = pkg.cases.cantilever_system(nx, ny, nz)
g = pkg.GpuPCG(s["l"], s["u"], s["nCells"])
g.set_matrix(s["diag"], s["upper"])
info = g.plan_info()
print({k: info[k] for k in ("nTiles", "tileRowsMax", "maxRounds", "sweepSmemBytes", "extForward", "extBackward")})
t = g.time_kernel("dic", reps=2, flushL2=False)
tr = g.tile_trace().astype(np.int64)
t0 = tr[:, 0].min()
tr -= t0
print("dic ms", t, "trace span us", tr.max()/1e3)
n = tr.shape[0]
for k in list(range(0, min(n, 24))) + list(range(n//2, n//2 + 6)) + list(range(max(0, n - 6), n)):
    print("tile %4d start %9.1f us  staged +%6.1f  swept +%7.1f" % (k, tr[k, 0]/1e3, (tr[k, 1] - tr[k, 0])/1e3, (tr[k, 2] - tr[k, 1])/1e3))
dur = (tr[:, 2] - tr[:, 1])/1e3
print("sweep duration per tile: mean %.1f us, min %.1f, max %.1f; staging mean %.1f us" % (dur.mean(), dur.min(), dur.max(), ((tr[:, 1] - tr[:, 0])/1e3).mean()))
# concurrency: how many tiles are between staged and swept at mid time
mid = tr[:, 2].max()//2
print("tiles in sweep phase at mid time:", int(((tr[:, 1] <= mid) & (tr[:, 2] >= mid)).sum()), " resident:", int(((tr[:, 0] <= mid) & (tr[:, 2] >= mid)).sum()))
