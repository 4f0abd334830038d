"""Per-brick-wave timeline of one forward DIC sweep: finish time vs brick wave index (slope = cost per tile hop).
env CELLS=nx,ny,nz  GPUPCG_BRICK=bx,by,bz  GPUPCG_TILE_ROWS"""
import os, sys
os.environ["GPUPCG_TILE_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fluidstructureinteraction_b200 as pkg
nx, ny, nz = [int(v) for v in os.environ.get("CELLS", "200,50,100").split(",")]
bx, by, bz = [int(v) for v in os.environ.get("GPUPCG_BRICK", "8,8,4").split(",")]
s = pkg.cases.cantilever_system(nx, ny, nz)
g = pkg.GpuPCG(s["l"], s["u"], s["nCells"])
g.set_matrix(s["diag"], s["upper"])
info = g.plan_info()
t = g.time_kernel("dic", reps=2, flushL2=False)
tr = g.tile_trace().astype(np.int64)
tr -= tr[:, 0].min()
tr = tr/1e3
bl = []
for c in range(-(-nz//bz)):
    for b in range(-(-ny//by)):
        for a in range(-(-nx//bx)):
            bl.append((a + b + c, c, b, a))
bl.sort()
assert len(bl) == tr.shape[0], (len(bl), tr.shape)
wave = np.array([x[0] for x in bl])
print("brick", (bx, by, bz), "tiles", len(bl), "waves", wave.max() + 1, "rounds/tile", info["maxRounds"], "dic ms %.4f" % t,
      "smem", info["sweepSmemBytes"])
print("staging us: mean %.2f max %.2f;  sweep (staged->done) mean %.2f min %.2f max %.2f" % (
    (tr[:, 1] - tr[:, 0]).mean(), (tr[:, 1] - tr[:, 0]).max(), (tr[:, 2] - tr[:, 1]).mean(), (tr[:, 2] - tr[:, 1]).min(), (tr[:, 2] - tr[:, 1]).max()))
W = wave.max() + 1
rows = []
for w in range(W):
    m = wave == w
    rows.append((w, int(m.sum()), tr[m, 0].min(), tr[m, 1].max(), tr[m, 2].min(), tr[m, 2].max()))
step = max(1, W//16)
for r in rows[::step] + [rows[-1]]:
    print("wave %3d n=%4d  first start %7.1f  last staged %7.1f  finish min %7.1f max %7.1f" % r)
fin = np.array([r[5] for r in rows])
print("finish(max) slope over middle half: %.3f us/wave" % ((fin[3*W//4] - fin[W//4])/(3*W//4 - W//4)))
# corner-to-corner chain along a=b=c diagonal-ish: tile (a,0,0) for a = 0..: finish times
idx = {x[1:]: k for k, x in enumerate(bl)}
ch = [tr[idx[(0, 0, a)], 2] for a in range(-(-nx//bx))]
print("x-chain finish (c=0,b=0):", np.round(ch[:12], 2))
ch = [tr[idx[(c, 0, 0)], 2] for c in range(-(-nz//bz))]
print("z-chain finish (b=0,a=0):", np.round(ch[:12], 2))
