import os, sys
os.environ["GPUPCG_TILE_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fluidstructureinteraction_b200 as pkg
nx, ny, nz = [int(v) for v in os.environ.get("CELLS", "200,50,100").split(",")]
s = pkg.cases.cantilever_system(nx, ny, nz)
g = pkg.GpuPCG(s["l"], s["u"], s["nCells"])
g.set_matrix(s["diag"], s["upper"])
t = g.time_kernel("dic", reps=2, flushL2=False)
tr = g.tile_trace().astype(np.int64)
tr -= tr[:, 0].min()
fin = tr[:, 2]/1e3
stg = tr[:, 1]/1e3
print("finish times (us) of tiles 0..59, 10 per row (a row = one z-plane of pencils, roughly):")
for r in range(6):
    print(" ".join("%7.1f" % fin[10*r + c] for c in range(10)))
print("staged times:")
for r in range(3):
    print(" ".join("%7.1f" % stg[10*r + c] for c in range(10)))
d = np.diff(fin[::10][:40])
print("plane-to-plane lag of pencil 0 (us):", np.round(d, 1))
