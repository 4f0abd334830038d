import os, sys
os.environ["GPUPCG_TILE_TRACE"] = "1"
os.environ["GPUPCG_TILE_ROWS"] = "512"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fluidstructureinteraction_b200 as pkg
for dims in [(8, 8, 8), (8, 8, 16), (8, 8, 32), (8, 8, 64), (16, 8, 8), (64, 8, 8), (32, 32, 8)]:
    nx, ny, nz = dims
    s = pkg.cases.cantilever_system(nx, ny, nz)
    g = pkg.GpuPCG(s["l"], s["u"], s["nCells"])
    g.set_matrix(s["diag"], s["upper"])
    info = g.plan_info()
    t = g.time_kernel("dic", reps=3, flushL2=False)
    tr = g.tile_trace().astype(np.int64)
    tr -= tr[:, 0].min()
    fin = tr[:, 2]/1e3; stg = tr[:, 1]/1e3
    print(dims, "tiles", info["nTiles"], "rounds", info["maxRounds"], "levels", info["nLevels"], "dic ms %.3f" % t)
    print("   staged:", np.round(stg[:10], 1), " finish:", np.round(fin[:10], 1))
    g.close()
