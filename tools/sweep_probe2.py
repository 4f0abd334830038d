"""Scratch probe: per-level latency of the DIC sweeps for different mesh shapes (run on the GPU box)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fluidstructureinteraction_b200 as pkg
for cells in [(4000, 1, 1), (2000, 2, 1), (1000, 32, 1), (1000, 64, 1), (500, 500, 1), (100, 100, 100), (200, 50, 100)]:
    nx, ny, nz = cells
    s = pkg.cases.cantilever_system(nx, ny, nz)
    g = pkg.GpuPCG(s["l"], s["u"], s["nCells"])
    g.set_matrix(s["diag"], s["upper"])
    info = g.plan_info()
    t = g.time_kernel("dic", reps=5, flushL2=False)
    tf = g.time_kernel("dic_factor", reps=5, flushL2=False)
    print(cells, "levels", info["nLevels"], "slices", info["nSlices"], "sweepGrid", info["sweepGrid"],
          "dic ms %.3f -> %.0f ns/level-step; factor ms %.3f -> %.0f ns/level" %
          (t, 1e6*t/(2*info["nLevels"]), tf, 1e6*tf/info["nLevels"]), flush=True)
    g.close()
