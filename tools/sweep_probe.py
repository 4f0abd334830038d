"""Scratch probe: DIC sweep time vs spin back-off and persistent grid size (run on the GPU box)."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    import fluidstructureinteraction_b200 as pkg
    nx, ny, nz = [int(v) for v in os.environ.get("CELLS", "200,50,100").split(",")]
    s = pkg.cases.cantilever_system(nx, ny, nz)
    g = pkg.GpuPCG(s["l"], s["u"], s["nCells"])
    g.set_matrix(s["diag"], s["upper"])
    out = {k: g.time_kernel(k, reps=5, flushL2=True) for k in ("amul", "dic", "dic_factor", "p_update", "xr_update")}
    out["plan"] = g.plan_info()["nLevels"]
    print(json.dumps(out))
else:
    for cells in ("200,50,100", "400,100,200"):
        for ns in (0, 20, 200):
            for bps in (1, 2, 4, 8):
                env = dict(os.environ, GPUPCG_SPIN_NS=str(ns), GPUPCG_SWEEP_BPS=str(bps), CELLS=cells)
                r = subprocess.run([sys.executable, __file__, "child"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
                print(cells, "ns", ns, "bps", bps, r.stdout.strip().split("\n")[-1], flush=True)
