"""Step time / hop cost of the streaming pencil sweeps: DIC precondition (forward + backward) timed on boxes made of
1, 2, ... pencils.  usage: python tools/pencil_probe.py [big]"""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fluidstructureinteraction_b200 as pkg

shapes = [(200, 4, 8), (400, 4, 8), (200, 8, 8), (200, 16, 32), (200, 50, 100)]
if len(sys.argv) > 1 and sys.argv[1] == "big":
    shapes = [(200, 50, 100), (400, 100, 200)]
out = []
for shape in shapes:
    s = pkg.cases.cantilever_system(*shape)
    n = s["nCells"]
    g = pkg.GpuPCG(s["l"], s["u"], n, device=0)
    info = g.plan_info()
    g.set_matrix(s["diag"], s["upper"])
    t1 = g.time_kernel("dic", reps=20, flushL2=True)
    dC = np.ascontiguousarray(np.stack([s["diag"]]*3, axis=1))
    g.set_matrix_vector(dC, s["upper"])
    t3 = g.time_kernel_vector(3, "dic", reps=20, flushL2=True)
    a3 = g.time_kernel_vector(3, "amul", reps=20, flushL2=True)
    out.append(dict(shape=shape, nTiles=info["nTiles"], rows=info["nRows"], scalar_us=round(1e3*t1, 1), batched3_us=round(1e3*t3, 1),
                    amul3_us=round(1e3*a3, 1), env={k: v for k, v in os.environ.items() if k.startswith("GPUPCG_")}))
    print(out[-1], flush=True)
    g.close()
