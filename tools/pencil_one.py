"""One batched DIC precondition on a box (ncu target).  usage: python tools/pencil_one.py nx ny nz [reps]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fluidstructureinteraction_b200 as pkg
shape = tuple(int(v) for v in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
s = pkg.cases.cantilever_system(*shape)
g = pkg.GpuPCG(s["l"], s["u"], s["nCells"], device=0)
dC = np.ascontiguousarray(np.stack([s["diag"]]*3, axis=1))
g.set_matrix_vector(dC, s["upper"])
print(shape, 1e3*g.time_kernel_vector(3, "dic", reps=reps, flushL2=True), "us")
g.close()
