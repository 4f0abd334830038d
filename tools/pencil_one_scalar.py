"""One scalar DIC precondition (forward + backward pencil sweep of ONE system) and one scalar Amul on a box (ncu target
for the kernels of the concurrent component solves).  usage: python tools/pencil_one_scalar.py nx ny nz [reps]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fluidstructureinteraction_b200 as pkg
shape = tuple(int(v) for v in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
s = pkg.cases.cantilever_system(*shape)
g = pkg.GpuPCG(s["l"], s["u"], s["nCells"], device=0)
g.set_matrix(s["diag"], s["upper"])
print(shape, "dic", 1e3*g.time_kernel("dic", reps=reps, flushL2=True), "us", "amul", 1e3*g.time_kernel("amul", reps=reps, flushL2=True), "us")
g.close()
