"""Batched-3 DIC precondition at 1 M (and optionally 8 M) cells over the fetch-lane knobs of the pencil sweeps.
usage: python tools/pencil_knobs.py [big]"""
import sys, os, subprocess, json
HERE = os.path.dirname(os.path.abspath(__file__))
code = r'''
import sys, os
import numpy as np
sys.path.insert(0, %r)
import fluidstructureinteraction_b200 as pkg
for shape in SHAPES:
    s = pkg.cases.cantilever_system(*shape)
    g = pkg.GpuPCG(s["l"], s["u"], s["nCells"], device=0)
    dC = np.ascontiguousarray(np.stack([s["diag"]]*3, axis=1))
    g.set_matrix_vector(dC, s["upper"])
    print(shape, round(1e3*g.time_kernel_vector(3, "dic", reps=20, flushL2=True), 1), end="  ")
    g.close()
print()
''' % os.path.dirname(HERE)
shapes = "[(200, 50, 100), (400, 100, 200)]" if len(sys.argv) > 1 else "[(200, 50, 100)]"
code = code.replace("SHAPES", shapes)
grid = []
for split, cps in (("1", "1"), ("0", "2"), ("0", "1")):
    grid.append({"GPUPCG_PENCIL_SPLIT": split, "GPUPCG_PENCIL_CTAS_PER_SM": cps})
for env in grid:
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    print(" ".join("%s=%s" % (k.replace("GPUPCG_PENCIL_", ""), v) for k, v in env.items()), r.stdout.strip()[-200:], flush=True)
