"""Assembly time at 8 M cells for several grid sizes (GPUPCG_ASM_GRID_MULT = CTAs per SM of the grid-stride kernels)."""
import os, sys
sys.path.insert(0, ".")
import numpy as np
import fluidstructureinteraction_b200 as pkg
from fluidstructureinteraction_b200 import geometry, meshes, cases
nx, ny, nz = 400, 100, 200
m = meshes.hex_box(nx, ny, nz); geom = geometry.compute(m)
patches = {"fixed": "fixedDisplacement", "loaded": "tractionDisplacement", "free": "tractionDisplacement"}
gm = pkg.GpuMesh(m, geom, patches)
rng = np.random.default_rng(0)
nC, nF, nBF = m.nCells, m.nFaces, m.nFaces - m.nInternalFaces
mu, lam = cases.lame(cases.STEEL["E"], cases.STEEL["nu"])
D = 1e-5*rng.standard_normal((nC, 3)); G = 1e-5*rng.standard_normal((nC, 9)); Gf = 1e-5*rng.standard_normal((nF, 9))
trac = np.zeros((nBF, 3)); pres = np.zeros(nBF); fixed = np.zeros((nBF, 3))
a = (np.full(nC, mu), np.full(nC, lam), np.full(nC, 7854.0), np.zeros(3), D, G, Gf, trac, pres, fixed, 1.0)
gm.assemble_D(*a, reps=3)
for rep in range(2):
    for mult in ("16", "8", "24", "32", "64", "0"):
        os.environ["GPUPCG_ASM_GRID_MULT"] = mult
        t = min(gm.assemble_D(*a, reps=10)["kernel_ms"] for _ in range(3))
        print("GRID_MULT", mult, "%.4f ms  frac %.3f" % (t, 856.0*nC/t/1e6/6549.8), flush=True)
