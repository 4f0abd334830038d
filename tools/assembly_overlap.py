"""Assembly at 8 M cells: one chunk vs 2-4 chunks with the cell kernel of chunk i overlapping the face kernel of chunk i+1."""
import os, sys
sys.path.insert(0, ".")
import numpy as np
import fluidstructureinteraction_b200 as pkg
from fluidstructureinteraction_b200 import geometry, meshes, cases
nx, ny, nz = 400, 100, 200
m = meshes.hex_box(nx, ny, nz); geom = geometry.compute(m)
patches = {"fixed": "fixedDisplacement", "loaded": "tractionDisplacement", "free": "tractionDisplacement"}
rng = np.random.default_rng(0)
nC, nF, nBF = m.nCells, m.nFaces, m.nFaces - m.nInternalFaces
mu, lam = cases.lame(cases.STEEL["E"], cases.STEEL["nu"])
D = 1e-5*rng.standard_normal((nC, 3)); G = 1e-5*rng.standard_normal((nC, 9)); Gf = 1e-5*rng.standard_normal((nF, 9))
trac = np.zeros((nBF, 3)); pres = np.zeros(nBF); fixed = np.zeros((nBF, 3))
a = (np.full(nC, mu), np.full(nC, lam), np.full(nC, 7854.0), np.zeros(3), D, G, Gf, trac, pres, fixed, 1.0)
ref = None
for chunk in (0, 4000000, 2700000, 2000000, 1000000):
    if chunk: os.environ["GPUPCG_ASM_CHUNK_CELLS"] = str(chunk)
    else: os.environ.pop("GPUPCG_ASM_CHUNK_CELLS", None)
    gm = pkg.GpuMesh(m, geom, patches)
    for ov in (("1", "0") if chunk else ("1",)):
        os.environ["GPUPCG_ASM_OVERLAP"] = ov
        r = gm.assemble_D(*a, reps=3)
        if ref is None: ref = r
        same = all(np.array_equal(r[k], ref[k]) for k in ("upper", "diag", "source", "internalCoeffs", "boundaryCoeffs"))
        t = min(gm.assemble_D(*a, reps=10)["kernel_ms"] for _ in range(3))
        print("chunk", chunk, "overlap", ov, "%.4f ms  frac %.3f" % (t, 856.0*nC/t/1e6/6549.8), "bit-identical" if same else "DIFFERENT", flush=True)
    gm.close()
