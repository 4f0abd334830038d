"""Generates the committed fixtures under tests/golden/ -- run HERE (the build container), never on the GPU box.

  1. *.polymesh.npz : the four polyMeshes the reference ships (run/fsiFoam/{beamInCrossFlow,3dTube}/
     {solid,fluid}/constant/polyMesh), converted verbatim to numpy arrays.  /root/reference does not exist
     on the GPU box, so tests read these copies.  (Mesh DATA of the reference's tutorial cases, not source.)
  2. *.golden.npz   : oracle outputs (Amul, DIC factor, DIC apply, PCG residual history, solution) for a
     deterministic laplacian-type LDU matrix on each mesh.  They pin the oracle against regressions; they do
     NOT pin it against foam-extend (parity unpinned, see oracle/ASSUMPTIONS.md).

usage: python tests/golden/make_fixtures.py [--polymesh-only]
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from fluidstructureinteraction_b200 import meshes  # noqa: E402
import oracle  # noqa: E402

REF = "/root/reference/run/fsiFoam"
CASES = {
    "beam_solid": "beamInCrossFlow/solid",
    "beam_fluid": "beamInCrossFlow/fluid",
    "tube_solid": "3dTube/solid",
    "tube_fluid": "3dTube/fluid",
}


def synthetic_matrix(l, u, nCells, seed):
    """SPD, diagonally dominant LDU matrix with laplacian structure and O(1)-varying coefficients."""
    rng = np.random.default_rng(seed)
    coef = rng.uniform(0.5, 2.0, size=l.shape[0])
    upper = -coef
    diag = np.bincount(l, weights=coef, minlength=nCells) + np.bincount(u, weights=coef, minlength=nCells)
    diag = diag + rng.uniform(0.0, 0.05, size=nCells)*(1.0 + diag)
    b = rng.standard_normal(nCells)
    x0 = 0.1*rng.standard_normal(nCells)
    return diag, upper, b, x0


def main():
    git = subprocess.run(["git", "-C", ROOT, "rev-parse", "HEAD"], stdout=subprocess.PIPE, text=True).stdout.strip()
    for k, (name, rel) in enumerate(sorted(CASES.items())):
        m = meshes.read_polymesh(os.path.join(REF, rel, "constant/polyMesh"))
        assert m.is_upper_triangular(), name
        m.save_npz(os.path.join(HERE, name + ".polymesh.npz"))
        if "--polymesh-only" in sys.argv:
            print(name, [(p["name"], p["type"], p["size"]) for p in m.patches])
            continue
        l, u = m.lower_upper()
        diag, upper, b, x0 = synthetic_matrix(l, u, m.nCells, 1000 + k)
        xt = np.sin(0.37*np.arange(m.nCells))
        Ax = oracle.amul(l, u, diag, upper, xt)
        rD = oracle.dic_factor(l, u, diag, upper)
        wA = oracle.precondition(l, u, upper, rD, b)
        x = x0.copy()
        perf, hist = oracle.pcg_solve(l, u, diag, upper, x, b, tolerance=1e-9, relTol=0.0, maxIter=1000)
        np.savez_compressed(os.path.join(HERE, name + ".golden.npz"), seed=np.int64(1000 + k), Ax=Ax, rD=rD, wA=wA,
                            history=hist, x=x, initialResidual=perf["initialResidual"],
                            finalResidual=perf["finalResidual"], nIterations=np.int64(perf["nIterations"]),
                            normFactor=perf["normFactor"], oracle_git=np.array(git))
        print(name, "cells", m.nCells, "faces", m.nInternalFaces, "iters", perf["nIterations"],
              "final", perf["finalResidual"])


if __name__ == "__main__":
    main()
