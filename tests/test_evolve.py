"""Outer loop of the solid model: unsTotalLagrangianStress::evolve() (unsTotalLagrangianStress.C:769-1009) and
residual() (unsTotalLagrangianStressSolve.C:647-672).  The GPU-driven loop (C++ host side host/stressModelEvolve.C:
gpupcg_assemble_D -> fvMatrix solve through gpuPCG -> gpupcg_relax_residual -> gpupcg_gradients -> gpupcg_sigma) against
the oracle's restatement of the same loop (oracle/evolve.py) with the same vol-to-point interpolation callback."""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle
import oracle.evolve as oevolve
import fluidstructureinteraction_b200 as pkg
from fluidstructureinteraction_b200 import cases, geometry, hostapi, meshes

PATCHES = {"fixed": "fixedDisplacement", "loaded": "tractionDisplacement", "free": "tractionDisplacement"}


def point_average(mesh):
    """Deterministic stand-in for leastSquaresVolPointInterpolation: a point takes the mean of the
    cells around it."""
    rows, cols = [], []
    fs = mesh.faceStart
    for f in range(mesh.nFaces):
        pts = mesh.facePoints[fs[f]:fs[f + 1]]
        cs = [mesh.owner[f]] + ([mesh.neighbour[f]] if f < mesh.nInternalFaces else [])
        for c in cs:
            rows.extend(pts); cols.extend([c]*len(pts))
    W = sp.coo_matrix((np.ones(len(rows)), (rows, cols)), shape=(mesh.nPoints, mesh.nCells)).tocsr()
    W.data[:] = 1.0                                        # a (point, cell) pair is listed once per shared face
    W = sp.diags(1.0/np.asarray(W.sum(axis=1)).ravel()) @ W
    return lambda D: np.ascontiguousarray(W @ D)


def problem(kind):
    m = meshes.hex_box(10, 4, 6) if kind == "hex" else meshes.tet_box(5, 3, 4, jitter=0.1)
    g = geometry.compute(m)
    nC, nBF = m.nCells, m.nFaces - m.nInternalFaces
    mu, lam = cases.lame(cases.STEEL["E"], cases.STEEL["nu"])
    trac = np.zeros((nBF, 3)); pres = np.zeros(nBF); fixed = np.zeros((nBF, 3))
    p = m.patch("loaded"); b0 = p["start"] - m.nInternalFaces
    trac[b0:b0 + p["size"], 2] = -1.0e4
    return m, g, np.full(nC, mu), np.full(nC, lam), np.full(nC, cases.STEEL["rho"]), np.array([0.0, 0.0, -9.81]), trac, pres, fixed


def test_oracle_residual_definition():
    rng = np.random.default_rng(3)
    D = rng.standard_normal((50, 3)); Dp = D + 1e-3*rng.standard_normal((50, 3)); Do = np.zeros((50, 3))
    r = oracle.relax_residual(D.copy(), Dp, Do)
    maxDD = np.max(np.sqrt(np.sum((D - Do)**2, axis=1)))
    assert r == pytest.approx(np.max(np.sqrt(np.sum((D - Dp)**2, axis=1))/(maxDD + 1e-15)), rel=1e-14)
    D2 = D.copy()
    oracle.relax_residual(D2, Dp, Do, 0.7)
    assert np.allclose(D2, Dp + 0.7*(D - Dp), rtol=1e-15, atol=0)


def test_oracle_evolve_outer_loop_contracts():
    """The fixed-point iteration on the explicit stress terms contracts: relative residual falls monotonically after the
    first couple of outer iterations, and the loop stops on nCorrectors / on the tolerance exactly as the do-while does."""
    m, g, mu, lam, rho, grav, trac, pres, fixed = problem("hex")
    v2p = point_average(m)
    r = oevolve.evolve(m, g, PATCHES, mu, lam, rho, grav, trac, pres, fixed, v2p, nCorrectors=8, convergenceTolerance=1e-12)
    h = r["residualHistory"]
    assert r["nOuterIterations"] == 8 and len(h) == 8          # ++iCorr < nCorr failed after the 8th pass
    assert h[0] == pytest.approx(1.0, rel=1e-6)                 # first pass: D - prevIter == D - oldTime (up to SMALL)
    assert np.all(np.diff(h[1:]) < 0)
    r2 = oevolve.evolve(m, g, PATCHES, mu, lam, rho, grav, trac, pres, fixed, v2p, nCorrectors=50, convergenceTolerance=h[3]*1.0001)
    assert r2["nOuterIterations"] == 3 and len(r2["residualHistory"]) == 4


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["hex", "tet"])
def test_gpu_relax_residual_bit_exact(kind):
    m, g, *_ = problem(kind)
    gm = pkg.GpuMesh(m, g, PATCHES)
    rng = np.random.default_rng(5)
    D = 1e-4*rng.standard_normal((m.nCells, 3)); Dp = D + 1e-6*rng.standard_normal((m.nCells, 3)); Do = 1e-5*rng.standard_normal((m.nCells, 3))
    for alpha in (None, 0.35):
        a, b = D.copy(), D.copy()
        assert gm.relax_residual(a, Dp, Do, alpha) == oracle.relax_residual(b, Dp, Do, alpha)
        assert np.array_equal(a, b)
    gm.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,relax,resident", [("hex", None, "yes"), ("tet", 0.8, "yes"), ("hex", None, "no"), ("tet", 0.8, "no")])
def test_gpu_evolve_matches_oracle_loop(kind, relax, resident):
    """residentMatrix yes (default): the assembled D equation never leaves the device (gpupcg_solve_assembled_D) and the
    gradients stay there between outer iterations; no: the fvMatrix is rebuilt on the host every outer iteration and
    solved through lduMatrix::solver::New -- the same loop statistics and fields either way."""
    m, g, mu, lam, rho, grav, trac, pres, fixed = problem(kind)
    v2p = point_average(m)
    nCorr = 6
    ref = oevolve.evolve(m, g, PATCHES, mu, lam, rho, grav, trac, pres, fixed, v2p, nCorrectors=nCorr,
                         convergenceTolerance=1e-12, relConvergenceTolerance=1e-9, relaxD=relax)
    got = hostapi.evolve(m, g, PATCHES, mu, lam, rho, grav, trac, pres, fixed, v2p,
                         "nCorrectors %d; convergenceTolerance 1e-12; relConvergenceTolerance 1e-9;" % nCorr,
                         "solver gpuPCG; preconditioner DIC; tolerance 1e-09; relTol 0.01; residentMatrix %s;" % resident, relaxD=relax)
    assert got["nOuterIterations"] == ref["nOuterIterations"]
    assert got["totalPcgIterations"] == ref["totalPcgIterations"]
    # the pieces are bit-exact except the PCG reductions (re-associated): the solutions agree to ~1e-12 of their size
    # after each solve, and the outer residuals -- ratios of max-norms of increments -- to that relative to the increments
    assert np.allclose(got["residualHistory"], ref["residualHistory"], rtol=1e-6, atol=0)
    scale = np.max(np.abs(ref["D"]))
    assert np.max(np.abs(got["D"] - ref["D"])) <= 1e-9*scale
    assert np.max(np.abs(got["sigma"] - ref["sigma"])) <= 1e-7*np.max(np.abs(ref["sigma"]))
    assert got["initialResidual"] == pytest.approx(ref["initialResidual"], rel=1e-10)
