"""The C++ host side: foam-extend API stand-in + the gpuPCG adaptor source + segregated fvMatrix solve, driven
through fvSolution-style dictionaries (the way `DEqn.solve()` of the reference reaches the solver)."""
import numpy as np
import pytest

import oracle
from fluidstructureinteraction_b200 import cases, hostapi

FVSOLUTION_D = """
    solver          gpuPCG;
    preconditioner  DIC;        // run/stressFoam/plateHole/plateHole/system/fvSolution:18-27
    tolerance       1e-09;
    relTol          0.01;
"""


def vector_system(nx, ny, nz):
    """Synthetic fvVectorMatrix on the hex cantilever: shared upper, per-component boundary diag and source."""
    s = cases.cantilever_system(nx, ny, nz)
    n = s["nCells"]
    rng = np.random.default_rng(3)
    i = np.arange(n) % nx
    fixed = np.flatnonzero(i == 0).astype(np.int32)
    loaded = np.flatnonzero(i == nx - 1).astype(np.int32)
    diag0 = s["diag"].copy()
    diag0[fixed] -= diag0[fixed]*0.0        # boundary diag comes through internalCoeffs below
    gam = s["gamma"]
    hx, hy, hz = 2.0/nx, 0.5/ny, 1.0/nz
    coefB = gam*(hy*hz)/(0.5*hx)
    base = np.bincount(s["l"], weights=-s["upper"], minlength=n) + np.bincount(s["u"], weights=-s["upper"], minlength=n)
    patches = [
        (fixed, np.tile([coefB, coefB*1.5, coefB*2.0], (fixed.size, 1)), rng.standard_normal((fixed.size, 3))*1e3),
        (loaded, np.zeros((loaded.size, 3)), np.tile([0.5e4, 0.25e4, -1e4], (loaded.size, 1))*hy*hz),
    ]
    source = rng.standard_normal((n, 3))*10.0
    return s["l"], s["u"], base, s["upper"], source, patches


def reference_segregated(l, u, diag, upper, source, patches, tol, relTol):
    """Python restatement of fvMatrix<vector>::solve over the CPU oracle's PCG."""
    n = diag.shape[0]
    total = source.copy()
    for fc, ic, bc in patches:
        np.add.at(total, fc, bc)
    psi = np.zeros((n, 3))
    perfs = []
    for c in range(3):
        d = diag.copy()
        for fc, ic, bc in patches:
            np.add.at(d, fc, ic[:, c])
        x = np.zeros(n)
        p, h = oracle.pcg_solve(l, u, d, upper, x, np.ascontiguousarray(total[:, c]), tol, relTol)
        psi[:, c] = x
        perfs.append(p)
    return psi, perfs


def test_dictionary_and_selection_errors():
    assert hostapi.dict_lookup(FVSOLUTION_D, "solver") == "gpuPCG"
    assert hostapi.dict_lookup("solvers { D { solver gpuPCG; relTol 0.1; } }", "solvers/D/relTol") == "0.1"
    l, u, diag, upper, source, patches = vector_system(4, 2, 2)
    psi = np.zeros_like(source)
    with pytest.raises(hostapi.FoamFatalError) as e:
        hostapi.solve("D", l, u, diag, upper, source, patches, "solver PCG; preconditioner DIC;", psi)
    assert "Unknown symmetric matrix solver PCG" in str(e.value) and "gpuPCG" in str(e.value)
    with pytest.raises(hostapi.FoamFatalError) as e:
        hostapi.solve("D", l, u, diag, upper, source, patches, "solver gpuPCG; preconditioner GAMG;", psi)
    assert "DIC" in str(e.value)
    with pytest.raises(hostapi.FoamFatalError) as e:
        hostapi.solve("D", l, u, diag, upper, source, patches, "preconditioner DIC;", psi)
    assert "keyword solver is undefined" in str(e.value)


def test_adaptor_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    l, u, diag, upper, source, patches = vector_system(4, 2, 2)
    with pytest.raises(hostapi.FoamFatalError) as e:
        hostapi.solve("D", l, u, diag, upper, source, patches, FVSOLUTION_D, np.zeros_like(source))
    assert "gpupcg_create failed" in str(e.value) and "no CPU fallback" in str(e.value)


@pytest.mark.gpu
@pytest.mark.parametrize("batch", ["yes", "no"])
def test_segregated_vector_solve_matches_reference_flow(batch):
    """DEqn.solve() through the C++ host side: `batchComponents yes` (default) hands the three components to
    gpuPCG::solveVector (one device launch sequence), `no` runs fvMatrix<vector>::solve's component loop over
    gpuPCG::solve -- both must reproduce the reference flow component by component."""
    l, u, diag, upper, source, patches = vector_system(30, 8, 12)
    psi = np.zeros_like(source)
    worst, per = hostapi.solve("D", l, u, diag, upper, source, patches, FVSOLUTION_D + "    batchComponents %s;\n" % batch, psi)
    ref, perfs = reference_segregated(l, u, diag, upper, source, patches, 1e-9, 0.01)
    assert [p["fieldName"] for p in per] == ["Dx", "Dy", "Dz"]
    assert all(p["solverName"] == "DICgpuPCG" for p in per)
    for p, q in zip(per, perfs):
        assert p["nIterations"] == q["nIterations"] and p["converged"] == q["converged"]
        assert p["initialResidual"] == pytest.approx(q["initialResidual"], rel=1e-10)
        assert p["finalResidual"] == pytest.approx(q["finalResidual"], rel=1e-9)
    assert np.max(np.abs(psi - ref)) <= 1e-9*np.max(np.abs(ref))
    assert worst["initialResidual"] == max(p["initialResidual"] for p in per)


@pytest.mark.gpu
def test_scalar_solve_T_equation_style():
    """fvScalarMatrix::solve (TEqn.solve(), thermalStressFoam/TEqn.H:15): one component, fieldName 'T'."""
    l, u, diag, upper, source, patches = vector_system(16, 6, 6)
    patches1 = [(fc, ic[:, 0].copy(), bc[:, 0].copy()) for fc, ic, bc in patches]
    src = np.ascontiguousarray(source[:, 0])
    psi = np.zeros_like(src)
    worst, per = hostapi.solve("T", l, u, diag, upper, src, patches1, "solver gpuPCG; preconditioner DIC; tolerance 1e-9; relTol 0;", psi)
    assert len(per) == 1 and per[0]["fieldName"] == "T"
    d = diag.copy(); tot = src.copy()
    for fc, ic, bc in patches1:
        np.add.at(d, fc, ic); np.add.at(tot, fc, bc)
    x = np.zeros_like(src)
    p, h = oracle.pcg_solve(l, u, d, upper, x, tot, 1e-9, 0.0)
    assert per[0]["nIterations"] == p["nIterations"]
    assert np.max(np.abs(psi - x)) <= 1e-9*np.max(np.abs(x))


@pytest.mark.gpu
@pytest.mark.parametrize("reuse", ["yes", "no"])
def test_reuse_upper_notices_a_change_at_a_single_face(reuse):
    """a fresh fvMatrix routinely lands on the storage of the previous one, so the address of upper() says
    nothing.  One coefficient changed in place (a face no strided sample would hit) must reach the device: the second
    solve has to be that of the modified matrix, with `reuseUpper yes` (full hash) as with the default `no`."""
    s = cases.cantilever_system(24, 8, 10)
    l, u, diag, upper, n = s["l"], s["u"], s["diag"], s["upper"], s["nCells"]
    b = 1e4*np.sin(0.37*np.arange(n))            # the cantilever's own source gives a solution that varies along x only
    f = 1237                                    # not a multiple of any small stride
    delta = 0.37*upper[f]
    upper2 = upper.copy(); upper2[f] += delta
    diag2 = diag.copy(); diag2[l[f]] -= delta; diag2[u[f]] -= delta       # keep the row sums (negSumDiag)
    d = "solver gpuPCG; preconditioner DIC; tolerance 1e-12; relTol 0; reuseUpper %s;" % reuse
    p1, p2 = hostapi.solve_twice(l, u, diag, diag2, upper, f, delta, b, d)
    x1, x2 = np.zeros(n), np.zeros(n)
    oracle.pcg_solve(l, u, diag, upper, x1, b, 1e-12, 0.0)
    oracle.pcg_solve(l, u, diag2, upper2, x2, b, 1e-12, 0.0)
    assert np.max(np.abs(p1 - x1)) <= 1e-9*np.max(np.abs(x1))
    assert np.max(np.abs(p2 - x2)) <= 1e-9*np.max(np.abs(x2))
    assert np.max(np.abs(x2 - x1)) > 1e-6*np.max(np.abs(x1))            # the change matters
