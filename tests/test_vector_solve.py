"""GPU parity tests (-m gpu) of the component-batched solve: gpupcg_set_matrix_vector + gpupcg_solve_vector, i.e. the
component loop of fvMatrix<vector>::solve [EXT fvMatrixSolve.C] (DEqn.solve(), unsTotalLagrangianStress.C:908) in one
launch sequence.  Every component must reproduce the reference's PCG::solve on ITS system: same iteration count as the
oracle, residual history within 1e-10 relative, field within 1e-9 relative L-infinity -- whatever the other components
do (they stop at different iterations)."""
import numpy as np
import pytest

import oracle
import fluidstructureinteraction_b200 as pkg
from fluidstructureinteraction_b200 import cases, meshes
from conftest import synthetic_matrix
from test_gpu_solver import FIELD_RTOL, HIST_RTOL, hist_close, random_ldu

pytestmark = pytest.mark.gpu


def vector_system(l, u, n, nCmpt, seed):
    """Shared upper, per-component diag (addBoundaryDiag adds different internalCoeffs per component) and source."""
    diag, upper, b, x0 = synthetic_matrix(l, u, n, seed)
    rng = np.random.default_rng(seed + 100)
    dC = np.stack([diag*(1.0 + 0.05*k) + rng.uniform(0, 0.2, n)*np.abs(diag)*(rng.random(n) < 0.2) for k in range(nCmpt)], axis=1)
    bC = np.stack([b*(k + 1.0) + rng.standard_normal(n)*np.abs(b).mean()*k for k in range(nCmpt)], axis=1)
    xC = np.stack([x0*(1.0 - 0.5*k) for k in range(nCmpt)], axis=1)
    return np.ascontiguousarray(dC), upper, np.ascontiguousarray(bC), np.ascontiguousarray(xC)


def check_vector_solve(g, l, u, dC, upper, bC, x0C, tol, relTol, minIter=0, maxIter=1000, expectMode=None):
    nCmpt = dC.shape[1]
    g.set_matrix_vector(dC, upper)
    if expectMode is not None:
        assert g.vector_mode(nCmpt) == expectMode
    xg = x0C.copy()
    perfs, hists = g.solve_vector(xg, bC, tol, relTol, minIter, maxIter)
    out = []
    for k in range(nCmpt):
        xc = np.ascontiguousarray(x0C[:, k])
        pc, hc = oracle.pcg_solve(l, u, np.ascontiguousarray(dC[:, k]), upper, xc, np.ascontiguousarray(bC[:, k]), tol, relTol,
                                  minIter, maxIter)
        pg, hg = perfs[k], hists[k]
        assert pg["nIterations"] == pc["nIterations"], (k, pg, pc)
        assert pg["converged"] == pc["converged"] and pg["singular"] == pc["singular"], (k, pg, pc)
        assert pg["normFactor"] == pytest.approx(pc["normFactor"], rel=1e-13)
        assert pg["initialResidual"] == pytest.approx(pc["initialResidual"], rel=HIST_RTOL, abs=1e-300)
        if len(hc):
            assert hist_close(hg, hc, pc["initialResidual"]), k
            assert pg["finalResidual"] == pytest.approx(pc["finalResidual"], rel=HIST_RTOL)
        assert np.max(np.abs(xg[:, k] - xc)) <= FIELD_RTOL*np.max(np.abs(xc)) + 1e-300, k
        out.append((pg, pc))
    return out


@pytest.mark.parametrize("dims,nCmpt", [((1, 1, 1), 3), ((33, 1, 1), 3), ((5, 4, 3), 2), ((24, 8, 12), 3), ((40, 10, 20), 3),
                                        ((40, 10, 20), 2)])
def test_hex_vector_solve(dims, nCmpt):
    nx, ny, nz = dims
    l, u = meshes.hex_ldu(nx, ny, nz)
    n = nx*ny*nz
    dC, upper, bC, xC = vector_system(l, u, n, nCmpt, 11)
    g = pkg.GpuPCG(l, u, n)
    check_vector_solve(g, l, u, dC, upper, bC, xC, 1e-9, 0.0, expectMode=1)
    check_vector_solve(g, l, u, dC, upper, bC, xC, 1e-9, 0.01)
    g.close()


def test_components_stop_at_different_iterations():
    """One component has a zero source (converged before the first iteration), one a loose start, one runs long; plus the
    stop() edge cases per component: maxIter cap (not a multiple of the host's check batch), minIter."""
    s = cases.cantilever_system(60, 15, 30)
    l, u, diag, upper, b, n = s["l"], s["u"], s["diag"], s["upper"], s["b"], s["nCells"]
    rng = np.random.default_rng(3)
    dC = np.ascontiguousarray(np.stack([diag, diag*1.5, diag + rng.uniform(0, 1, n)*diag*(rng.random(n) < 0.1)], axis=1))
    bC = np.ascontiguousarray(np.stack([np.zeros(n), b, rng.standard_normal(n)*np.abs(b).max()], axis=1))
    x0 = np.zeros((n, 3))
    g = pkg.GpuPCG(l, u, n)
    res = check_vector_solve(g, l, u, dC, upper, bC, x0, 1e-9, 0.01, expectMode=1)
    its = [pg["nIterations"] for pg, _ in res]
    assert its[0] == 0 and len(set(its)) == 3, its
    check_vector_solve(g, l, u, dC, upper, bC, x0, 1e-30, 0.0, maxIter=13)
    check_vector_solve(g, l, u, dC, upper, bC, x0, 1e-9, 0.999999, minIter=3)
    check_vector_solve(g, l, u, dC, upper, bC, x0, 1e-9, 0.1)
    g.close()


def test_vector_solve_equals_component_solves_and_reuses_upper():
    """Against the scalar GPU path on the same handle (same kernels' arithmetic per component: iteration counts equal,
    fields equal to 1e-12), with upper=None on the second outer iteration as evolve() does."""
    s = cases.cantilever_system(40, 10, 20)
    l, u, diag, upper, b, n = s["l"], s["u"], s["diag"], s["upper"], s["b"], s["nCells"]
    dC = np.ascontiguousarray(np.stack([diag, diag*1.1, diag*1.2], axis=1))
    bC = np.ascontiguousarray(np.stack([b*0.5, b*0.25, -b], axis=1))
    g = pkg.GpuPCG(l, u, n)
    g.set_matrix_vector(dC, upper)
    xv = np.zeros((n, 3))
    pv, _ = g.solve_vector(xv, bC, 1e-9, 0.01)
    for k in range(3):
        g.set_matrix(np.ascontiguousarray(dC[:, k]), upper if k == 0 else None)
        xs = np.zeros(n)
        ps, _ = g.solve(xs, np.ascontiguousarray(bC[:, k]), 1e-9, 0.01)
        assert ps["nIterations"] == pv[k]["nIterations"]
        assert np.max(np.abs(xs - xv[:, k])) <= 1e-12*np.max(np.abs(xs))
    g.set_matrix_vector(dC*1.01, None)
    x2 = xv.copy()
    p2, _ = g.solve_vector(x2, bC, 1e-9, 0.01)
    for k in range(3):
        xc = np.ascontiguousarray(xv[:, k])
        pc, _ = oracle.pcg_solve(l, u, np.ascontiguousarray(dC[:, k]*1.01), upper, xc, np.ascontiguousarray(bC[:, k]), 1e-9, 0.01)
        assert p2[k]["nIterations"] == pc["nIterations"]
        assert np.max(np.abs(x2[:, k] - xc)) <= FIELD_RTOL*np.max(np.abs(xc))
    g.close()


def test_tet_mesh_vector_solve():
    m = meshes.tet_box(10, 4, 6, jitter=0.15)
    l, u = m.lower_upper()
    dC, upper, bC, xC = vector_system(l, u, m.nCells, 3, 21)
    g = pkg.GpuPCG(l, u, m.nCells)
    check_vector_solve(g, l, u, dC, upper, bC, xC, 1e-9, 0.0, expectMode=1)
    g.close()


def test_singular_component_and_errors():
    s = cases.cantilever_system(8, 4, 4)
    l, u, diag, upper, b, n = s["l"], s["u"], s["diag"], s["upper"], s["b"], s["nCells"]
    g = pkg.GpuPCG(l, u, n)
    with pytest.raises(pkg.GpuPcgError) as e:
        g.solve_vector(np.zeros((n, 3)), np.zeros((n, 3)))
    assert e.value.code == -5                       # solve before set_matrix_vector
    with pytest.raises(pkg.GpuPcgError) as e:
        g.set_matrix_vector(np.ones((n, 4)), upper)
    assert e.value.code == -1                       # nCmpt must be 2 or 3
    # one singular component (wApA underflows) beside two regular ones
    z = np.zeros_like(upper)
    dC = np.ones((n, 3))
    bC = np.ascontiguousarray(np.stack([np.full(n, 1e-301), np.sin(np.arange(n)), np.cos(np.arange(n))], axis=1))
    res = check_vector_solve(g, l, u, dC, z, bC, np.zeros((n, 3)), 0.0, 0.0, maxIter=4)
    assert res[0][0]["singular"] == 1 and res[0][0]["nIterations"] == 0
    g.close()


def test_polyhedral_rows_take_sequential_component_solves():
    """More than four faces per row and side: the table variant does not apply, the vector entry points run the
    components one after the other on the scalar GPU kernels (mode 0) -- same results contract."""
    l, u = random_ldu(300, 900, 4)
    dC, upper, bC, xC = vector_system(l, u, 300, 3, 31)
    g = pkg.GpuPCG(l, u, 300)
    check_vector_solve(g, l, u, dC, upper, bC, xC, 1e-9, 0.0, expectMode=0)
    g.close()


@pytest.mark.parametrize("mode", ["concurrent", "batched"])
def test_vector_solve_1M_properties(mode, monkeypatch):
    """Full benchmark size: true residual of every component of the returned field, and agreement with the scalar path --
    for the default of this size (three concurrent scalar solves, gpupcg_vector_mode 2) and for the component-batched
    kernels (GPUPCG_VECTOR_CONCURRENT=0, mode 1)."""
    if mode == "batched":
        monkeypatch.setenv("GPUPCG_VECTOR_CONCURRENT", "0")
    s = cases.cantilever_system(200, 50, 100)
    l, u, diag, upper, b, n = s["l"], s["u"], s["diag"], s["upper"], s["b"], s["nCells"]
    dC = np.ascontiguousarray(np.stack([diag, diag, diag], axis=1))
    bC = np.ascontiguousarray(np.stack([b*0.5e4, b*0.25e4, b*-1e4], axis=1))
    g = pkg.GpuPCG(l, u, n)
    g.set_matrix_vector(dC, upper)
    assert g.vector_mode(3) == (2 if mode == "concurrent" else 1)
    x = np.zeros((n, 3))
    pv, _ = g.solve_vector(x, bC, 1e-9, 0.01)
    g.set_matrix(diag, upper)
    for k in range(3):
        assert pv[k]["converged"] == 1 and 0 < pv[k]["nIterations"] < 1000
        r = bC[:, k] - g.amul(np.ascontiguousarray(x[:, k]))
        assert np.sum(np.abs(r))/pv[k]["normFactor"] == pytest.approx(pv[k]["finalResidual"], rel=1e-6)
        assert pv[k]["finalResidual"] <= 0.01*pv[k]["initialResidual"]
    xs = np.zeros(n)
    ps, _ = g.solve(xs, np.ascontiguousarray(bC[:, 2]), 1e-9, 0.01)
    assert ps["nIterations"] == pv[2]["nIterations"]
    # 150 CG iterations to relTol 0.01: the batched and the scalar kernels sum their dot products over different grids;
    # a concurrent component IS the scalar solve
    assert np.max(np.abs(xs - x[:, 2])) <= (1e-4 if mode == "batched" else 1e-11)*np.max(np.abs(xs))
    g.close()
