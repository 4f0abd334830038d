"""CPU tests of the oracle itself (no GPU): the C restatement against an independent pure-Python restatement
of the same published loops, against the committed golden vectors, and basic solver properties."""
import numpy as np
import pytest

import oracle
from fluidstructureinteraction_b200 import cases, meshes
from conftest import FIXTURE_MESHES, load_fixture_mesh, load_golden, synthetic_matrix


# ---- independent, loop-for-loop Python restatement ---------------------------
def py_amul(l, u, diag, upper, x):
    Ax = [diag[c]*x[c] for c in range(len(diag))]
    for f in range(len(l)):
        Ax[u[f]] += upper[f]*x[l[f]]
        Ax[l[f]] += upper[f]*x[u[f]]
    return np.array(Ax)


def py_dic_factor(l, u, diag, upper):
    rD = [float(d) for d in diag]
    for f in range(len(l)):
        rD[u[f]] -= upper[f]*upper[f]/rD[l[f]]
    return np.array([1.0/d for d in rD])


def py_precondition(l, u, upper, rD, rA):
    wA = [rD[c]*rA[c] for c in range(len(rA))]
    for f in range(len(l)):
        wA[u[f]] -= rD[u[f]]*upper[f]*wA[l[f]]
    for f in range(len(l) - 1, -1, -1):
        wA[l[f]] -= rD[l[f]]*upper[f]*wA[u[f]]
    return np.array(wA)


def py_pcg(l, u, diag, upper, x, b, tol, relTol, minIter=0, maxIter=1000):
    n = len(x)
    x = [float(v) for v in x]

    def seqsum(vals):
        s = 0.0
        for v in vals:
            s += v
        return s

    wA = py_amul(l, u, diag, upper, x)
    rA = [b[i] - wA[i] for i in range(n)]
    xRef = seqsum(x)/n
    tmp = py_amul(l, u, diag, upper, [xRef]*n)
    normFactor = seqsum(abs(wA[i] - tmp[i]) + abs(b[i] - tmp[i]) for i in range(n)) + 1e-20
    init = seqsum(abs(r) for r in rA)/normFactor
    final = init
    nIter = 0
    hist = []

    def stop():
        if nIter < minIter:
            return False
        return nIter >= maxIter or final < tol or (relTol > 1e-15 and final <= relTol*init)

    if not stop():
        rD = py_dic_factor(l, u, diag, upper)
        wArA = 1e20
        pA = [0.0]*n
        while True:
            wArAold = wArA
            wA = py_precondition(l, u, upper, rD, rA)
            wArA = seqsum(wA[i]*rA[i] for i in range(n))
            if nIter == 0:
                pA = [float(w) for w in wA]
            else:
                beta = wArA/wArAold
                pA = [wA[i] + beta*pA[i] for i in range(n)]
            wA = py_amul(l, u, diag, upper, pA)
            wApA = seqsum(wA[i]*pA[i] for i in range(n))
            if not (abs(wApA)/normFactor > 1e-300):
                break
            alpha = wArA/wApA
            for i in range(n):
                x[i] += alpha*pA[i]
                rA[i] -= alpha*wA[i]
            final = seqsum(abs(r) for r in rA)/normFactor
            hist.append(final)
            nIter += 1
            if stop():
                break
    return np.array(x), init, final, nIter, np.array(hist)


@pytest.mark.parametrize("dims", [(1, 1, 1), (2, 1, 1), (3, 2, 1), (4, 3, 2), (5, 4, 3)])
def test_c_oracle_matches_python_restatement(dims):
    nx, ny, nz = dims
    l, u = meshes.hex_ldu(nx, ny, nz)
    n = nx*ny*nz
    diag, upper, b, x0 = synthetic_matrix(l, u, n, 7)
    xt = np.sin(0.37*np.arange(n))
    assert np.array_equal(oracle.amul(l, u, diag, upper, xt), py_amul(l, u, diag, upper, xt))
    rD = oracle.dic_factor(l, u, diag, upper)
    assert np.array_equal(rD, py_dic_factor(l, u, diag, upper))
    assert np.array_equal(oracle.precondition(l, u, upper, rD, b), py_precondition(l, u, upper, rD, b))
    x = x0.copy()
    perf, hist = oracle.pcg_solve(l, u, diag, upper, x, b, 1e-9, 0.0, maxIter=200)
    xp, init, final, nIter, hp = py_pcg(l, u, diag, upper, x0, b, 1e-9, 0.0, maxIter=200)
    assert perf["nIterations"] == nIter
    assert perf["initialResidual"] == init and perf["finalResidual"] == final
    assert np.array_equal(hist, hp)
    assert np.array_equal(x, xp)


@pytest.mark.parametrize("name", FIXTURE_MESHES)
def test_golden_vectors(name):
    """The committed golden vectors (tests/golden/make_fixtures.py) on the reference's shipped polyMeshes."""
    m = load_fixture_mesh(name)
    g = load_golden(name)
    assert m.is_upper_triangular()
    l, u = m.lower_upper()
    diag, upper, b, x0 = synthetic_matrix(l, u, m.nCells, int(g["seed"]))
    xt = np.sin(0.37*np.arange(m.nCells))
    assert np.array_equal(oracle.amul(l, u, diag, upper, xt), g["Ax"])
    rD = oracle.dic_factor(l, u, diag, upper)
    assert np.array_equal(rD, g["rD"])
    assert np.array_equal(oracle.precondition(l, u, upper, rD, b), g["wA"])
    x = x0.copy()
    perf, hist = oracle.pcg_solve(l, u, diag, upper, x, b, 1e-9, 0.0, maxIter=1000)
    assert perf["nIterations"] == int(g["nIterations"])
    assert np.array_equal(hist, g["history"])
    assert np.array_equal(x, g["x"])
    # and it is a solution: |b - Ax|_1 / normFactor below the tolerance
    r = b - oracle.amul(l, u, diag, upper, x)
    assert np.sum(np.abs(r))/perf["normFactor"] < 2e-9


def test_fixture_mesh_sizes():
    """Sizes quoted in the owner-file headers of the reference."""
    expect = {"beam_solid": (256, 640), "beam_fluid": (14592, 41600), "tube_solid": (6400, 17200),
              "tube_fluid": (16000, 45800)}
    for name, (nc, nif) in expect.items():
        m = load_fixture_mesh(name)
        assert (m.nCells, m.nInternalFaces) == (nc, nif)


def test_pcg_controls():
    s = cases.cantilever_system(12, 4, 6)
    l, u, diag, upper, b, n = s["l"], s["u"], s["diag"], s["upper"], s["b"], s["nCells"]
    # zero rhs and x0 = 0: converged on entry, no iterations (normFactor = small_)
    x = np.zeros(n)
    perf, hist = oracle.pcg_solve(l, u, diag, upper, x, np.zeros(n), 1e-9, 0.01)
    assert perf["nIterations"] == 0 and perf["converged"] == 1 and perf["initialResidual"] == 0.0
    # maxIter cap
    x = np.zeros(n)
    perf, hist = oracle.pcg_solve(l, u, diag, upper, x, b, 1e-30, 0.0, maxIter=5)
    assert perf["nIterations"] == 5 and perf["converged"] == 0 and len(hist) == 5
    # minIter forces iterations although relTol is met at once
    x = np.zeros(n)
    perf, hist = oracle.pcg_solve(l, u, diag, upper, x, b, 1e-9, 0.999999, minIter=3)
    assert perf["nIterations"] >= 3
    # relTol stops early, tolerance later
    x1, x2 = np.zeros(n), np.zeros(n)
    p1, _ = oracle.pcg_solve(l, u, diag, upper, x1, b, 1e-9, 0.1)
    p2, _ = oracle.pcg_solve(l, u, diag, upper, x2, b, 1e-9, 0.0)
    assert 0 < p1["nIterations"] < p2["nIterations"]
    assert p2["finalResidual"] < 1e-9
    # singular: wApA underflows to 0 (identity matrix, |b| ~ 1e-301, tolerance 0) -> checkSingularity breaks
    # out of the loop before x is touched
    x = np.zeros(n)
    perf, hist = oracle.pcg_solve(l, u, np.ones(n), np.zeros_like(upper), x, np.full(n, 1e-301), 0.0, 0.0)
    assert perf["singular"] == 1 and perf["nIterations"] == 0 and np.all(x == 0.0)


def test_decomposed_oracle_matches_global_solution():
    """Two sub-domains with processor interfaces: same converged solution as the undecomposed system
    (iteration counts differ legitimately: DIC is block-local)."""
    from fluidstructureinteraction_b200 import decompose
    s = cases.cantilever_system(10, 4, 6)
    n = s["nCells"]
    cellProc = (np.arange(n) % 10 >= 5).astype(np.int32)        # split at x = L/2
    doms = decompose.decompose_ldu(s["l"], s["u"], s["diag"], s["upper"], s["b"], np.zeros(n), cellProc, 2)
    perf, hist = oracle.pcg_solve_multi(doms, 1e-12, 0.0, maxIter=500)
    xg = np.zeros(n)
    perfG, _ = oracle.pcg_solve(s["l"], s["u"], s["diag"], s["upper"], xg, s["b"], 1e-12, 0.0, maxIter=500)
    xd = decompose.reconstruct(doms, n)
    assert perf["converged"] == 1 and perfG["converged"] == 1
    assert np.max(np.abs(xd - xg)) < 1e-9*np.max(np.abs(xg))
    assert perf["initialResidual"] == pytest.approx(perfG["initialResidual"], rel=1e-12)


def test_other_scalar_systems_converge_in_the_oracle():
    """TEqn-like (positive definite) and pEqn-like (negative definite, reference cell) systems."""
    from fluidstructureinteraction_b200 import cases, geometry, meshes
    m = meshes.hex_box(8, 4, 6)
    g = geometry.compute(m)
    for kw, x0 in ((dict(fixedValue={"fixed": 300.0, "loaded": 500.0}, rate=1.0e5), 300.0), (dict(refCell=0), 0.0)):
        s = cases.scalar_laplacian_system(m, g, 3.0, **kw)
        diag = s["diag"].copy(); b = s["source"].copy()
        for fc, ic, bc in s["patches"]:
            np.add.at(diag, fc, ic); np.add.at(b, fc, bc)
        if "refCell" in kw:
            rng = np.random.default_rng(2); dv = rng.standard_normal(m.nCells); b = b + (dv - dv.mean())*g.V
        else:
            b = b + 1.0e5*g.V*300.0
        x = np.full(m.nCells, x0)
        perf, hist = oracle.pcg_solve(s["l"], s["u"], diag, s["upper"], x, b, 1e-10, 0.0)
        assert perf["converged"] and not perf["singular"]
        r = b - oracle.amul(s["l"], s["u"], diag, s["upper"], x)
        assert np.sum(np.abs(r)) <= 1e-8*max(np.sum(np.abs(b)), 1e-300) + 1e-10*perf["normFactor"]
        if "refCell" not in kw:
            assert 300.0 - 1e-6 <= x.min() and x.max() <= 500.0 + 1e-6      # maximum principle of the heat equation
