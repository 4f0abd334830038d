"""the other symmetric scalar systems that enter the same lduMatrix::solver boundary, ASSEMBLED on the
device (csrc/assembly_scalar.cuh) and solved by the same kernels:
  TEqn  src/solvers/thermalStressFoam/TEqn.H:9-15               rhoC*fvm::ddt(T) == fvm::laplacian(k, T)
  pEqn  flowModels/consistentIcoFlow/consistentIcoFlow.C:211-224  fvm::laplacian(1/interpolate(AU), p) == fvc::div(phi), setReference
Orthogonal part of the laplacian (exact on the orthogonal meshes used here; the explicit skewCorrected correction: tests/test_scalar_correction.py); material of the shipped thermal case
(run/thermalStressFoam/flange/constant/thermalProperties: C 500, k 50; rheologyProperties rho 7800; 0/T: 273 K / 573 K)."""
import numpy as np
import pytest

import oracle
from fluidstructureinteraction_b200 import cases, geometry, meshes

PT = {"fixed": "fixedDisplacement", "loaded": "tractionDisplacement", "free": "tractionDisplacement"}


def _bf(m, table, default=(0, 0.0)):
    nIF = m.nInternalFaces
    kind = np.zeros(m.nFaces - nIF, dtype=np.int32); val = np.zeros(m.nFaces - nIF)
    for p in m.patches:
        k, v = table.get(p["name"], default)
        kind[p["start"] - nIF:p["start"] - nIF + p["size"]] = k; val[p["start"] - nIF:p["start"] - nIF + p["size"]] = v
    return kind, val


def _teqn_inputs(m, g, seed=0):
    rng = np.random.default_rng(seed)
    k = 50.0*(1.0 + 0.2*rng.random(m.nCells)); rhoC = 7800.0*500.0*(1.0 + 0.1*rng.random(m.nCells))
    Told = 273.0 + 300.0*rng.random(m.nCells)
    kind, val = _bf(m, {"fixed": (1, 273.0), "loaded": (1, 573.0)})
    return dict(sign=1, bfKind=kind, bfValue=val, gammaCells=k, ddt="Euler", deltaT=10.0, rate=rhoC, psiOld=Told)


def _peqn_inputs(m, g, seed=1):
    rng = np.random.default_rng(seed)
    rAUf = 1e-3*(1.0 + 0.3*rng.random(m.nFaces))
    phi = 1e-4*rng.standard_normal(m.nFaces)
    kind, val = _bf(m, {})                                   # all-Neumann: the reference cell makes it regular
    return dict(sign=-1, bfKind=kind, bfValue=val, gammaFaces=rAUf, phi=phi, refCell=3, refValue=0.0)


def test_oracle_scalar_assembly_agrees_with_the_numpy_form_and_is_consistent():
    m = meshes.hex_box(9, 5, 6); g = geometry.compute(m)
    a = oracle.assemble_scalar(m, g, PT, **_teqn_inputs(m, g))
    # symmetric, diagonally dominant M-matrix of (rate/dt) V - laplacian: row sums = ddt diagonal + fixedValue contributions
    l, u = m.lower_upper()
    rows = a["diag"].copy(); np.add.at(rows, l, a["upper"]); np.add.at(rows, u, a["upper"])
    inp = _teqn_inputs(m, g)
    assert np.allclose(rows, inp["rate"]*g.V/inp["deltaT"], rtol=1e-12) and np.all(a["upper"] < 0)
    # a uniform gamma reproduces cases.scalar_laplacian_system (same formulas, numpy's summation order: 1e-13)
    inp = dict(inp, gammaCells=np.full(m.nCells, 50.0), rate=np.full(m.nCells, 3.9e6))
    b = oracle.assemble_scalar(m, g, PT, **inp)
    c = cases.scalar_laplacian_system(m, g, 50.0, fixedValue={"fixed": 273.0, "loaded": 573.0}, rate=3.9e6/10.0)
    assert np.allclose(b["upper"], c["upper"], rtol=1e-14) and np.allclose(b["diag"], c["diag"], rtol=1e-12)
    # pEqn form: negative definite laplacian, the reference cell doubles its diagonal, source = V*div(phi) sums to the boundary flux
    p = oracle.assemble_scalar(m, g, PT, **_peqn_inputs(m, g))
    q = oracle.assemble_scalar(m, g, PT, **dict(_peqn_inputs(m, g), refCell=-1))
    assert np.all(p["upper"] > 0) and p["diag"][3] == 2*q["diag"][3] and np.array_equal(np.delete(p["diag"], 3), np.delete(q["diag"], 3))
    phi = _peqn_inputs(m, g)["phi"]
    assert abs(q["source"].sum() - phi[m.nInternalFaces:].sum()) < 1e-15*np.abs(phi).sum()*10


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["hex", "tet"])
def test_gpu_scalar_assembly_bit_exact(kind):
    import fluidstructureinteraction_b200 as pkg
    m = meshes.hex_box(24, 10, 12) if kind == "hex" else meshes.tet_box(6, 4, 5, jitter=0.15)
    g = geometry.compute(m)
    gm = pkg.GpuMesh(m, g, PT, device=0)
    for inp in (_teqn_inputs(m, g), _peqn_inputs(m, g), dict(_teqn_inputs(m, g), ddt="steadyState", rate=None, psiOld=None)):
        a = gm.assemble_scalar(**inp)
        r = oracle.assemble_scalar(m, g, PT, **inp)
        for k in ("upper", "diag", "source", "internalCoeffs", "boundaryCoeffs"):
            assert np.array_equal(a[k], r[k]), (k, inp["sign"])
    gm.close()


@pytest.mark.gpu
def test_gpu_thermal_step_assembled_and_solved_on_the_device():
    """One implicit Euler step of the T equation of thermalStressFoam (TEqn.H:9-15) on a bar between 273 K and 573 K: the
    equation is assembled by gpupcg_assemble_scalar, fvMatrix's addBoundaryDiag / addBoundarySource applied, solved by gpuPCG,
    and compared with the oracle's assembly + PCG (iteration counts, 1e-9) -- config 4's scalar half through the GPU path."""
    import fluidstructureinteraction_b200 as pkg
    m = meshes.hex_box(40, 8, 10); g = geometry.compute(m)
    kind, val = _bf(m, {"fixed": (1, 273.0), "loaded": (1, 573.0)})
    inp = dict(sign=1, bfKind=kind, bfValue=val, gammaCells=np.full(m.nCells, 50.0), ddt="Euler", deltaT=100.0,
               rate=np.full(m.nCells, 7800.0*500.0), psiOld=np.full(m.nCells, 273.0))
    gm = pkg.GpuMesh(m, g, PT, device=0)
    l, u = m.lower_upper()
    gs = pkg.GpuPCG(l, u, m.nCells, device=0)
    fc = m.owner[m.nInternalFaces:]

    def step(asm, solve, T):
        a = asm(**dict(inp, psiOld=T))
        diag = a["diag"].copy(); b = a["source"].copy()
        np.add.at(diag, fc, a["internalCoeffs"]); np.add.at(b, fc, a["boundaryCoeffs"])
        x = T.copy()
        perf = solve(diag, a["upper"], x, b)
        return x, perf

    def gsolve(diag, upper, x, b):
        gs.set_matrix(diag, upper)
        return gs.solve(x, b, tolerance=1e-8, relTol=0.0)[0]

    def osolve(diag, upper, x, b):
        return oracle.pcg_solve(l, u, diag, upper, x, b, tolerance=1e-8, relTol=0.0)[0]

    Tg = Tc = np.full(m.nCells, 273.0)
    for _ in range(3):
        Tg, pg = step(gm.assemble_scalar, gsolve, Tg)
        Tc, pc = step(lambda **kw: oracle.assemble_scalar(m, g, PT, **kw), osolve, Tc)
        assert pg["nIterations"] == pc["nIterations"] and pg["nIterations"] > 3
        assert np.abs(Tg - Tc).max() <= 1e-9*np.abs(Tc).max()
    assert 273.0 - 1e-6 <= Tg.min() and Tg.max() <= 573.0 + 1e-6 and Tg.max() > 300.0          # heat enters from the hot end, maximum principle
    gm.close(); gs.close()


# ---------------------------------------------------------------------------------------------------------------------
# config 5's fluid half: the pressure equation of the PISO loop on the SHIPPED fluid meshes with the shipped 0/p boundary types
# ---------------------------------------------------------------------------------------------------------------------
def _shipped_fluid(name):
    import json, os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    m = meshes.PolyMesh.load_npz(os.path.join(gold, name + ".polymesh.npz"))
    with open(os.path.join(gold, name + ".case.json")) as f:
        d = json.load(f)
    # 0/p: fixedValue (and timeVaryingUniformFixedValue, a fixedValue whose value comes from a table) fix p; extrapolatedPressure
    # (fixedGradient) and symmetryPlane contribute no implicit coefficient
    table = {}
    for patch, b in d["p"]["boundaryField"].items():
        if b["type"] in ("fixedValue", "timeVaryingUniformFixedValue"):
            table[patch] = (1, 0.0 if b["type"] == "fixedValue" else 1333.2)
    sol = d["fvSolution"]["solvers"]["p"]
    return m, geometry.compute(m), table, float(sol["tolerance"]), float(sol["relTol"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["beam_fluid", "tube_fluid"])
def test_gpu_pressure_equation_on_the_shipped_fluid_meshes(name):
    """fvm::laplacian(1/interpolate(AU), p) == fvc::div(phi) (consistentIcoFlow.C:211-224) on run/fsiFoam/{beamInCrossFlow,
    3dTube}/fluid: assembled on the device bit for bit as the oracle assembles it, then solved by gpuPCG with the shipped
    tolerance (the shipped solver entry is GAMG; `solver gpuPCG; preconditioner DIC;` is what this path offers) against the
    oracle's PCG: same iteration count, 1e-9."""
    import fluidstructureinteraction_b200 as pkg
    m, g, table, tol, relTol = _shipped_fluid(name)
    assert any(k == 1 for k, _ in table.values())            # a fixedValue patch: no reference cell needed (setRefCell does nothing)
    kind, val = _bf(m, table)
    rng = np.random.default_rng(11)
    rAUf = 1e-3*(1.0 + 0.3*rng.random(m.nFaces))
    phi = 1e-6*rng.standard_normal(m.nFaces)
    pt = {p["name"]: "tractionDisplacement" for p in m.patches}          # the mesh handle's D patch table only tells `empty`
    inp = dict(sign=-1, bfKind=kind, bfValue=val, gammaFaces=rAUf, phi=phi)
    gm = pkg.GpuMesh(m, g, pt, device=0)
    a = gm.assemble_scalar(**inp)
    r = oracle.assemble_scalar(m, g, pt, **inp)
    for k in ("upper", "diag", "source", "internalCoeffs", "boundaryCoeffs"):
        assert np.array_equal(a[k], r[k]), k
    fc = m.owner[m.nInternalFaces:]
    diag = a["diag"].copy(); b = a["source"].copy()
    np.add.at(diag, fc, a["internalCoeffs"]); np.add.at(b, fc, a["boundaryCoeffs"])
    l, u = m.lower_upper()
    gs = pkg.GpuPCG(l, u, m.nCells, device=0)
    gs.set_matrix(diag, a["upper"])
    xg, xc = np.zeros(m.nCells), np.zeros(m.nCells)
    pg, hg = gs.solve(xg, b, tolerance=tol, relTol=relTol)
    pc, hc = oracle.pcg_solve(l, u, diag, a["upper"], xc, b, tolerance=tol, relTol=relTol)
    assert pg["nIterations"] == pc["nIterations"] and pg["nIterations"] > 10 and pg["converged"]
    assert np.max(np.abs(hg - hc)/np.abs(hc)) < 1e-10 + 1e-13*np.abs(hc).max()/np.abs(hc).min()
    assert np.abs(xg - xc).max() <= 1e-9*np.abs(xc).max()
    gm.close(); gs.close()
