"""The explicit correction of `laplacian(k,T) Gauss linear skewCorrected 1` with `snGradCorr(T) leastSquares`
(run/thermalStressFoam/flange/system/fvSchemes; numerics/skewCorrectedSnGrad/skewCorrectedSnGrad.C:57-294; [EXT]
leastSquaresGrad / leastSquaresVectors; [EXT] gaussLaplacianScheme: source -= V*div(gammaMagSf*correction)) -- what the T
equation of config 4 needs on its real, non-orthogonal mesh (flange.ans -> tests/golden/flange.polymesh.npz: 5712 cells, faces up
to 44 degrees off orthogonality).

Pin from outside the restatement: with the correction the discrete laplacian is EXACT for a linear field on every cell without
boundary faces, whatever the mesh (the corrected face gradient is g.n for a linear field); without it the error is O(1)."""
import os

import numpy as np
import pytest

import oracle
from fluidstructureinteraction_b200 import cases, geometry, meshes

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
F = cases.FLANGE


def _mesh(kind):
    if kind == "tet":
        m = meshes.tet_box(8, 5, 6, jitter=0.2)
    else:
        m = meshes.PolyMesh.load_npz(os.path.join(GOLD, "flange.polymesh.npz"))
    g = geometry.compute(m)
    return m, g, geometry.least_squares_vectors(m, g), {p["name"]: "tractionDisplacement" for p in m.patches}


def _residual(m, asm, psi):
    nIF = m.nInternalFaces
    own, nei = m.owner.astype(np.int64), m.neighbour.astype(np.int64)
    diag = asm["diag"].copy(); b = asm["source"].copy()
    np.add.at(diag, own[nIF:], asm["internalCoeffs"]); np.add.at(b, own[nIF:], asm["boundaryCoeffs"])
    Ax = diag*psi
    np.add.at(Ax, own[:nIF], asm["upper"]*psi[nei]); np.add.at(Ax, nei, asm["upper"]*psi[own[:nIF]])
    return np.abs(Ax - b)/np.abs(diag*psi), diag, b


@pytest.mark.parametrize("kind", ["tet", "flange"])
def test_oracle_corrected_laplacian_is_exact_for_linear_fields(kind):
    m, g, ls, pt = _mesh(kind)
    nIF = m.nInternalFaces
    nBF = m.nFaces - nIF
    a = np.array([0.7, -1.3, 2.1])
    L = np.abs(g.C).max()
    psi = g.C@a/L + 0.5; psib = g.Cf[nIF:]@a/L + 0.5
    interior = np.ones(m.nCells, bool); interior[m.owner[nIF:]] = False
    assert interior.sum() > 100
    kind1 = np.ones(nBF, dtype=np.int32)
    gam = np.full(m.nCells, 2.0)
    r0, _, _ = _residual(m, oracle.assemble_scalar(m, g, pt, 1, kind1, psib, gammaCells=gam), psi)
    r1, _, _ = _residual(m, oracle.assemble_scalar(m, g, pt, 1, kind1, psib, gammaCells=gam, psi=psi, ls=ls, limitCoeff=1.0), psi)
    assert r1[interior].max() < 1e-12 and r0[interior].max() > 1e-3
    # the same through the pEqn form (laplacian == div(phi), sign -1)
    r2, _, _ = _residual(m, oracle.assemble_scalar(m, g, pt, -1, kind1, psib, gammaFaces=np.full(m.nFaces, 2.0), psi=psi, ls=ls), psi)
    assert r2[interior].max() < 1e-12


def _fields(m, g, seed=3):
    rng = np.random.default_rng(seed)
    nIF = m.nInternalFaces
    nBF = m.nFaces - nIF
    L = np.abs(g.C).max()
    psi = 300.0 + 50.0*np.sin(3.0*g.C[:, 0]/L)*np.cos(2.0*g.C[:, 2]/L) + rng.standard_normal(m.nCells)
    zf = g.Cf[nIF:, 2]
    kind = np.where((zf < zf.min() + 0.02*np.ptp(zf)) | (zf > zf.max() - 0.02*np.ptp(zf)), 1, 0).astype(np.int32)
    val = np.where(zf > zf.mean(), 573.0, 273.0)
    return psi, kind, val, rng


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["tet", "flange"])
def test_gpu_scalar_correction_bit_exact(kind):
    import fluidstructureinteraction_b200 as pkg
    m, g, ls, pt = _mesh(kind)
    psi, bk, bv, rng = _fields(m, g)
    gm = pkg.GpuMesh(m, g, pt, device=0)
    rhoC = np.full(m.nCells, F["rho"]*F["C"]); kT = F["k"]*(1.0 + 0.2*rng.random(m.nCells))
    for lim in (1.0, 0.5):
        inp = dict(sign=1, bfKind=bk, bfValue=bv, gammaCells=kT, ddt="Euler", deltaT=0.5, rate=rhoC, psiOld=psi - 1.0, psi=psi, ls=ls,
                   limitCoeff=lim)
        a, r = gm.assemble_scalar(**inp), oracle.assemble_scalar(m, g, pt, **inp)
        for k in ("upper", "diag", "source", "internalCoeffs", "boundaryCoeffs"):
            assert np.array_equal(a[k], r[k]), (lim, k)
        r0 = oracle.assemble_scalar(m, g, pt, **{k: v for k, v in inp.items() if k not in ("psi", "ls", "limitCoeff")})
        assert np.abs(r["source"] - r0["source"]).max() > 1e-6*np.abs(r0["source"]).max()          # the correction is there
    inp = dict(sign=-1, bfKind=bk, bfValue=bv, gammaFaces=1e-3*(1.0 + 0.3*rng.random(m.nFaces)), phi=1e-6*rng.standard_normal(m.nFaces),
               psi=psi, ls=ls)
    a, r = gm.assemble_scalar(**inp), oracle.assemble_scalar(m, g, pt, **inp)
    for k in ("upper", "diag", "source", "internalCoeffs", "boundaryCoeffs"):
        assert np.array_equal(a[k], r[k]), k
    gm.close()


@pytest.mark.gpu
def test_gpu_T_equation_with_the_correction_on_the_flange_mesh():
    """TEqn.H:1-31 on the mesh of config 4 with its schemes: the outer loop re-assembles the explicit correction from the latest T
    and solves with gpuPCG -- against the oracle's sequence: same PCG iteration counts, T within 1e-9; the corrected steady field
    differs visibly from the orthogonal-only one (the mesh is up to 44 degrees non-orthogonal)."""
    import fluidstructureinteraction_b200 as pkg
    m, g, ls, pt = _mesh("flange")
    _, bk, bv, _ = _fields(m, g)
    nIF = m.nInternalFaces
    fc = m.owner[nIF:].astype(np.int64)
    l, u = m.lower_upper()
    gm = pkg.GpuMesh(m, g, pt, device=0)
    gs = pkg.GpuPCG(l, u, m.nCells, device=0)
    kT = np.full(m.nCells, F["k"])

    def run(asm, solve, corrected, nOuter=8):
        T = np.full(m.nCells, 273.0)
        its = []
        for _ in range(nOuter):
            kw = dict(psi=T, ls=ls, limitCoeff=1.0) if corrected else {}
            a = asm(sign=1, bfKind=bk, bfValue=bv, gammaCells=kT, **kw)
            diag = a["diag"].copy(); b = a["source"].copy()
            np.add.at(diag, fc, a["internalCoeffs"]); np.add.at(b, fc, a["boundaryCoeffs"])
            its.append(solve(diag, a["upper"], T, b))
        return T, its

    def solve_gpu(diag, upper, T, b):
        gs.set_matrix(diag, upper)
        p, _ = gs.solve(T, b, tolerance=1e-10, relTol=0.0)
        return p["nIterations"]

    def solve_cpu(diag, upper, T, b):
        p, _ = oracle.pcg_solve(l, u, diag, upper, T, b, tolerance=1e-10, relTol=0.0)
        return p["nIterations"]

    Tg, ig = run(gm.assemble_scalar, solve_gpu, True)
    Tc, ic = run(lambda **kw: oracle.assemble_scalar(m, g, pt, **kw), solve_cpu, True)
    assert ig == ic and ig[0] > 10
    assert np.abs(Tg - Tc).max() <= 1e-9*np.abs(Tc).max()
    T0, _ = run(lambda **kw: oracle.assemble_scalar(m, g, pt, **kw), solve_cpu, False, nOuter=1)
    assert 273.0 - 1e-6 <= Tc.min() and Tc.max() <= 573.0 + 30.0
    assert np.abs(Tc - T0).max() > 1.0                      # kelvin: the non-orthogonal correction matters on this mesh
    gm.close(); gs.close()


def _flange_case():
    m, g, kind, val, _ = cases.flange_case(os.path.join(GOLD, "flange.polymesh.npz"))
    return m, g, kind, val


def test_flange_boundary_surfaces_give_the_shipped_T_patches():
    m, g, kind, val = _flange_case()
    assert [p["size"] for p in m.patches] == [348, 384, 288, 3268 - 348 - 384 - 288]
    assert (kind == 1).sum() == 348 + 384 and (val[kind == 1] == F["Thot"]).sum() == 384
    assert m.is_upper_triangular() and np.all(g.V > 0)


@pytest.mark.gpu
def test_gpu_config_4_sequence_on_the_flange_mesh(monkeypatch):
    """thermalStressFoam.C:60-80 on the real mesh of the shipped case: every time step the T equation (TEqn.H, Euler ddt,
    `laplacian(k,T) Gauss linear skewCorrected 1` with its least-squares correction, shipped T boundary conditions), then
    stress->evolve() with the thermal terms (Euler d2dt2 / ddt, skewCorrected 1, relaxation 0.5) -- both assembled and solved on
    the GPU, against the oracle's restatement of the same sequence.

    With the global sums formed in the reference's order (GPUPCG_ORDERED_SUMS) the whole chain is the oracle's: identical PCG
    iteration counts in all 450 solves of an evolve() and fields equal to rounding.  (With the production tree reductions one
    solve in ~450 stops an iteration earlier or later on this case: late in the relaxed outer loop the initial residual of a
    solve is a difference of nearly equal numbers, and the oracle's own stopping margins there are 1e-6 ... 1e-3.)"""
    import fluidstructureinteraction_b200 as pkg
    from oracle import evolve as oev
    from fluidstructureinteraction_b200.solid import GpuSolidCase
    monkeypatch.setenv("GPUPCG_ORDERED_SUMS", "1")
    m, g, kind, val = _flange_case()
    ls = geometry.least_squares_vectors(m, g)
    pt = {"cold": "tractionDisplacement", "hot": "tractionDisplacement", "fixed": "fixedDisplacement", "free": "tractionDisplacement"}
    nIF = m.nInternalFaces
    fc = m.owner[nIF:].astype(np.int64)
    l, u = m.lower_upper()
    mu, lam = cases.lame(F["E"], F["nu"])
    threeK = cases.three_k(F["E"], F["nu"], F["rho"])
    gm = pkg.GpuMesh(m, g, pt, device=0)
    gs = pkg.GpuPCG(l, u, m.nCells, device=0)
    dT = 2.0e4           # the shipped mesh is in millimetres read as metres: a time step on its diffusion time scale
    rhoC = np.full(m.nCells, F["rho"]*F["C"]); kT = np.full(m.nCells, F["k"])

    def t_step(asm, solve, T, Told):
        """TEqn.H:1-31."""
        iCorr, its = 0, []
        while True:
            Tprev = T.copy()
            a = asm(sign=1, bfKind=kind, bfValue=val, gammaCells=kT, ddt="Euler", deltaT=dT, rate=rhoC, psiOld=Told, psi=T, ls=ls,
                    limitCoeff=1.0)
            diag = a["diag"].copy(); b = a["source"].copy()
            np.add.at(diag, fc, a["internalCoeffs"]); np.add.at(b, fc, a["boundaryCoeffs"])
            T = T.copy()
            its.append(solve(diag, a["upper"], T, b))
            rel = (np.abs(T - Tprev)/(np.abs(T - Told).max() + 1e-15)).max()
            iCorr += 1
            if not (rel > 1e-6 and iCorr < 10):
                break
        return T, its

    def gsolve(diag, upper, x, b):
        gs.set_matrix(diag, upper)
        return gs.solve(x, b, tolerance=1e-8, relTol=0.0)[0]["nIterations"]

    def osolve(diag, upper, x, b):
        return oracle.pcg_solve(l, u, diag, upper, x, b, tolerance=1e-8, relTol=0.0)[0]["nIterations"]

    ref = oev.SolidCase(m, g, pt, mu, lam, F["rho"], threeK=threeK, alpha=F["alpha"])
    gc = GpuSolidCase(m, g, pt, mu, lam, F["rho"], threeK=threeK, alpha=F["alpha"])
    Tg = Tc = np.full(m.nCells, F["Tcold"])
    kw = dict(nCorrectors=150, convergenceTolerance=F["convergenceTolerance"], tolerance=1e-10, relTol=0.1, d2dt2="Euler", ddt="Euler",
              deltaT=dT, thermal=True, relaxD=F["relaxD"])
    for step in range(2):
        Tg_old, Tc_old = Tg, Tc
        Tg, ig = t_step(gm.assemble_scalar, gsolve, Tg, Tg_old)
        Tc, ic = t_step(lambda **k: oracle.assemble_scalar(m, g, pt, **k), osolve, Tc, Tc_old)
        assert ig == ic and len(ic) > 2 and np.abs(Tg - Tc).max() <= 1e-9*np.abs(Tc).max()

        def dtb(T):
            b = T[fc] - F["T0"]
            b[kind == 1] = val[kind == 1] - F["T0"]
            return b
        # the D equation of both sides from the SAME temperature (the T fields agree to 1e-9, asserted above; 150 relaxed outer
        # iterations on this mesh would carry that difference into the PCG iteration counts)
        ref.DT, ref.DTb = Tg - F["T0"], dtb(Tg)
        gc.set("DT", Tg - F["T0"]); gc.set("DTb", dtb(Tg))
        ref.new_time_step(); gc.new_time_step()
        r, rg = ref.evolve(**kw), gc.evolve(**kw)
        assert rg["nOuterIterations"] == r["nOuterIterations"] and r["nOuterIterations"] > 10
        pi, pg = np.array(r["pcgIterations"]), rg["pcgIterations"]
        assert np.array_equal(pi, pg), (step, np.argwhere(pi != pg)[:5], np.array(r["pcgStopMargin"])[pi != pg][:5])
        assert np.array_equal(r["residualHistory"], rg["residualHistory"])
        for k in ("D", "pointD", "sigma", "U"):
            a = getattr(ref, k); b = gc.get(k).reshape(a.shape)
            assert np.abs(a - b).max() <= 1e-12*np.abs(a).max() + 1e-300, (step, k)
    assert Tc.max() > 400.0 and np.abs(ref.D).max() > 0 and np.abs(ref.sigma).max() > 0
    gm.close(); gs.close(); gc.close()


def test_oracle_steady_conduction_on_a_skewed_mesh_recovers_the_linear_profile():
    """Known answer for the corrected T equation as a whole (deferred correction iterated as TEqn.H's loop does): steady
    conduction between a cold and a hot wall, insulated elsewhere, has the linear profile whatever the mesh.  On jittered tets the
    orthogonal-only laplacian misses it by several kelvin; with the skewCorrected correction the error drops by an order of
    magnitude (what is left sits in the boundary cells, whose wall faces carry no correction in the reference either)."""
    m = meshes.tet_box(10, 4, 4, L=2.0, W=0.5, H=0.5, jitter=0.2)
    g = geometry.compute(m)
    ls = geometry.least_squares_vectors(m, g)
    pt = {p["name"]: "tractionDisplacement" for p in m.patches}
    nIF = m.nInternalFaces
    xf = g.Cf[nIF:, 0]
    kind = np.where((xf < 1e-9) | (xf > 2.0 - 1e-9), 1, 0).astype(np.int32)
    val = np.where(xf > 1.0, 573.0, 273.0)
    exact = 273.0 + 300.0*g.C[:, 0]/2.0
    fc = m.owner[nIF:].astype(np.int64)
    l, u = m.lower_upper()
    kT = np.full(m.nCells, F["k"])

    def solve(corrected, nOuter):
        T = np.full(m.nCells, 273.0)
        for _ in range(nOuter):
            kw = dict(psi=T, ls=ls, limitCoeff=1.0) if corrected else {}
            a = oracle.assemble_scalar(m, g, pt, sign=1, bfKind=kind, bfValue=val, gammaCells=kT, **kw)
            diag = a["diag"].copy(); b = a["source"].copy()
            np.add.at(diag, fc, a["internalCoeffs"]); np.add.at(b, fc, a["boundaryCoeffs"])
            oracle.pcg_solve(l, u, diag, a["upper"], T, b, tolerance=1e-12, relTol=0.0)
        return T

    e0 = np.abs(solve(False, 1) - exact).max()
    e1 = np.abs(solve(True, 25) - exact).max()
    assert e0 > 2.0 and e1 < 0.25*e0, (e0, e1)
