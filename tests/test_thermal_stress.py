"""BASELINE.json configs[3]: the thermalStressFoam sequence (src/solvers/thermalStressFoam/thermalStressFoam.C:60-80) --
every time step the T equation (TEqn.H:1-35: outer loop on the relative temperature residual, DT = T - T0), then
stress->evolve() with the thermal terms of the D equation (unsTotalLagrangianStress.C:883-890, :986-989, traction BC
tractionDisplacementFvPatchVectorField.C:233-251) -- with both equations assembled and solved on the GPU, against the oracle's
restatement of the same sequence.  Material and controls of the shipped case (run/thermalStressFoam/flange: cases.FLANGE); the
shipped mesh's patch split is not recorded (tests/test_case_formats.py), so the geometry here is a bar between the cold and the
hot face.  A third known answer pins the thermal branch from outside: free thermal expansion, D = alpha*DT*x, sigma = 0."""
import numpy as np
import pytest

import oracle
from oracle import evolve as oev
from fluidstructureinteraction_b200 import cases, geometry, meshes

F = cases.FLANGE
SMALL = 1e-15


def _bf(m, table):
    nIF = m.nInternalFaces
    kind = np.zeros(m.nFaces - nIF, dtype=np.int32); val = np.zeros(m.nFaces - nIF)
    for p in m.patches:
        k, v = table.get(p["name"], (0, 0.0))
        kind[p["start"] - nIF:p["start"] - nIF + p["size"]] = k; val[p["start"] - nIF:p["start"] - nIF + p["size"]] = v
    return kind, val


def test_oracle_free_thermal_expansion_is_exact():
    """A block on three symmetry planes, traction-free elsewhere, heated uniformly: D = alpha*DT*(x, y, z), sigma = 0 -- a linear
    field, which the least-squares interpolation and the vertex-based gradients reproduce exactly; what is left is the outer
    loop's convergence tolerance."""
    m = meshes.hex_box(6, 4, 5, L=0.6, W=0.4, H=0.5)
    g = geometry.compute(m)
    # the box's patches: fixed = x = 0, loaded = x = L, free = the other four; make x = 0 a symmetry plane and pin y, z by symmetry
    # through the exact solution's own boundary values: x = 0 symmetryDisplacement, the rest traction-free -> rigid motion in y, z is
    # removed by fixing the displacement of the patch `fixed` to the exact field instead
    pt = {"fixed": "fixedDisplacement", "loaded": "tractionDisplacement", "free": "tractionDisplacement"}
    DTv = 100.0
    mu, lam = cases.lame(F["E"], F["nu"])
    threeK = cases.three_k(F["E"], F["nu"], F["rho"])
    nIF = m.nInternalFaces
    exact = lambda X: F["alpha"]*DTv*X
    fv = np.zeros((m.nFaces - nIF, 3))
    p = m.patch("fixed")
    fv[p["start"] - nIF:p["start"] - nIF + p["size"]] = exact(g.Cf[p["start"]:p["start"] + p["size"]])
    c = oev.SolidCase(m, g, pt, mu, lam, F["rho"], fixedValue=fv, threeK=threeK, alpha=F["alpha"])
    c.DT = np.full(m.nCells, DTv); c.DTb = np.full(m.nFaces - nIF, DTv)
    r = c.evolve(1000, 1e-9, 0.0, 1e-12, 0.01, thermal=True)
    assert r["res"] <= 1e-9 and r["nOuterIterations"] < 999
    scale = np.abs(exact(g.C)).max()
    assert np.abs(c.D - exact(g.C)).max() < 1e-6*scale
    assert np.abs(c.pointD - exact(m.points)).max() < 1e-6*scale
    assert np.abs(c.sigma).max() < 1e-5*threeK*F["alpha"]*DTv            # stress-free; without the thermal terms it is O(threeK alpha DT)
    c2 = oev.SolidCase(m, g, pt, mu, lam, F["rho"], fixedValue=fv)
    c2.evolve(1000, 1e-9, 0.0, 1e-12, 0.01)
    assert np.abs(c2.sigma).max() > 1e-2*threeK*F["alpha"]*DTv


def _t_step(asm, solve, m, T, Told, kind, val):
    """TEqn.H:1-31: do { T.storePrevIter(); TEqn; solve; relResidual } while (relResidual > 1e-6 && ++iCorr < 10)."""
    fc = m.owner[m.nInternalFaces:]
    iCorr = 0
    its = []
    while True:
        Tprev = T.copy()
        a = asm(sign=1, bfKind=kind, bfValue=val, gammaCells=np.full(m.nCells, F["k"]), ddt="Euler", deltaT=F["deltaT"]*1e4,
                rate=np.full(m.nCells, F["rho"]*F["C"]), psiOld=Told)
        diag = a["diag"].copy(); b = a["source"].copy()
        np.add.at(diag, fc, a["internalCoeffs"]); np.add.at(b, fc, a["boundaryCoeffs"])
        T = T.copy()
        its.append(solve(diag, a["upper"], T, b)["nIterations"])
        maxDT = np.abs(T - Told).max()
        rel = (np.abs(T - Tprev)/(maxDT + SMALL)).max()
        iCorr += 1
        if not (rel > 1e-6 and iCorr < 10):
            break
    return T, its


@pytest.mark.gpu
def test_gpu_thermal_stress_time_steps_match_the_oracle_sequence():
    import fluidstructureinteraction_b200 as pkg
    from fluidstructureinteraction_b200.solid import GpuSolidCase
    m = meshes.hex_box(24, 6, 8); g = geometry.compute(m)
    pt = {"fixed": "fixedDisplacement", "loaded": "tractionDisplacement", "free": "tractionDisplacement"}
    nIF = m.nInternalFaces
    kind, val = _bf(m, {"fixed": (1, F["Tcold"]), "loaded": (1, F["Thot"])})
    mu, lam = cases.lame(F["E"], F["nu"])
    threeK = cases.three_k(F["E"], F["nu"], F["rho"])
    l, u = m.lower_upper()
    gm = pkg.GpuMesh(m, g, pt, device=0)
    gs = pkg.GpuPCG(l, u, m.nCells, device=0)

    def gsolve(diag, upper, x, b):
        gs.set_matrix(diag, upper)
        return gs.solve(x, b, tolerance=1e-8, relTol=0.0)[0]

    def osolve(diag, upper, x, b):
        return oracle.pcg_solve(l, u, diag, upper, x, b, tolerance=1e-8, relTol=0.0)[0]

    ref = oev.SolidCase(m, g, pt, mu, lam, F["rho"], threeK=threeK, alpha=F["alpha"])
    gc = GpuSolidCase(m, g, pt, mu, lam, F["rho"], threeK=threeK, alpha=F["alpha"])
    Tg = Tc = np.full(m.nCells, F["Tcold"])
    kw = dict(nCorrectors=F["nCorrectors"], convergenceTolerance=F["convergenceTolerance"], tolerance=1e-10, relTol=0.1,
              d2dt2="Euler", ddt="Euler", deltaT=F["deltaT"]*1e4, thermal=True, relaxD=F["relaxD"])
    for step in range(2):
        Tg_old, Tc_old = Tg, Tc
        Tg, ig = _t_step(gm.assemble_scalar, gsolve, m, Tg, Tg_old, kind, val)
        Tc, ic = _t_step(lambda **k: oracle.assemble_scalar(m, g, pt, **k), osolve, m, Tc, Tc_old, kind, val)
        assert ig == ic and np.abs(Tg - Tc).max() <= 1e-9*np.abs(Tc).max()
        # DT = T - T0 (TEqn.H:34), its boundary values: the fixedValue patches' values, the cell value on zeroGradient patches
        def dtb(T):
            b = T[m.owner[nIF:]] - F["T0"]
            b[kind == 1] = val[kind == 1] - F["T0"]
            return b
        ref.DT, ref.DTb = Tc - F["T0"], dtb(Tc)
        gc.set("DT", Tg - F["T0"]); gc.set("DTb", dtb(Tg))
        ref.new_time_step(); gc.new_time_step()
        r, rg = ref.evolve(**kw), gc.evolve(**kw)
        assert rg["nOuterIterations"] == r["nOuterIterations"] and r["nOuterIterations"] > 2
        assert np.array_equal(rg["pcgIterations"], np.array(r["pcgIterations"]))
        for k in ("D", "pointD", "sigma", "U"):
            a = getattr(ref, k); b = gc.get(k).reshape(a.shape)
            assert np.abs(a - b).max() <= 1e-8*np.abs(a).max() + 1e-300, k
    assert np.abs(ref.D).max() > 0 and np.abs(ref.sigma).max() > 0
    gm.close(); gs.close(); gc.close()
