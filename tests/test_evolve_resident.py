"""The complete outer loop of unsTotalLagrangianStress::evolve() (unsTotalLagrangianStress.C:769-1009) resident on the
GPU -- every branch (Euler d2dt2 / damping / nonLinear / thermal / 2-D), the reference's least-squares vol->point
interpolation, D.boundaryField() tracking -- against the oracle's restatement of the same loop (oracle/evolve.py
SolidCase).  Kernels are compared bit for bit; whole loops by iteration counts + fields at the north_star bars."""
import numpy as np
import pytest

import oracle
from oracle import evolve as oev
from fluidstructureinteraction_b200 import cases, geometry, lsq, meshes

PT3 = {"fixed": "fixedDisplacement", "loaded": "tractionDisplacement", "free": "tractionDisplacement"}
PT2D = {"fixed": "fixedDisplacement", "loaded": "tractionDisplacement", "sides": "tractionDisplacement", "frontAndBack": "empty"}


def _mesh(kind):
    if kind == "hex":
        m = meshes.hex_box(10, 4, 6); pt = dict(PT3)
    elif kind == "hexsym":
        m = meshes.hex_box(8, 4, 5); pt = dict(PT3, free="symmetryDisplacement")
    elif kind == "tet":
        m = meshes.tet_box(5, 3, 4, jitter=0.15); pt = dict(PT3)
    elif kind == "2d":
        m = meshes.hex_box(12, 6, 1, split_free=True); pt = dict(PT2D)
    elif kind == "flange":          # the imported mesh of config 4: hexes and prisms, faces up to 44 degrees off orthogonality
        import os
        m, g, _, _, pt = cases.flange_case(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "flange.polymesh.npz"))
        return m, g, pt
    else:
        raise KeyError(kind)
    return m, geometry.compute(m), pt


# ---------------------------------------------------------------------------------------------- CPU: oracle and host logic
@pytest.mark.parametrize("kind", ["hex", "tet", "2d"])
def test_lsq_builder_matches_oracle_and_is_exact_on_linear_fields(kind):
    m, g, pt = _mesh(kind)
    L = oracle.Lsq(m, g, pt)
    e, b = L.export(), lsq.build(m, g, pt)
    for k in e:
        assert np.array_equal(e[k], b[k]), k
    a = np.array([0.1, -0.2, 0.3]); B = np.array([[1, 2, 3], [4, 5, 6], [7, 8, 10.0]])*1e-3
    if kind == "2d":
        a[2] = 0.0; B[2, :] = 0.0; B[:, 2] = 0.0
    nIF = m.nInternalFaces
    pD = L.interpolate(a + g.C@B.T, a + g.Cf[nIF:]@B.T)
    assert np.abs(pD - (a + m.points@B.T)).max() < 1e-14


def test_extended_assembly_reduces_to_the_round1_one():
    m, g, pt = _mesh("hexsym")
    rng = np.random.default_rng(3)
    nC, nF, nBF = m.nCells, m.nFaces, m.nFaces - m.nInternalFaces
    mu, lam = cases.lame(2e11, 0.3)
    mu, lam, rho = np.full(nC, mu), np.full(nC, lam), np.full(nC, 7854.0)
    D = 1e-5*rng.standard_normal((nC, 3)); G = 1e-5*rng.standard_normal((nC, 9)); Gf = 1e-5*rng.standard_normal((nF, 9))
    tr = 1e4*rng.standard_normal((nBF, 3)); pr = 1e3*rng.standard_normal(nBF); fv = 1e-6*rng.standard_normal((nBF, 3))
    gv = np.array([0.0, 0.0, -9.81])
    a = oracle.assemble_D(m, g, pt, mu, lam, rho, gv, D, G, Gf, tr, pr, fv, 0.7)
    b = oracle.assemble_D_ex(m, g, pt, oracle.controls(limitCoeff=0.7), mu, lam, rho, gv, D, D, D, G, Gf, tr, pr, fv)
    for k in a:
        assert np.array_equal(a[k], b[k]), k


def test_oracle_transient_nonlinear_loop_runs_two_steps():
    m, g, pt = _mesh("hex")
    mu, lam = cases.lame(1e4, 0.4)
    nBF, nIF = m.nFaces - m.nInternalFaces, m.nInternalFaces
    tr = np.zeros((nBF, 3)); p = m.patch("loaded"); tr[p["start"] - nIF:p["start"] - nIF + p["size"], 2] = -5.0
    c = oev.SolidCase(m, g, pt, mu, lam, 1000.0, traction=tr)
    its = []
    for _ in range(2):
        c.new_time_step()
        r = c.evolve(300, 1e-7, 1e-3, relTol=0.1, d2dt2="Euler", ddt="Euler", deltaT=0.01, nonLinear=True, relaxD=0.5)
        its.append(r["nOuterIterations"])
        assert r["res"] <= max(1e-7, 1e-3*r["maxRes"]) and not r["enforceLinear"]
    assert np.abs(c.D).max() > 0 and np.abs(c.U).max() > 0 and its[0] > 1


# ---------------------------------------------------------------------------------------------- GPU
def _random_state(m, seed=0, scale=1e-5):
    rng = np.random.default_rng(seed)
    nC, nF, nBF = m.nCells, m.nFaces, m.nFaces - m.nInternalFaces
    return dict(D=scale*rng.standard_normal((nC, 3)), Dold=scale*rng.standard_normal((nC, 3)),
                Doldold=scale*rng.standard_normal((nC, 3)), gradD=scale*rng.standard_normal((nC, 9)),
                gradDf=scale*rng.standard_normal((nF, 9)), traction=1e4*rng.standard_normal((nBF, 3)),
                pressure=1e3*rng.standard_normal(nBF), fixedValue=1e-6*rng.standard_normal((nBF, 3)),
                threeK=5e11*(1 + 0.1*rng.random(nC)), alpha=1e-5*(1 + 0.1*rng.random(nC)), DT=50*rng.standard_normal(nC),
                DTb=50*rng.standard_normal(nBF))


BRANCHES = [dict(), dict(d2dt2="Euler", ddt="Euler", deltaT=0.01, deltaT0=0.02), dict(ddt="Euler", deltaT=0.01, K=3.0),
            dict(nonLinear=True), dict(thermal=True),
            dict(d2dt2="Euler", ddt="Euler", deltaT=0.01, K=2.0, nonLinear=True, thermal=True, limitCoeff=0.5)]


def _check_resident_chain(m, g, pt, branch, scale, lsq_data=None):
    """assemble -> D.correctBoundaryConditions -> vol->point -> grad / fGrad -> sigma / U on the device, every array bit for
    bit against the oracle on the same state."""
    import fluidstructureinteraction_b200 as pkg
    from fluidstructureinteraction_b200 import capi
    st = _random_state(m, seed=branch, scale=scale)
    nC = m.nCells
    mu, lam = cases.lame(2e11, 0.3)
    mu, lam, rho = np.full(nC, mu), np.full(nC, lam), np.full(nC, 7854.0)
    gv = np.array([0.0, 0.3, -9.81])
    kw = BRANCHES[branch]
    ctl = oracle.controls(**kw)
    ref = oracle.assemble_D_ex(m, g, pt, ctl, mu, lam, rho, gv, st["D"], st["Dold"], st["Doldold"], st["gradD"], st["gradDf"],
                               st["traction"], st["pressure"], st["fixedValue"], st["threeK"], st["alpha"], st["DT"], st["DTb"])
    gm = pkg.GpuMesh(m, g, pt, device=0)
    for k, v in (("mu", mu), ("lambda_", lam), ("rho", rho)):
        gm.set_field(k, v)
    for k in ("D", "Dold", "Doldold", "gradD", "gradDf", "traction", "pressure", "fixedValue", "threeK", "alpha", "DT", "DTb"):
        gm.set_field(k, st[k])
    gm.assemble_resident(capi.eqn_controls(**kw), gv)
    for k in ("upper", "diag", "source", "internalCoeffs", "boundaryCoeffs"):
        assert np.array_equal(gm.get_field(k), ref[k].reshape(gm.get_field(k).shape)), (k, kw)
    nonempty = np.ones(m.nFaces - m.nInternalFaces, dtype=bool)
    for p in m.patches:
        if pt[p["name"]] == "empty":
            nonempty[p["start"] - m.nInternalFaces:p["start"] - m.nInternalFaces + p["size"]] = False
    assert np.array_equal(gm.get_field("patchGradient")[nonempty], ref["patchGradient"][nonempty])
    # boundary evaluate, least-squares interpolation, gradients, stress on the same state
    gm.patch_evaluate()
    Db = oracle.patch_evaluate(m, g, pt, st["D"], st["gradD"], st["fixedValue"], ref["patchGradient"])
    assert np.array_equal(gm.get_field("Db")[nonempty], Db[nonempty])
    gm.set_field("Db", Db)
    L = oracle.Lsq(m, g, pt)
    gm.set_lsq(lsq_data(L) if lsq_data else lsq.build(m, g, pt))
    gm.vol_to_point()
    pD = L.interpolate(st["D"], Db)
    assert np.array_equal(gm.get_field("pointD"), pD)
    gm.gradients_on_device(ctl.limitCoeff)
    G2, Gb2 = oracle.grad_ex(m, g, pt, st["D"], pD, st["gradD"], st["fixedValue"], ref["patchGradient"])
    Gf2 = oracle.fgrad_ex(m, g, pt, st["D"], pD, G2, st["fixedValue"], ref["patchGradient"], ctl.limitCoeff)
    assert np.array_equal(gm.get_field("gradD"), G2)
    assert np.array_equal(gm.get_field("gradDb")[nonempty], Gb2[nonempty])
    nef = np.concatenate([np.ones(m.nInternalFaces, dtype=bool), nonempty])
    assert np.array_equal(gm.get_field("gradDf")[nef], Gf2[nef])
    gm.sigma_resident(capi.eqn_controls(**kw))
    eps, sig = oracle.sigma_ex(mu, lam, G2, kw.get("nonLinear", False), kw.get("thermal", False), st["threeK"], st["alpha"], st["DT"])
    assert np.array_equal(gm.get_field("epsilon"), eps) and np.array_equal(gm.get_field("sigma"), sig)
    assert np.array_equal(gm.get_field("U"), oracle.ddt(st["D"], st["Dold"], kw.get("ddt", "steadyState"), kw.get("deltaT", 1.0)))
    gm.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["hexsym", "tet", "2d", "flange"])
@pytest.mark.parametrize("branch", range(len(BRANCHES)))
def test_gpu_resident_assembly_every_branch_bit_exact(kind, branch):
    m, g, pt = _mesh(kind)
    if kind == "2d" and branch >= 3:
        scale = 1e-3        # make the non-linear terms visible
    else:
        scale = 1e-5 if branch < 3 else 1e-2
    _check_resident_chain(m, g, pt, branch, scale)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["hex", "tet"])
def test_gpu_resident_chain_bit_exact_at_one_million_cells(kind):
    """The same bit-for-bit comparison at a size that fills the machine (grid-stride loops with 24 CTAs per SM, every SM
    busy, 32-bit index products past 2^31 bytes): 1.02 M hexes / 1.05 M jittered tets, all branches switched on."""
    if kind == "hex":
        m = meshes.hex_box(160, 64, 100); pt = dict(PT3, free="symmetryDisplacement")
    else:
        m = meshes.tet_box(56, 56, 56, jitter=0.15); pt = dict(PT3)
    assert m.nCells > 1000000
    g = geometry.compute(m)
    _check_resident_chain(m, g, pt, len(BRANCHES) - 1, 1e-2, lsq_data=lambda L: L.export())


def _loaded_case(m, g, pt, E, nu, rho, tz, cls, **extra):
    mu, lam = cases.lame(E, nu, planeStress=extra.pop("planeStress", False))
    nBF, nIF = m.nFaces - m.nInternalFaces, m.nInternalFaces
    tr = np.zeros((nBF, 3)); p = m.patch("loaded")
    tr[p["start"] - nIF:p["start"] - nIF + p["size"], extra.pop("dir", 2)] = tz
    return cls(m, g, pt, mu, lam, rho, traction=tr, **extra)


def _compare_loops(c, gcase, r, rg, fields=("D", "pointD", "gradD", "sigma", "epsilon", "U")):
    assert rg["nOuterIterations"] == r["nOuterIterations"], (rg["nOuterIterations"], r["nOuterIterations"])
    assert np.array_equal(rg["pcgIterations"], np.array(r["pcgIterations"])), "PCG iteration counts differ"
    assert rg["enforceLinear"] == int(r["enforceLinear"])
    h, hg = r["residualHistory"], rg["residualHistory"]
    # res = max|D - D.prevIter| / max|D - D.oldTime| is a difference quotient: near convergence it carries the 1e-13-level
    # differences of the PCG solutions (re-associated global sums) divided by res itself -- hence the floor relative to the first residual
    assert np.all(np.abs(h - hg) <= 1e-6*np.abs(h) + 1e-10*np.abs(h).max())
    for k in fields:
        a, b = getattr(c, k), gcase.get(k)
        scale = np.abs(a).max()
        assert np.abs(a - b.reshape(a.shape)).max() <= 1e-9*scale + 1e-300, k


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["hex", "tet"])
def test_gpu_resident_evolve_steady_matches_oracle(kind):
    from fluidstructureinteraction_b200.solid import GpuSolidCase
    m, g, pt = _mesh(kind)
    c = _loaded_case(m, g, pt, 2e11, 0.3, 7854.0, -1e4, oev.SolidCase)
    gc = _loaded_case(m, g, pt, 2e11, 0.3, 7854.0, -1e4, GpuSolidCase)
    kw = dict(nCorrectors=60, convergenceTolerance=1e-6, relConvergenceTolerance=0.0, tolerance=1e-9, relTol=0.01)
    r, rg = c.evolve(**kw), gc.evolve(**kw)
    _compare_loops(c, gc, r, rg)
    gc.close()


@pytest.mark.gpu
def test_gpu_resident_evolve_two_time_steps_nonlinear_euler_relaxed():
    """The beamInCrossFlow solid settings (run/fsiFoam/beamInCrossFlow/solid: nonLinear yes, d2dt2(D) Euler, ddt(D) Euler,
    relaxationFactors D 0.5, relTol 0.1, convergenceTolerance 1e-7, relConvergenceTolerance 1e-3, E 1e4, nu 0.4, rho 1000)."""
    from fluidstructureinteraction_b200.solid import GpuSolidCase
    m, g, pt = _mesh("hex")
    c = _loaded_case(m, g, pt, 1e4, 0.4, 1000.0, -5.0, oev.SolidCase)
    gc = _loaded_case(m, g, pt, 1e4, 0.4, 1000.0, -5.0, GpuSolidCase)
    kw = dict(nCorrectors=1000, convergenceTolerance=1e-7, relConvergenceTolerance=1e-3, tolerance=1e-9, relTol=0.1,
              d2dt2="Euler", ddt="Euler", deltaT=0.01, nonLinear=True, relaxD=0.5, K=0.5)
    for step in range(2):
        c.new_time_step(); gc.new_time_step()
        r, rg = c.evolve(**kw), gc.evolve(**kw)
        _compare_loops(c, gc, r, rg)
    gc.close()


@pytest.mark.gpu
def test_gpu_resident_evolve_thermal():
    from fluidstructureinteraction_b200.solid import GpuSolidCase
    m, g, pt = _mesh("hexsym")
    E, nu = 2e11, 0.3
    threeK = E/(1 - 2*nu)*1.0
    c = _loaded_case(m, g, pt, E, nu, 7854.0, -1e4, oev.SolidCase, threeK=threeK, alpha=1.1e-5)
    gc = _loaded_case(m, g, pt, E, nu, 7854.0, -1e4, GpuSolidCase, threeK=threeK, alpha=1.1e-5)
    DT = 20.0 + 30.0*g.C[:, 0]; DTb = 20.0 + 30.0*g.Cf[m.nInternalFaces:, 0]
    c.DT, c.DTb = DT, DTb
    gc.set("DT", DT); gc.set("DTb", DTb)
    kw = dict(nCorrectors=40, convergenceTolerance=1e-6, tolerance=1e-9, relTol=0.01, d2dt2="Euler", ddt="Euler", deltaT=1.0,
              thermal=True, relaxD=0.5)
    c.new_time_step(); gc.new_time_step()
    r, rg = c.evolve(**kw), gc.evolve(**kw)
    _compare_loops(c, gc, r, rg)
    gc.close()


@pytest.mark.gpu
def test_gpu_resident_evolve_2d_empty_two_components():
    from fluidstructureinteraction_b200.solid import GpuSolidCase
    m, g, pt = _mesh("2d")
    c = _loaded_case(m, g, pt, 2e11, 0.3, 7854.0, -1e4, oev.SolidCase, planeStress=True, dir=1, validCmpts=(1, 1, 0))
    gc = _loaded_case(m, g, pt, 2e11, 0.3, 7854.0, -1e4, GpuSolidCase, planeStress=True, dir=1, validCmpts=(1, 1, 0))
    kw = dict(nCorrectors=50, convergenceTolerance=1e-6, tolerance=1e-9, relTol=0.01)
    r, rg = c.evolve(**kw), gc.evolve(**kw)
    _compare_loops(c, gc, r, rg, fields=("D", "pointD", "gradD", "sigma"))
    assert np.all(gc.get("D")[:, 2] == 0.0)
    gc.close()
