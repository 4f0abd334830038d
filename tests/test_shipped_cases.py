"""The REAL D equation of the reference's shipped solid cases -- shipped polyMesh, shipped material (rheologyProperties),
shipped 0/D and 0/pointD boundary types, shipped stressProperties / fvSolution / fvSchemes settings (tests/golden/
<case>.case.json = the parsed dictionaries, written by tests/golden/make_case_fixtures.py) -- assembled, solved and iterated
by the device-resident evolve() and by the oracle's restatement of the same loop: identical outer and PCG iteration counts,
fields at the north_star bars.

  beam_solid  run/fsiFoam/beamInCrossFlow/solid: nonLinear yes, d2dt2(D) Euler, ddt(D) Euler, relaxationFactors D 0.5,
              tractionDisplacement interface / fixedDisplacement bottom / symmetryPlane; pointD fixedValue + symmetryPlane.
              The interface load comes from the fluid in an FSI run; here a uniform traction stands in for it.
              snGrad(D) `corrected` [EXT]: the mesh is orthogonal (checked below), where it and skewCorrected 1 both vanish.
  tube_solid  run/fsiFoam/3dTube/solid: linear, skewCorrected 1, fixedValue inlet/outlet, two symmetry planes, pressure on
              the inner wall, d2dt2(D)/ddt(D) `backward` (numerics/backwardD2dt2Scheme/backwardD2dt2Scheme.C:230-280 over [EXT]
              backwardDdtScheme; five time steps, across its first-order start-up); also run steadyState (checked against
              Lame's thick-walled cylinder) and Euler."""
import json
import os

import numpy as np
import pytest

from oracle import evolve as oev
from fluidstructureinteraction_b200 import cases, foamdict, geometry, meshes

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def _load(name):
    m = meshes.PolyMesh.load_npz(os.path.join(GOLD, name + ".polymesh.npz"))
    with open(os.path.join(GOLD, name + ".case.json")) as f:
        d = json.load(f)
    return m, geometry.compute(m), cases.solid_case_settings(d), d


def _make(cls, m, g, S, load):
    tr, pr, fv = cases.boundary_fields(m, S, **load)
    mu, lam = cases.lame(S["E"], S["nu"], S["planeStress"])
    return cls(m, g, S["patchTypes"], mu, lam, S["rho"], traction=tr, pressure=pr, fixedValue=fv,
               pointPatchTypes=S["pointPatchTypes"])


def _evolve(c, S):
    return c.evolve(S["nCorrectors"], S["convergenceTolerance"], S["relConvergenceTolerance"], S["tolerance"], S["relTol"],
                    d2dt2=S["d2dt2"], ddt=S["ddt"], deltaT=S["deltaT"], nonLinear=S["nonLinear"], limitCoeff=S["limitCoeff"],
                    relaxD=S["relaxD"], maxIter=S["maxIter"])


BEAM_LOAD = dict(traction={"interface": [2.0, 0.0, 0.0]})
TUBE_LOAD = dict(pressure={"inner-wall": 1333.2})


# ------------------------------------------------------------------------------------------------------------------ CPU
def test_shipped_dictionaries_parse_to_the_shipped_settings():
    _, _, S, d = _load("beam_solid")
    assert (S["E"], S["nu"], S["rho"], S["planeStress"]) == (1e4, 0.4, 1000.0, False)
    assert (S["nCorrectors"], S["convergenceTolerance"], S["relConvergenceTolerance"], S["nonLinear"]) == (1000, 1e-7, 1e-3, True)
    assert (S["tolerance"], S["relTol"], S["relaxD"], S["d2dt2"], S["ddt"], S["deltaT"]) == (1e-9, 0.1, 0.5, "Euler", "Euler", 1e-3)
    assert S["patchTypes"] == {"interface": "tractionDisplacement", "bottom": "fixedDisplacement", "symmetry": "symmetryPlane"}
    assert S["pointPatchTypes"] == {"interface": "calculated", "bottom": "fixedValue", "symmetry": "symmetryPlane"}
    _, _, S, _ = _load("tube_solid")
    assert (S["E"], S["nu"], S["rho"], S["nCorrectors"], S["nonLinear"], S["d2dt2"], S["limitCoeff"]) == (3e5, 0.3, 1200.0, 200, False, "backward", 1.0)
    assert S["patchTypes"]["inlet"] == "fixedValue" and S["patchTypes"]["symmetry-x"] == "symmetryPlane"
    assert S["patchTypes"]["inner-wall"] == "tractionDisplacement"
    with open(os.path.join(GOLD, "plateHole.case.json")) as f:
        P = cases.solid_case_settings(json.load(f))
    for k in ("E", "nu", "rho", "planeStress", "nCorrectors", "convergenceTolerance", "tolerance", "relTol", "limitCoeff",
              "patchTypes", "pointPatchTypes"):
        assert P[k] == cases.PLATE_HOLE[k], k
    assert (P["d2dt2"], P["ddt"], P["nonLinear"], P["relaxD"]) == ("steadyState", "steadyState", False, None)


@pytest.mark.skipif(not os.path.exists("/root/reference/run"), reason="reference tree not present (GPU box)")
def test_case_fixtures_are_the_shipped_dictionaries():
    for name in ("beam_solid", "tube_solid", "plateHole", "flange"):
        with open(os.path.join(GOLD, name + ".case.json")) as f:
            d = json.load(f)
        for key, rel in (("D", "0/D"), ("fvSolution", "system/fvSolution"), ("stressProperties", "constant/stressProperties")):
            fresh = json.loads(json.dumps(foamdict.read(os.path.join("/root/reference", d["source"], rel))))
            assert fresh == d[key], (name, key)


def test_dictionary_reader_syntax():
    d = foamdict.parse("""
        FoamFile { version 2.0; class dictionary; }
        // comment
        a 1; b 2.5e-3; c word; /* block
        comment */ s "quoted string";
        rho rho [1 -3 0 0 0 0 0] 7854;
        sub { laplacian(DD,D) Gauss linear skewCorrected 1; interpolate(grad(U)) linear; n (2 2 1); empty (); }
        list 3 ( (0 0 0) (1 0 0) (1 1 0) );
        value uniform (0 0 -9.81);
        dup 1; dup 2;
    """)
    assert d["a"] == 1 and d["b"] == 2.5e-3 and d["c"] == "word" and d["s"] == "quoted string"
    assert foamdict.dimensioned(d["rho"]) == 7854.0 and d["rho"][1] == (1, -3, 0, 0, 0, 0, 0)
    assert d["sub"]["laplacian(DD,D)"] == ["Gauss", "linear", "skewCorrected", 1] and d["sub"]["interpolate(grad(U))"] == "linear"
    assert d["sub"]["n"] == [2, 2, 1] and d["sub"]["empty"] == []
    assert d["list"] == [3, [[0, 0, 0], [1, 0, 0], [1, 1, 0]]]
    assert foamdict.uniform(d["value"], 2, 3) == [[0.0, 0.0, -9.81]]*2 and d["dup"] == 2
    assert foamdict.switch("yes") and not foamdict.switch("off")


def test_beam_mesh_is_orthogonal_so_corrected_equals_skew_corrected():
    m, g, _, _ = _load("beam_solid")
    nIF = m.nInternalFaces
    own, nei = m.owner[:nIF], m.neighbour
    nh = g.Sf[:nIF]/g.magSf[:nIF, None]
    d = g.C[nei] - g.C[own]
    # [EXT] correctedSnGrad's nonOrthCorrectionVectors = nHat - delta*deltaCoeffs, skewCorrected's kP, kN: all ~ 1e-14
    assert np.abs(nh - d*g.deltaCoeffs[:nIF, None]).max() < 1e-12
    kP = g.Cf[:nIF] - g.C[own]
    kP = kP - nh*np.sum(kP*nh, axis=1)[:, None]
    assert np.abs(kP).max() < 1e-14*1e2


def test_oracle_tube_matches_lame_thick_walled_cylinder():
    """Pressurised tube, ends held: away from the ends the hoop stress at the inner wall follows Lame,
    p (ro^2 + ri^2)/(ro^2 - ri^2); a second known answer the restated loop has to reproduce (shipped mesh and material)."""
    m, g, S, _ = _load("tube_solid")
    S = dict(S, d2dt2="steadyState", ddt="steadyState")
    c = _make(oev.SolidCase, m, g, S, TUBE_LOAD)
    r = _evolve(c, S)
    assert r["res"] <= max(S["convergenceTolerance"], S["relConvergenceTolerance"]*r["maxRes"])
    ri, ro, p = 0.005, 0.006, TUBE_LOAD["pressure"]["inner-wall"]
    C = g.C
    rad = np.hypot(C[:, 0], C[:, 1])
    mid = np.abs(C[:, 2] - 0.025) < 0.004
    # hoop stress of a cell: t . sigma . t with t the circumferential direction
    t = np.stack([-C[:, 1]/rad, C[:, 0]/rad], axis=1)
    s = c.sigma
    hoop = t[:, 0]*t[:, 0]*s[:, 0] + 2*t[:, 0]*t[:, 1]*s[:, 1] + t[:, 1]*t[:, 1]*s[:, 3]
    lame = p*ri*ri/(ro*ro - ri*ri)*(1.0 + ro*ro/(rad*rad))
    assert np.abs(hoop[mid]/lame[mid] - 1.0).max() < 0.02
    # radial displacement of the same cells (plane strain, ends held): u_r = p ri^2 (1+nu)/(E (ro^2-ri^2)) ((1-2nu) r + ro^2/r)
    ur = (c.D[:, 0]*C[:, 0] + c.D[:, 1]*C[:, 1])/rad
    ua = p*ri*ri*(1 + S["nu"])/(S["E"]*(ro*ro - ri*ri))*((1 - 2*S["nu"])*rad + ro*ro/rad)
    assert np.abs(ur[mid]/ua[mid] - 1.0).max() < 0.02


def test_oracle_backward_d2dt2_reproduces_uniform_acceleration():
    """Known answer for the restated `backward` d2dt2 / ddt: a free body under a uniform body force accelerates uniformly,
    D = g t^2/2.  With the exact history in the four old-time levels one step of the scheme must land on the exact
    displacement and velocity (second-order backward differences are exact for quadratics) -- to rounding."""
    m = meshes.hex_box(6, 3, 4, L=0.6, W=0.3, H=0.4)
    g = geometry.compute(m)
    pt = {"fixed": "tractionDisplacement", "loaded": "tractionDisplacement", "free": "tractionDisplacement"}
    mu, lam = cases.lame(2e5, 0.3)
    acc, dT = np.array([0.3, -0.2, 1.0]), 1e-3
    exact = lambda t: 0.5*acc*t*t
    c = oev.SolidCase(m, g, pt, mu, lam, 1000.0, g=acc)
    c.old = oev.OldTimes()
    c.old.now, c.old.ti = 5, 4
    c.D = np.tile(exact(4*dT), (m.nCells, 1))
    c.old.lev = [[np.tile(exact((3 - j)*dT), (m.nCells, 1)), 3 - j] for j in range(4)]
    c.Db[:] = exact(4*dT)
    r = c.evolve(50, 1e-12, 0.0, 1e-14, 0.0, d2dt2="backward", ddt="backward", deltaT=dT)
    assert all(c.lastBackward[k] == dT for k in ("eD2", "eDdtD", "eDdtD0", "eDdtD00", "eU"))
    assert np.abs(c.D - exact(5*dT)).max() <= 1e-13*np.abs(exact(5*dT)).max()
    assert np.abs(c.U - acc*5*dT).max() <= 1e-11*np.abs(acc*5*dT).max()


def test_backward_start_up_follows_the_old_time_levels():
    """backwardD2dt2Scheme.C:58-76 / [EXT] backwardDdtScheme::deltaT0_: which uses see GREAT in the first time steps of a run
    (oracle/ASSUMPTIONS.md #13) -- the oracle's and the product's bookkeeping (timelevels.OldTimes, here over a recorder
    instead of a device mesh) must agree step by step, and from the 4th step on every use sees the real deltaT0."""
    from fluidstructureinteraction_b200 import timelevels

    class Recorder(object):
        def __init__(self): self.copies = []
        def copy_field(self, dst, src): self.copies.append((dst, src))

    D = np.zeros((4, 3))
    o, rec = oev.OldTimes(), Recorder()
    o.now = o.ti = 1
    p = timelevels.OldTimes(rec)
    seen = []
    for step in range(1, 7):
        if step > 1:
            o.advance(); p.advance()
        for outer in range(2):
            eo = oev.backward_controls(o, D, "backward", "backward", True, 2e-3)
            ep = p.assembly("backward", "backward", True, 2e-3)
            assert eo == ep, (step, outer)
            seen.append((step, outer, eo["eD2"] == oev.GREAT, eo["eDdtD00"] == oev.GREAT))
        o.ensure(1, D); eUo = oev.GREAT if o.n_old(0) < 2 else 2e-3; o.ensure(2, D)
        assert p.velocity(2e-3) == eUo == 2e-3
        D = D + 1.0
    assert [s[2] for s in seen if s[1] == 0] == [True, True, True, False, False, False]       # eD2: GREAT in steps 1-3
    assert [s[3] for s in seen] == [True] + [False]*11                                          # eDdtD00: the very first use only
    assert rec.copies[:4] == [("Dold", "D"), ("Doldold", "Dold"), ("Dooo", "Doldold"), ("Doooo", "Dooo")]
    assert rec.copies[4:8] == [("Doooo", "Dooo"), ("Dooo", "Doldold"), ("Doldold", "Dold"), ("Dold", "D")]


@pytest.mark.parametrize("d2dt2,ddt,damping", [("Euler", "backward", True), ("steadyState", "backward", True),
                                                ("steadyState", "backward", False), ("backward", "Euler", True)])
def test_old_time_bookkeeping_of_mixed_schemes(d2dt2, ddt, damping):
    """The same bookkeeping for the other combinations fvSchemes allows: which levels each scheme's calls create, and when the
    backward ddt of the damping term / of U = fvc::ddt(D) sees GREAT (vf.nOldTimes() < 2) -- oracle and product side by side."""
    from fluidstructureinteraction_b200 import timelevels

    class Recorder(object):
        def __init__(self): self.copies = []
        def copy_field(self, dst, src): self.copies.append((dst, src))

    D = np.zeros((3, 3))
    o, p = oev.OldTimes(), timelevels.OldTimes(Recorder())
    o.now = o.ti = 1
    eUs = []
    for step in range(1, 5):
        if step > 1:
            o.advance(); p.advance()
        for outer in range(2):
            assert oev.backward_controls(o, D, d2dt2, ddt, damping, 1e-3) == p.assembly(d2dt2, ddt, damping, 1e-3), (step, outer)
        if ddt == "backward":
            o.ensure(1, D); eU = oev.GREAT if o.n_old(0) < 2 else 1e-3; o.ensure(2, D)
            assert p.velocity(1e-3) == eU
            eUs.append(eU)
        assert len(o.lev) == len(p.ti) and [t for _, t in o.lev] == p.ti
        D = D + 1.0
    if ddt == "backward":
        # only a steadyState d2dt2 without damping leaves D with a single old level (residual()'s) at the first fvc::ddt(D)
        first = oev.GREAT if (d2dt2 == "steadyState" and not damping) else 1e-3
        assert eUs == [first, 1e-3, 1e-3, 1e-3]


# ------------------------------------------------------------------------------------------------------------------ GPU
def _compare(c, gc, r, rg):
    assert rg["nOuterIterations"] == r["nOuterIterations"], (rg["nOuterIterations"], r["nOuterIterations"])
    assert np.array_equal(rg["pcgIterations"], np.array(r["pcgIterations"])), "PCG iteration counts differ"
    assert rg["enforceLinear"] == int(r["enforceLinear"])
    h, hg = r["residualHistory"], rg["residualHistory"]
    # res = max|D - D.prevIter| / max|D - D.oldTime| is a difference quotient: near convergence it carries the 1e-13-level
    # differences of the PCG solutions (re-associated global sums) divided by res itself -- hence the floor relative to the first residual
    assert np.all(np.abs(h - hg) <= 1e-6*np.abs(h) + 1e-10*np.abs(h).max())
    for k in ("D", "pointD", "gradD", "sigma", "epsilon", "U"):
        a = getattr(c, k)
        b = gc.get(k).reshape(a.shape)
        assert np.abs(a - b).max() <= 1e-9*np.abs(a).max() + 1e-300, k


@pytest.mark.gpu
def test_gpu_beam_solid_shipped_settings_three_time_steps():
    from fluidstructureinteraction_b200.solid import GpuSolidCase
    m, g, S, _ = _load("beam_solid")
    c, gc = _make(oev.SolidCase, m, g, S, BEAM_LOAD), _make(GpuSolidCase, m, g, S, BEAM_LOAD)
    for step in range(3):
        c.new_time_step(); gc.new_time_step()
        r, rg = _evolve(c, S), _evolve(gc, S)
        assert r["nOuterIterations"] > 3
        _compare(c, gc, r, rg)
    gc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", ["steadyState", "Euler", "backward"])
def test_gpu_tube_solid_shipped_settings(scheme):
    """`backward` = the shipped fvSchemes of the case, unchanged."""
    from fluidstructureinteraction_b200.solid import GpuSolidCase
    m, g, S, _ = _load("tube_solid")
    if scheme == "backward":
        assert (S["d2dt2"], S["ddt"]) == ("backward", "backward")
    S = dict(S, d2dt2=scheme, ddt=scheme)
    c, gc = _make(oev.SolidCase, m, g, S, TUBE_LOAD), _make(GpuSolidCase, m, g, S, TUBE_LOAD)
    for step in range({"steadyState": 1, "Euler": 2, "backward": 5}[scheme]):
        c.new_time_step(); gc.new_time_step()
        r, rg = _evolve(c, S), _evolve(gc, S)
        _compare(c, gc, r, rg)
    gc.close()
