"""BASELINE.json configs[0]: the shipped run/stressFoam/plateHole case, end to end, against the only known answer the
reference itself holds -- the Kirsch solution coded in its function object (run/stressFoam/plateHole/setPlateHoleBC/
setPlateHoleBC.C:57-152) and the error norms its calcError() prints (:250-388).

Mesh: the shipped blockMeshDict (five blocks, arc edges; tests/golden/plateHole.blockMeshDict.json is its parsed content,
written by tests/golden/make_fixtures.py) through the blockMesh restatement.  Case: 0/D, 0/pointD, rheologyProperties
(planeStress), stressProperties (nCorrectors 1000, convergenceTolerance 1e-8), fvSolution (PCG/DIC 1e-9 / 0.01), fvSchemes
(steadyState, skewCorrected 1) -- cases.PLATE_HOLE.  setBC() (:202-247) overwrites the tractions with n & sigma_Kirsch.

The analytical solution is independent of every restatement in oracle/: assembly, boundary conditions, least-squares
vol->point interpolation (+ the symmetryPlane point patches), vertex-based gradients, segregated PCG/DIC solves and the
outer loop have to be right together for D to come out at discretisation error and to converge at second order."""
import json
import os

import numpy as np
import pytest

from oracle import evolve as oev
import oracle
from fluidstructureinteraction_b200 import blockmesh, cases, geometry, lsq

HERE = os.path.dirname(os.path.abspath(__file__))
DICT = os.path.join(HERE, "golden", "plateHole.blockMeshDict.json")
REF_DICT = "/root/reference/run/stressFoam/plateHole/plateHole/constant/polyMesh/blockMeshDict"
PH = cases.PLATE_HOLE


def _dict():
    with open(DICT) as f:
        return json.load(f)


def _case(refine, cls, **kw):
    m = blockmesh.block_mesh(_dict(), refine)
    g = geometry.compute(m)
    mu, lam = cases.lame(PH["E"], PH["nu"], PH["planeStress"])
    tr = cases.plate_hole_tractions(m, g, PH["patchTypes"])
    c = cls(m, g, PH["patchTypes"], mu, lam, PH["rho"], traction=tr, validCmpts=(1, 1, 0), pointPatchTypes=PH["pointPatchTypes"], **kw)
    return m, g, c


def _run(c):
    return c.evolve(PH["nCorrectors"], PH["convergenceTolerance"], 0.0, PH["tolerance"], PH["relTol"], limitCoeff=PH["limitCoeff"])


# ------------------------------------------------------------------------------------------------------------------ CPU
@pytest.mark.skipif(not os.path.exists(REF_DICT), reason="reference tree not present (GPU box)")
def test_fixture_is_the_shipped_block_mesh_dict():
    assert blockmesh.read_block_mesh_dict(REF_DICT) == _dict()


def test_block_mesh_reproduces_the_shipped_case_sizes():
    """C1 plateHole N = 1000, F = 1930; patches of the shipped blockMeshDict:62-100 in dictionary order."""
    m = blockmesh.block_mesh(_dict())
    assert (m.nCells, m.nInternalFaces, m.nFaces, m.nPoints) == (1000, 1930, 4070, 2142)
    assert m.is_upper_triangular()
    assert [(p["name"], p["type"], p["size"]) for p in m.patches] == [
        ("left", "symmetryPlane", 30), ("right", "patch", 30), ("down", "symmetryPlane", 30), ("up", "patch", 30),
        ("hole", "patch", 20), ("frontAndBack", "empty", 2000)]
    g = geometry.compute(m)
    # closed cells, positive volumes, total volume = plate minus the (polygonal) quarter hole, 0.5 thick
    s = np.zeros((m.nCells, 3))
    np.add.at(s, m.owner, g.Sf); np.subtract.at(s, m.neighbour, g.Sf[:m.nInternalFaces])
    assert np.abs(s).max() < 1e-15 and g.V.min() > 0
    exact = (4.0 - np.pi*0.25/4.0)*0.5
    assert exact < g.V.sum() < exact*(1 + 1e-4)
    # arcs: the hole's points lie on r = 0.5, the block interface 1-4-9 on r = 1
    p = m.patch("hole")
    fp = m.facePoints.reshape(-1, 4)[p["start"]:p["start"] + p["size"]].ravel()
    r = np.hypot(m.points[fp, 0], m.points[fp, 1])
    assert np.abs(r - 0.5).max() < 2e-6            # the dictionary's arc points carry six digits
    # cells are numbered block by block, i fastest: cell 0 sits at the hole next to the y axis side of block 0
    assert np.hypot(*g.C[0, :2]) < 0.56 and g.C[0, 0] < g.C[0, 1]


def test_block_mesh_grading_and_straight_blocks():
    d = dict(scale=2.0, vertices=[[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]],
             blocks=[dict(v=list(range(8)), n=[4, 3, 2], grading=[2.0, 1.0, 0.5])], edges=[],
             patches=[dict(type="patch", name="all", faces=[[0, 4, 7, 3], [1, 2, 6, 5], [0, 1, 5, 4], [3, 7, 6, 2], [0, 3, 2, 1], [4, 5, 6, 7]])])
    m = blockmesh.block_mesh(d)
    assert (m.nCells, m.nPoints, m.nInternalFaces) == (24, 60, 3*3*2 + 4*2*2 + 4*3*1)
    x = np.unique(np.round(m.points[:, 0], 12))
    dx = np.diff(x)
    assert abs(dx[-1]/dx[0] - 2.0) < 1e-10 and abs(x[-1] - 2.0) < 1e-15
    z = np.unique(np.round(m.points[:, 2], 12))
    assert abs(np.diff(z)[-1]/np.diff(z)[0] - 0.5) < 1e-10
    g = geometry.compute(m)
    assert abs(g.V.sum() - 8.0) < 1e-13


def test_point_constraint_builders_agree_and_project():
    m = blockmesh.block_mesh(_dict())
    g = geometry.compute(m)
    a = lsq.point_constraints(m, g, PH["pointPatchTypes"])
    b = oracle.point_constraints_build(m, g, PH["pointPatchTypes"])
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert a["point"].shape[0] == 2*2*31               # two symmetry planes, 31 x 2 points each
    pD = np.random.default_rng(0).standard_normal((m.nPoints, 3))
    q = oracle.point_constraints_apply(a, pD.copy())
    left = np.abs(m.points[:, 0]) < 1e-12
    down = np.abs(m.points[:, 1]) < 1e-12
    assert np.abs(q[left, 0]).max() < 1e-15 and np.abs(q[down, 1]).max() < 1e-15
    free = ~(left | down)
    assert np.array_equal(q[free], pD[free])
    # a fixedValue point patch wins over what was interpolated; a later symmetryPlane patch then projects it
    c = lsq.point_constraints(m, g, dict(PH["pointPatchTypes"], right="fixedValue"), {"right": [1.0, 2.0, 3.0]})
    q = oracle.point_constraints_apply(c, pD.copy())
    right = np.abs(m.points[:, 0] - 2.0) < 1e-12
    assert np.array_equal(q[right & ~down], np.broadcast_to([1.0, 2.0, 3.0], (int((right & ~down).sum()), 3)))
    assert np.allclose(q[right & down], [1.0, 0.0, 3.0], atol=1e-15)


# bars on the reference's own error norms (calcError), as fractions of the analytical maxima on the same mesh
KIRSCH_BARS = {1: dict(D=0.004, sigma=0.03), 2: dict(D=0.001, sigma=0.008)}


def _check_kirsch(e, refine):
    assert e["DError"] < KIRSCH_BARS[refine]["D"]*e["DMax"], e
    assert e["pointDError"] < 1.5*KIRSCH_BARS[refine]["D"]*e["DMax"], e
    for k in ("sigmaXXErr", "sigmaXYErr", "sigmaYYErr"):
        assert e[k] < KIRSCH_BARS[refine]["sigma"]*e["sigmaXXMax"], (k, e)


def test_oracle_reproduces_kirsch_at_second_order():
    errs = {}
    for refine in (1, 2):
        m, g, c = _case(refine, oev.SolidCase)
        r = _run(c)
        assert r["res"] <= PH["convergenceTolerance"] and 0 < r["nOuterIterations"] < PH["nCorrectors"] - 1
        assert np.all(c.D[:, 2] == 0.0)
        errs[refine] = cases.plate_hole_errors(m, g, c.D, c.pointD, c.sigma)
        _check_kirsch(errs[refine], refine)
    # halving h: displacement error / 4 (second order), stress error at least / 3
    assert errs[1]["DError"]/errs[2]["DError"] > 3.5
    assert errs[1]["sigmaXXErr"]/errs[2]["sigmaXXErr"] > 3.0
    # stress concentration: 3 T at the top of the hole in the continuum, the cell next to it sees most of it
    assert 2.6*PH["T"] < errs[1]["sigmaXXMax"] < 3.0*PH["T"]


def test_kirsch_check_has_teeth():
    """The same loop with one ingredient wrong must fail the bars: plane strain instead of plane stress (the displacement
    solution of :116-152 is the plane-stress one), and the symmetryPlane point patches left out (0/pointD ignored)."""
    m = blockmesh.block_mesh(_dict()); g = geometry.compute(m)
    tr = cases.plate_hole_tractions(m, g, PH["patchTypes"])
    mu, lam = cases.lame(PH["E"], PH["nu"], False)
    c = oev.SolidCase(m, g, PH["patchTypes"], mu, lam, PH["rho"], traction=tr, validCmpts=(1, 1, 0), pointPatchTypes=PH["pointPatchTypes"])
    _run(c)
    e = cases.plate_hole_errors(m, g, c.D, c.pointD, c.sigma)
    assert e["DError"] > 10*KIRSCH_BARS[1]["D"]*e["DMax"]
    mu, lam = cases.lame(PH["E"], PH["nu"], True)
    c = oev.SolidCase(m, g, PH["patchTypes"], mu, lam, PH["rho"], traction=tr, validCmpts=(1, 1, 0))
    _run(c)
    e = cases.plate_hole_errors(m, g, c.D, c.pointD, c.sigma)
    assert e["DError"] > KIRSCH_BARS[1]["D"]*e["DMax"]


# ------------------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("refine", [1, 2])
def test_gpu_plate_hole_matches_oracle_and_kirsch(refine):
    """Config 1 end to end on the device-resident path: identical outer and inner iteration counts with the oracle,
    fields at the north_star bars, and the reference's own error norms against the analytical solution."""
    from fluidstructureinteraction_b200.solid import GpuSolidCase
    m, g, c = _case(refine, oev.SolidCase)
    _, _, gc = _case(refine, GpuSolidCase)
    r, rg = _run(c), _run(gc)
    assert rg["nOuterIterations"] == r["nOuterIterations"]
    assert np.array_equal(rg["pcgIterations"], np.array(r["pcgIterations"]))
    h, hg = r["residualHistory"], rg["residualHistory"]
    # res = max|D - D.prevIter| / max|D - D.oldTime| is a difference quotient: near convergence it carries the 1e-13-level
    # differences of the PCG solutions (re-associated global sums) divided by res itself -- hence the floor relative to the first residual
    assert np.all(np.abs(h - hg) <= 1e-6*np.abs(h) + 1e-10*np.abs(h).max())
    for k in ("D", "pointD", "gradD", "sigma"):
        a, b = getattr(c, k), gc.get(k).reshape(getattr(c, k).shape)
        assert np.abs(a - b).max() <= 1e-9*np.abs(a).max(), k
    D, pD, sg = gc.get("D"), gc.get("pointD"), gc.get("sigma")
    assert np.all(D[:, 2] == 0.0)
    _check_kirsch(cases.plate_hole_errors(m, g, D, pD, sg), refine)
    gc.close()


@pytest.mark.gpu
def test_gpu_point_constraints_bit_exact():
    import fluidstructureinteraction_b200 as pkg
    m = blockmesh.block_mesh(_dict()); g = geometry.compute(m)
    pc = lsq.point_constraints(m, g, dict(PH["pointPatchTypes"], right="fixedValue"), {"right": [1e-7, -2e-7, 0.0]})
    rng = np.random.default_rng(5)
    D = 1e-7*rng.standard_normal((m.nCells, 3)); Db = 1e-7*rng.standard_normal((m.nFaces - m.nInternalFaces, 3))
    gm = pkg.GpuMesh(m, g, PH["patchTypes"], device=0)
    gm.set_field("D", D); gm.set_field("Db", Db)
    gm.set_lsq(lsq.build(m, g, PH["patchTypes"]))
    gm.set_point_constraints(pc)
    gm.vol_to_point()
    ref = oracle.point_constraints_apply(pc, oracle.Lsq(m, g, PH["patchTypes"]).interpolate(D, Db))
    assert np.array_equal(gm.get_field("pointD"), ref)
    gm.close()
