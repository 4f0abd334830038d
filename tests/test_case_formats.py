"""the file formats either side of the path -- decomposePar's processor meshes and *ProcAddressing lists
(src/utilities/decomposePar/decomposeMesh.C:44-843, domainDecomposition.C:95-620), `method simple` ([EXT] simpleGeomDecomp),
the polyMesh writer / reader, and the ANSYS import of the thermal case (run/thermalStressFoam/flange/flange.ans)."""
import json
import os

import numpy as np
import pytest

import oracle
from fluidstructureinteraction_b200 import ansys, blockmesh, cases, decompose, geometry, meshes

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
FLANGE = "/root/reference/run/thermalStressFoam/flange/flange.ans"


def _plate_hole():
    with open(os.path.join(GOLD, "plateHole.blockMeshDict.json")) as f:
        m = blockmesh.block_mesh(json.load(f))
    return m, geometry.compute(m)


def test_simple_decomposition_method():
    # shipped run/stressFoam/plateHole/plateHole/system/decomposeParDict: method simple; n (2 2 1); delta 0.001
    m, g = _plate_hole()
    cp = decompose.simple_decomp(g.C, (2, 2, 1), 0.001)
    assert np.bincount(cp).tolist() == [252, 248, 248, 252] or sorted(np.bincount(cp).tolist()) == [248, 248, 252, 252]
    assert np.bincount(cp).sum() == 1000 and cp.max() == 3
    # groups along x are halves by sorted (slightly rotated) x: every cell of group 0 lies left of every cell of group 1
    gx = cp % 2
    assert g.C[gx == 0, 0].max() <= g.C[gx == 1, 0].min() + 2e-3*2.0
    assert decompose._assign_groups(10, 4).tolist() == [0, 0, 0, 1, 1, 1, 2, 2, 3, 3]
    # on a box whose edges divide evenly it is the block decomposition of the benchmark
    mb = meshes.hex_box(8, 4, 6); gb = geometry.compute(mb)
    assert np.array_equal(decompose.simple_decomp(gb.C, (2, 2, 2), 0.001), decompose.simple_cell_proc(8, 4, 6, 2, 2, 2))


def test_processor_meshes_and_addressing_round_trip(tmp_path):
    m, g = _plate_hole()
    cp = decompose.simple_decomp(g.C, (2, 2, 1), 0.001)
    doms = decompose.write_case(str(tmp_path), m, cp, 4)
    back = decompose.read_case(str(tmp_path), 4)
    assert sorted(os.listdir(tmp_path / "processor2" / "constant" / "polyMesh")) == sorted(
        ["points", "faces", "owner", "neighbour", "boundary", "cellProcAddressing", "faceProcAddressing",
         "pointProcAddressing", "boundaryProcAddressing"])
    cellSeen = np.zeros(m.nCells, dtype=int)
    faceSeen = np.zeros(m.nFaces, dtype=int)
    for p, (a, b) in enumerate(zip(doms, back)):
        ma, mb = a["mesh"], b["mesh"]
        assert np.array_equal(ma.points, mb.points) and np.array_equal(ma.facePoints, mb.facePoints)
        assert np.array_equal(ma.faceStart, mb.faceStart) and np.array_equal(ma.owner, mb.owner) and np.array_equal(ma.neighbour, mb.neighbour)
        assert ma.patches == mb.patches
        for k in ("cellProcAddressing", "faceProcAddressing", "pointProcAddressing", "boundaryProcAddressing"):
            assert np.array_equal(a[k], b[k]), k
        # a valid fvMesh on its own: upper-triangular internal faces, closed cells, cells / points / faces ascending global id
        assert ma.is_upper_triangular()
        gg = geometry.compute(ma)
        s = np.zeros((ma.nCells, 3)); np.add.at(s, ma.owner, gg.Sf); np.subtract.at(s, ma.neighbour, gg.Sf[:ma.nInternalFaces])
        assert np.abs(s).max() < 1e-15 and gg.V.min() > 0
        assert np.all(np.diff(a["cellProcAddressing"]) > 0) and np.all(np.diff(a["pointProcAddressing"]) > 0)
        assert np.array_equal(ma.points, m.points[a["pointProcAddressing"]])
        assert np.allclose(gg.V, g.V[a["cellProcAddressing"]], rtol=1e-13)
        # original patches first (all of them, empty ones kept), processor patches last with -1
        nOrig = len(m.patches)
        assert a["boundaryProcAddressing"][:nOrig].tolist() == list(range(nOrig)) and np.all(a["boundaryProcAddressing"][nOrig:] == -1)
        for pt in ma.patches[nOrig:]:
            assert pt["type"] == "processor" and pt["myProcNo"] == p and pt["name"] == "procBoundary%dto%d" % (p, pt["neighbProcNo"])
        cellSeen[a["cellProcAddressing"]] += 1
        faceSeen[np.abs(a["faceProcAddressing"]) - 1] += 1
        # the turning index: a turned face is the global face reversed, and this processor holds its neighbour cell
        fa = a["faceProcAddressing"]
        for i in np.flatnonzero(fa < 0)[:50]:
            gf = -fa[i] - 1
            loop = m.facePoints[m.faceStart[gf]:m.faceStart[gf + 1]]
            mine = a["pointProcAddressing"][ma.facePoints[ma.faceStart[i]:ma.faceStart[i + 1]]]
            assert mine.tolist() == [loop[0]] + loop[:0:-1].tolist()
            assert a["cellProcAddressing"][ma.owner[i]] == m.neighbour[gf]
    assert np.all(cellSeen == 1)
    nIF = m.nInternalFaces
    cut = cp[m.owner[:nIF]] != cp[m.neighbour]
    assert np.all(faceSeen[:nIF] == np.where(cut, 2, 1)) and np.all(faceSeen[nIF:] == 1)
    # the two sides of a processor boundary list the same global faces in the same order
    for p, a in enumerate(doms):
        for pt in a["mesh"].patches:
            if pt["type"] != "processor":
                continue
            q = pt["neighbProcNo"]
            other = [o for o in doms[q]["mesh"].patches if o["type"] == "processor" and o["neighbProcNo"] == p][0]
            fa = np.abs(a["faceProcAddressing"][pt["start"]:pt["start"] + pt["size"]])
            fb = np.abs(doms[q]["faceProcAddressing"][other["start"]:other["start"] + other["size"]])
            assert np.array_equal(fa, fb) and np.all(np.diff(fa) > 0)


def test_decomposed_scalar_system_reaches_the_serial_solution():
    m, g = _plate_hole()
    sysm = cases.scalar_laplacian_system(m, g, 50.0, fixedValue={"left": 273.0, "right": 573.0}, rate=7800.0*500.0/1.0)
    diag = sysm["diag"].copy(); b = sysm["source"].copy()
    for fc, ic, bc in sysm["patches"]:
        np.add.at(diag, fc, ic); np.add.at(b, fc, bc)
    x = np.zeros(m.nCells)
    oracle.pcg_solve(sysm["l"], sysm["u"], diag, sysm["upper"], x, b, 1e-12, 0.0)
    cp = decompose.simple_decomp(g.C, (2, 2, 1), 0.001)
    doms = decompose.decompose_ldu(sysm["l"], sysm["u"], diag, sysm["upper"], b, np.zeros(m.nCells), cp, 4)
    perf, hist = oracle.pcg_solve_multi(doms, 1e-12, 0.0)
    xr = decompose.reconstruct(doms, m.nCells)
    assert perf["converged"] and np.abs(xr - x).max() <= 1e-8*np.abs(x).max()


ANS = """/PREP7
N,1,0,0,0
N,2,1,0,0
N,3,1,1,0
N,4,0,1
N,5,0,0,1
N,6,1,0,1
N,7,1,1,1
N,8,0,1,1
N,9,2,0,0
N,10,2,0,1
N,12,0.5,0.5,2
N,20,1.5,0.3,2
EN,1,1,2,3,4,5,6,7,8
EN,2,2,9,3,3,6,10,7,7
EN,3,5,6,7,8,12,12,12,12
EN,4,6,10,7,7,20,20,20,20
"""


def test_ans_reader_shapes(tmp_path):
    p = tmp_path / "t.ans"
    p.write_text(ANS)
    pts, shapes = ansys.read_ans(str(p), 2.0)
    assert [s[0] for s in shapes] == ["hex", "prism", "pyr", "tet"] and pts.shape == (12, 3) and pts[8].tolist() == [4.0, 0.0, 0.0]
    m = ansys.ans_to_polymesh(str(p))
    assert m.nCells == 4 and m.nInternalFaces == 3 and m.is_upper_triangular()
    assert [(q["name"], q["size"]) for q in m.patches] == [("defaultFaces", 6 + 5 + 5 + 4 - 2*3)]
    g = geometry.compute(m)
    s = np.zeros((4, 3)); np.add.at(s, m.owner, g.Sf); np.subtract.at(s, m.neighbour, g.Sf[:3])
    assert np.abs(s).max() < 1e-15 and g.V.min() > 0
    assert abs(g.V[0] - 1.0) < 1e-14 and abs(g.V[1] - 0.5) < 1e-14 and abs(g.V[2] - 1.0/3.0) < 1e-14


@pytest.mark.skipif(not os.path.exists(FLANGE), reason="reference tree not present (GPU box)")
def test_flange_import_matches_the_shipped_boundary_file():
    """constant/polyMesh/boundary of the shipped case: startFace 15316 (= internal faces), patches of 2440 + 348 + 96 + 384 =
    3268 boundary faces; flange.ans: 7189 nodes, 5712 elements -- reference-held numbers for the importer and the mesher."""
    m = ansys.ans_to_polymesh(FLANGE, 0.001)
    assert (m.nCells, m.nPoints, m.nInternalFaces, m.nFaces - m.nInternalFaces) == (5712, 7189, 15316, 2440 + 348 + 96 + 384)
    assert m.is_upper_triangular()
    g = geometry.compute(m)
    s = np.zeros((m.nCells, 3)); np.add.at(s, m.owner, g.Sf); np.subtract.at(s, m.neighbour, g.Sf[:m.nInternalFaces])
    assert np.abs(s).max() < 1e-18 and g.V.min() > 0
    lab, sizes = ansys.boundary_surfaces(m, g, 80.0)
    assert 348 in sizes.tolist() and 384 in sizes.tolist()          # patch2 (flat face, 273 K) and patch4 (bore, 573 K)
    flat = np.flatnonzero(sizes == 348)[0]
    nIF = m.nInternalFaces
    assert np.allclose(g.Cf[nIF:][lab == flat, 1], -0.0075, atol=1e-9)


def test_repatch_regroups_boundary_faces_and_keeps_the_mesh():
    """PolyMesh.repatch ([EXT] autoPatch / createPatch-like, used to give the imported flange its T patches): internal faces and
    cells untouched, every boundary face kept once with its own points and owner, faces of a new patch in their old order."""
    from fluidstructureinteraction_b200 import geometry
    m = meshes.hex_box(4, 3, 2)
    nIF, nBF = m.nInternalFaces, m.nFaces - m.nInternalFaces
    g = geometry.compute(m)
    lab = (g.Cf[nIF:, 2] > g.Cf[nIF:, 2].mean()).astype(int) + 2*(g.Cf[nIF:, 0] < 1e-9)
    names = ["low", "high", "lowx0", "highx0"]
    m2, order = m.repatch(lab, names)
    assert m2.nCells == m.nCells and m2.nInternalFaces == nIF and np.array_equal(m2.neighbour, m.neighbour)
    assert np.array_equal(m2.owner[:nIF], m.owner[:nIF]) and sorted(order.tolist()) == list(range(nBF))
    assert [p["size"] for p in m2.patches] == np.bincount(lab, minlength=4).tolist() and m2.patches[0]["start"] == nIF
    for k, bf in enumerate(order):
        f2, f = nIF + k, nIF + bf
        assert m2.owner[f2] == m.owner[f]
        assert np.array_equal(m2.facePoints[m2.faceStart[f2]:m2.faceStart[f2 + 1]], m.facePoints[m.faceStart[f]:m.faceStart[f + 1]])
    for p in m2.patches:                         # stable inside a patch
        o = order[p["start"] - nIF:p["start"] - nIF + p["size"]]
        assert np.all(np.diff(o) > 0)
    g2 = geometry.compute(m2)
    assert np.array_equal(g2.V, g.V) and np.array_equal(g2.Cf[nIF:], g.Cf[nIF:][order])
