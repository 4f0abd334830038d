"""CPU tests of the product's host side: the C ABI loads and exports every declared symbol, fails loudly
without a GPU, and the wavefront / SELL-32 schedule reproduces the reference face order (checked by
emulating the device kernels' row walks in numpy -- test code, not a product CPU path)."""
import os
import re

import numpy as np
import pytest

import oracle
import fluidstructureinteraction_b200 as pkg
from fluidstructureinteraction_b200 import capi, meshes
from conftest import FIXTURE_MESHES, ROOT, load_fixture_mesh, synthetic_matrix


def test_library_exports_every_declared_symbol():
    lib = pkg.load_library()
    with open(os.path.join(ROOT, "include", "gpupcg.h")) as f:
        header = f.read()
    declared = set(re.findall(r"\b(gpupcg_[A-Za-z_0-9]+)\s*\(", header))
    assert declared, "no prototypes found in include/gpupcg.h"
    assert declared == set(capi.SYMBOLS), (declared ^ set(capi.SYMBOLS))
    for name in declared:
        assert hasattr(lib, name), name


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run (it must never silently compute on the CPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    l, u = meshes.hex_ldu(3, 3, 3)
    with pytest.raises(pkg.GpuPcgError) as e:
        pkg.GpuPCG(l, u, 27)
    assert e.value.code == -4 and "no CPU fallback" in str(e.value)


def test_plan_rejects_non_upper_triangular():
    l = np.array([0, 1, 0], dtype=np.int32)     # lowerAddr not ascending
    u = np.array([1, 2, 2], dtype=np.int32)
    with pytest.raises(pkg.GpuPcgError) as e:
        pkg.plan_export(l, u, 3)
    assert e.value.code == -2
    with pytest.raises(pkg.GpuPcgError) as e:
        pkg.plan_export(np.array([1], dtype=np.int32), np.array([0], dtype=np.int32), 2)    # l > u
    assert e.value.code == -2


def emulate(plan, diag, upper, l, u):
    """numpy emulation of the row walks of k_amul / k_dic_fwd / k_dic_bwd on the exported schedule: tile by
    tile, round by round, in-tile values from the tile-local array, cross-tile values through the external
    lists -- exactly the indexing the kernels use."""
    rowOld, slices = plan["rowOld"], plan["slices"]
    Ucol, UfaceOld, Lidx, Lcol = plan["Ucol"], plan["UfaceOld"], plan["Lidx"], plan["Lcol"]
    tileStart, tileRoundPtr, roundStart = plan["tileStart"], plan["tileRoundPtr"], plan["roundStart"]
    LcolTile, UcolTile = plan["LcolTile"], plan["UcolTile"]
    extFPtr, extFCol, extBPtr, extBCol = plan["extFPtr"], plan["extFCol"], plan["extBPtr"], plan["extBCol"]
    nT = plan["info"]["nTiles"]
    nR = rowOld.shape[0]
    real = rowOld >= 0
    d = np.where(real, diag[np.maximum(rowOld, 0)], 1.0)
    Uval = np.where(UfaceOld >= 0, upper[np.maximum(UfaceOld, 0)], 0.0)

    def rows_in(v):
        return np.where(real, v[np.maximum(rowOld, 0)], 0.0)

    def rows_out(v):
        out = np.empty(int(real.sum()))
        out[rowOld[real]] = v[real]
        return out

    def amul(x):
        xp = rows_in(x)
        Ax = np.zeros(nR)
        for r in range(nR):
            s, lane = divmod(r, 32)
            offU, wU, offL, wL = slices[s]
            acc = d[r]*xp[r]
            for j in range(wL):
                e = offL + j*32 + lane
                if Lcol[e] >= 0:
                    acc = acc + Uval[Lidx[e]]*xp[Lcol[e]]
            for j in range(wU):
                e = offU + j*32 + lane
                if Ucol[e] >= 0:
                    acc = acc + Uval[e]*xp[Ucol[e]]
            Ax[r] = acc
        return rows_out(Ax)

    def forward(init, coef_of, combine):
        """generic forward tile sweep: out[r] = init[r] (-) combine(coef, dep)"""
        out = np.full(nR, np.nan)
        for t in range(nT):
            r0, r1 = tileStart[t], tileStart[t + 1]
            rs = roundStart[tileRoundPtr[t]:tileRoundPtr[t + 1]]
            ext = extFCol[extFPtr[t]:extFPtr[t + 1]]
            assert np.all(ext < r0)                              # cross-tile inputs come from earlier tiles
            vs = init[r0:r1].copy()
            for k in range(len(rs) - 1):
                new = {}
                for q in range(rs[k], rs[k + 1]):
                    r = r0 + q
                    s, lane = divmod(r, 32)
                    offU, wU, offL, wL = slices[s]
                    acc = vs[q]
                    for j in range(wL):
                        e = offL + j*32 + lane
                        c = LcolTile[e]
                        if c == -1:
                            break
                        if c >= 0:
                            assert not any(rs[k] <= c < rs[k + 1] for _ in [0]), "dependency inside a round"
                            v = vs[c]
                        else:
                            v = out[ext[-2 - c]]
                            assert not np.isnan(v)
                        acc = combine(acc, coef_of(r, Uval[Lidx[e]]), v)
                    new[q] = acc
                for q, a in new.items():
                    vs[q] = a
                    out[r0 + q] = a
        return out

    def factor():
        piv = forward(d.copy(), lambda r, c: c*c, lambda acc, cf, v: acc - cf/v)
        return 1.0/piv

    def precondition(rD, rA):
        rp = rows_in(rA)
        y = forward(rD*rp, lambda r, c: rD[r]*c, lambda acc, cf, v: acc - cf*v)
        z = np.full(nR, np.nan)
        for t in range(nT - 1, -1, -1):
            r0, r1 = tileStart[t], tileStart[t + 1]
            rs = roundStart[tileRoundPtr[t]:tileRoundPtr[t + 1]]
            ext = extBCol[extBPtr[t]:extBPtr[t + 1]]
            assert np.all(ext >= r1)
            vs = y[r0:r1].copy()
            for k in range(len(rs) - 2, -1, -1):
                new = {}
                for q in range(rs[k], rs[k + 1]):
                    r = r0 + q
                    s, lane = divmod(r, 32)
                    offU, wU, offL, wL = slices[s]
                    acc = vs[q]
                    for j in range(wU - 1, -1, -1):
                        e = offU + j*32 + lane
                        c = UcolTile[e]
                        if c == -1:
                            continue
                        v = vs[c] if c >= 0 else z[ext[-2 - c]]
                        assert not np.isnan(v)
                        acc = acc - (rD[r]*Uval[e])*v
                    new[q] = acc
                for q, a in new.items():
                    vs[q] = a
                    z[r0 + q] = a
        return rows_out(z)

    return amul, factor, precondition, rows_out


def check_schedule(l, u, n, seed=3, tileRows=64, tileMode=1):
    plan = pkg.plan_export(l, u, n, tileRows, tileMode)
    info = plan["info"]
    rowOld = plan["rowOld"]
    # permutation is a bijection onto the real rows; tiles are whole slices
    real = rowOld[rowOld >= 0]
    assert np.array_equal(np.sort(real), np.arange(n))
    assert info["nRows"] % 32 == 0 and info["nSlices"]*32 == info["nRows"]
    assert np.all(plan["tileStart"] % 32 == 0) and plan["tileStart"][-1] == info["nRows"]
    pos = np.empty(n, dtype=np.int64)
    pos[rowOld[rowOld >= 0]] = np.flatnonzero(rowOld >= 0)
    # acyclic tile graph, and inside a tile every face goes from an earlier round to a later one
    tile = np.searchsorted(plan["tileStart"], pos, side="right") - 1
    assert np.all(tile[l] <= tile[u])
    rnd = np.empty(n, dtype=np.int64)
    for t in range(info["nTiles"]):
        rs = plan["roundStart"][plan["tileRoundPtr"][t]:plan["tileRoundPtr"][t + 1]]
        sel = tile == t
        rnd[sel] = np.searchsorted(rs, pos[sel] - plan["tileStart"][t], side="right") - 1
    same = tile[l] == tile[u]
    assert np.all(rnd[l][same] < rnd[u][same])
    # every face appears exactly once on each side
    assert np.array_equal(np.sort(plan["UfaceOld"][plan["UfaceOld"] >= 0]), np.arange(l.shape[0]))
    assert int((plan["Lcol"] >= 0).sum()) == l.shape[0]
    diag, upper, b, x0 = synthetic_matrix(l, u, n, seed)
    amul, factor, precondition, rows_out = emulate(plan, diag, upper, l, u)
    xt = np.sin(0.37*np.arange(n))
    assert np.array_equal(amul(xt), oracle.amul(l, u, diag, upper, xt))
    rDp = factor()
    rD = oracle.dic_factor(l, u, diag, upper)
    assert np.array_equal(rows_out(rDp), rD)
    assert np.array_equal(precondition(rDp, b), oracle.precondition(l, u, upper, rD, b))
    return info


@pytest.mark.parametrize("dims", [(1, 1, 1), (7, 1, 1), (5, 4, 3), (9, 8, 7)])
def test_schedule_reproduces_reference_order_hex(dims):
    nx, ny, nz = dims
    l, u = meshes.hex_ldu(nx, ny, nz)
    for tileRows, tileMode in ((32, 0), (32, 1), (64, 1), (64, 2), (32, 2), (1024, 2)):
        info = check_schedule(l, u, nx*ny*nz, tileRows=tileRows, tileMode=tileMode)
        assert info["nLevels"] == nx + ny + nz - 2        # hyperplanes i+j+k = const


def test_schedule_reproduces_reference_order_shipped_beam():
    m = load_fixture_mesh("beam_solid")
    l, u = m.lower_upper()
    info = check_schedule(l, u, m.nCells)
    assert info["nLevels"] == 18


def test_schedule_tet_mesh():
    m = meshes.tet_box(3, 2, 2, jitter=0.15)
    assert m.is_upper_triangular()
    l, u = m.lower_upper()
    check_schedule(l, u, m.nCells)


def test_level_counts_of_shipped_meshes():
    """Level structure quoted in SURVEY.md section 7.3 (tube solid: 102 levels, beam fluid: 88)."""
    expect = {"beam_solid": 18, "tube_solid": 102, "beam_fluid": 88}
    for name, nl in expect.items():
        m = load_fixture_mesh(name)
        l, u = m.lower_upper()
        assert pkg.plan_export(l, u, m.nCells)["info"]["nLevels"] == nl


def test_mesh_generators_closed_cells():
    """Sum of outward face-area vectors of every cell vanishes (hex and jittered tet box)."""
    for m in (meshes.hex_box(4, 3, 2), meshes.tet_box(3, 2, 2, jitter=0.15)):
        nF = m.nFaces
        Sf = np.zeros((nF, 3))
        for f in range(nF):
            p = m.points[m.facePoints[m.faceStart[f]:m.faceStart[f + 1]]]
            c = p.mean(axis=0)
            for k in range(p.shape[0]):
                Sf[f] += 0.5*np.cross(p[k] - c, p[(k + 1) % p.shape[0]] - c)
        acc = np.zeros((m.nCells, 3))
        np.add.at(acc, m.owner, Sf)
        np.subtract.at(acc, m.neighbour, Sf[:m.nInternalFaces])
        assert np.max(np.abs(acc)) < 1e-12
        assert m.is_upper_triangular()
        # volumes positive: divergence theorem with x
        vol = np.zeros(m.nCells)
        for f in range(nF):
            p = m.points[m.facePoints[m.faceStart[f]:m.faceStart[f + 1]]]
            fc = p.mean(axis=0)
            vol[m.owner[f]] += np.dot(Sf[f], fc)/3
            if f < m.nInternalFaces:
                vol[m.neighbour[f]] -= np.dot(Sf[f], fc)/3
        assert np.all(vol > 0)
        assert abs(vol.sum() - 1.0) < 1e-9           # 2 x 0.5 x 1 box


def test_pencil_plan_puts_every_dependency_one_step_back():
    """Tile mode 3 (structured block): rows are numbered base + 32*step + lane per pencil so that, inside a pencil, the
    three lower neighbours of a cell sit exactly one step back -- own lane (long axis), lane - 1, lane - w1 -- which is
    what lets the sweep warp of kernels_pencil.cuh run on registers and shuffles."""
    nx, ny, nz = 40, 9, 13
    l, u = meshes.hex_ldu(nx, ny, nz)
    n = nx*ny*nz
    from fluidstructureinteraction_b200 import plan_export
    P = plan_export(l, u, n, tileRows=256, tileMode=3)
    rowOld, ts = P["rowOld"], P["tileStart"]
    pos = np.full(n, -1, dtype=np.int64)
    real = np.flatnonzero(rowOld >= 0)
    pos[rowOld[real]] = real
    assert np.all(pos >= 0) and real.size == n                     # a permutation plus padding rows
    assert np.all(ts % 32 == 0) and np.all(np.diff(ts) == ts[1] - ts[0])
    tile = np.searchsorted(ts, pos, side="right") - 1
    step = (pos - ts[tile])//32
    lane = pos % 32
    rl, ru = pos[l], pos[u]
    same = tile[l] == tile[u]
    assert np.all(step[u][same] == step[l][same] + 1)
    d = (u - l)[same]
    dl = (lane[u] - lane[l])[same]
    assert np.all(dl[d == 1] == 0)                                  # long axis (x, the longest): own lane
    assert np.all(dl[d == nx] == 1) and np.all(dl[d == nx*ny] == 4)  # a1 = y: lane + 1, a2 = z: lane + w1
    # across pencils: the producer is an earlier tile (forward sweeps take tiles in ascending order)
    assert np.all(tile[l][~same] < tile[u][~same])
    # a general mesh keeps the tile schedule (no pencils)
    m = meshes.tet_box(4, 3, 3)
    lt, ut = m.lower_upper()
    Pt = plan_export(lt, ut, m.nCells, tileRows=256, tileMode=3)
    assert Pt["info"]["nRows"] >= m.nCells and Pt["info"]["maxRounds"] > 0
