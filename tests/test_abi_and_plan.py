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
    declared = set(re.findall(r"\b(gpupcg_[a-z_0-9]+)\s*\(", header))
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
    """numpy emulation of the row walks of k_amul / k_dic_factor / k_dic_sweeps on the exported schedule."""
    rowOld, slices = plan["rowOld"], plan["slices"]
    Ucol, UfaceOld, Lidx, Lcol = plan["Ucol"], plan["UfaceOld"], plan["Lidx"], plan["Lcol"]
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

    def entries(r):
        s, lane = divmod(r, 32)
        offU, wU, offL, wL = slices[s]
        lo = [(Uval[Lidx[offL + j*32 + lane]], Lcol[offL + j*32 + lane]) for j in range(wL)
              if Lcol[offL + j*32 + lane] >= 0]
        up = [(Uval[offU + j*32 + lane], Ucol[offU + j*32 + lane]) for j in range(wU)
              if Ucol[offU + j*32 + lane] >= 0]
        return lo, up

    def amul(x):
        xp = rows_in(x)
        Ax = np.zeros(nR)
        for r in range(nR):
            lo, up = entries(r)
            acc = d[r]*xp[r]
            for c, col in lo + up:
                acc = acc + c*xp[col]
            Ax[r] = acc
        return rows_out(Ax)

    def factor():
        piv = np.zeros(nR)
        for r in range(nR):             # rows are in topological order
            lo, _ = entries(r)
            acc = d[r]
            for c, col in lo:
                acc = acc - (c*c)/piv[col]
            piv[r] = acc
        return 1.0/piv

    def precondition(rD, rA):
        rp = rows_in(rA)
        y = np.zeros(nR)
        for r in range(nR):
            lo, _ = entries(r)
            acc = rD[r]*rp[r]
            for c, col in lo:
                acc = acc - (rD[r]*c)*y[col]
            y[r] = acc
        z = np.zeros(nR)
        for r in range(nR - 1, -1, -1):
            _, up = entries(r)
            acc = y[r]
            for c, col in reversed(up):
                acc = acc - (rD[r]*c)*z[col]
            z[r] = acc
        return rows_out(z)

    return amul, factor, precondition, rows_out


def check_schedule(l, u, n, seed=3):
    plan = pkg.plan_export(l, u, n)
    info = plan["info"]
    rowOld = plan["rowOld"]
    # permutation is a bijection onto the real rows, levels are warp-aligned
    real = rowOld[rowOld >= 0]
    assert np.array_equal(np.sort(real), np.arange(n))
    assert info["nRows"] % 32 == 0 and info["nSlices"]*32 == info["nRows"]
    pos = np.empty(n, dtype=np.int64)
    pos[rowOld[rowOld >= 0]] = np.flatnonzero(rowOld >= 0)
    # topological: every face goes from an earlier slice to a later one (so a warp never waits on itself)
    assert np.all(pos[l] // 32 < pos[u] // 32)
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
    info = check_schedule(l, u, nx*ny*nz)
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
