"""GPU parity tests (-m gpu): the CUDA path through the C ABI against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): Amul / DIC factor / DIC apply bit-exact (same operand order, no FMA);
PCG: identical iteration counts, per-iteration residual history within 1e-10 relative, final field within
1e-9 relative L-infinity (only the global reductions are re-associated on the GPU)."""
import numpy as np
import pytest

import oracle
import fluidstructureinteraction_b200 as pkg
from fluidstructureinteraction_b200 import cases, meshes
from conftest import FIXTURE_MESHES, load_fixture_mesh, load_golden, synthetic_matrix

pytestmark = pytest.mark.gpu

HIST_RTOL = 1e-10
# Two summation orders of the same doubles cannot agree below eps x the largest intermediate: once the residual
# has dropped many decades the 1e-10 RELATIVE bar gets an absolute floor of 1e-13 x the largest residual seen.
HIST_FLOOR = 1e-13
FIELD_RTOL = 1e-9


def hist_close(hg, hc, init):
    scale = max(float(np.max(np.abs(hc))), abs(init)) if len(hc) else abs(init)
    return np.all(np.abs(hg - hc) <= HIST_RTOL*np.abs(hc) + HIST_FLOOR*scale)


def check_kernels(l, u, n, diag, upper, b):
    g = pkg.GpuPCG(l, u, n)
    g.set_matrix(diag, upper)
    xt = np.sin(0.37*np.arange(n))
    assert np.array_equal(g.amul(xt), oracle.amul(l, u, diag, upper, xt))
    rD = oracle.dic_factor(l, u, diag, upper)
    assert np.array_equal(g.dic_factor(), rD)
    assert np.array_equal(g.precondition(b), oracle.precondition(l, u, upper, rD, b))
    return g


def check_solve(g, l, u, diag, upper, b, x0, tol, relTol, minIter=0, maxIter=1000):
    xg, xc = x0.copy(), x0.copy()
    pg, hg = g.solve(xg, b, tol, relTol, minIter, maxIter)
    pc, hc = oracle.pcg_solve(l, u, diag, upper, xc, b, tol, relTol, minIter, maxIter)
    assert pg["nIterations"] == pc["nIterations"], (pg, pc)
    assert pg["converged"] == pc["converged"] and pg["singular"] == pc["singular"]
    assert pg["normFactor"] == pytest.approx(pc["normFactor"], rel=1e-13)
    assert pg["initialResidual"] == pytest.approx(pc["initialResidual"], rel=HIST_RTOL, abs=1e-300)
    if len(hc):
        assert hist_close(hg, hc, pc["initialResidual"])
        assert pg["finalResidual"] == pytest.approx(pc["finalResidual"], rel=HIST_RTOL)
    scale = np.max(np.abs(xc))
    assert np.max(np.abs(xg - xc)) <= FIELD_RTOL*scale + 1e-300
    return pg, pc


@pytest.mark.parametrize("dims", [(1, 1, 1), (2, 1, 1), (33, 1, 1), (5, 4, 3), (24, 8, 12), (40, 10, 20)])
def test_hex_kernels_and_solve(dims):
    nx, ny, nz = dims
    l, u = meshes.hex_ldu(nx, ny, nz)
    n = nx*ny*nz
    diag, upper, b, x0 = synthetic_matrix(l, u, n, 11)
    g = check_kernels(l, u, n, diag, upper, b)
    check_solve(g, l, u, diag, upper, b, x0, 1e-9, 0.0)
    check_solve(g, l, u, diag, upper, b, x0, 1e-9, 0.01)
    g.close()


@pytest.mark.parametrize("name", FIXTURE_MESHES)
def test_shipped_meshes_against_golden(name):
    """The reference's own polyMeshes: CUDA path vs the committed golden vectors (and the live oracle)."""
    m = load_fixture_mesh(name)
    gold = load_golden(name)
    l, u = m.lower_upper()
    diag, upper, b, x0 = synthetic_matrix(l, u, m.nCells, int(gold["seed"]))
    g = check_kernels(l, u, m.nCells, diag, upper, b)
    xt = np.sin(0.37*np.arange(m.nCells))
    assert np.array_equal(g.amul(xt), gold["Ax"])
    assert np.array_equal(g.dic_factor(), gold["rD"])
    assert np.array_equal(g.precondition(b), gold["wA"])
    x = x0.copy()
    pg, hg = g.solve(x, b, 1e-9, 0.0)
    assert pg["nIterations"] == int(gold["nIterations"])
    assert hist_close(hg, gold["history"], float(gold["initialResidual"]))
    assert np.max(np.abs(x - gold["x"])) <= FIELD_RTOL*np.max(np.abs(gold["x"]))
    # shipped fvSolution controls (beamInCrossFlow/solid/system/fvSolution:17-32): tol 1e-9, relTol 0.1
    check_solve(g, l, u, diag, upper, b, x0, 1e-9, 0.1)
    g.close()


def test_cantilever_1M_slice_and_controls():
    """The benchmark system (steel cantilever, traction load) at a size the oracle solves in seconds."""
    s = cases.cantilever_system(100, 25, 50)        # 125 k cells, same aspect as the 1 M case
    l, u, diag, upper, b, n = s["l"], s["u"], s["diag"], s["upper"], s["b"], s["nCells"]
    g = check_kernels(l, u, n, diag, upper, b)
    x0 = np.zeros(n)
    pg, pc = check_solve(g, l, u, diag, upper, b, x0, 1e-9, 0.01)      # plateHole controls
    assert pg["nIterations"] > 10
    check_solve(g, l, u, diag, upper, b, x0, 1e-9, 0.1)                # FSI solid controls
    # edge cases of stop(): zero rhs, maxIter cap, minIter
    pz, _ = g.solve(np.zeros(n), np.zeros(n), 1e-9, 0.01)
    assert pz["nIterations"] == 0 and pz["converged"] == 1 and pz["initialResidual"] == 0.0
    check_solve(g, l, u, diag, upper, b, x0, 1e-30, 0.0, maxIter=5)
    check_solve(g, l, u, diag, upper, b, x0, 1e-9, 0.999999, minIter=3)
    check_solve(g, l, u, diag, upper, b, x0, 1e-9, 0.0, maxIter=13)    # not a multiple of the check batch
    g.close()


def test_component_reuse_upper_null():
    """fvMatrix<vector>::solve: three component solves share upper, only diag/b change (upper=None)."""
    s = cases.cantilever_system(20, 6, 10)
    l, u, diag, upper, b, n = s["l"], s["u"], s["diag"], s["upper"], s["b"], s["nCells"]
    g = pkg.GpuPCG(l, u, n)
    g.set_matrix(diag, upper)
    rng = np.random.default_rng(5)
    for cmpt in range(3):
        dc = diag*(1.0 + 0.1*cmpt) + rng.uniform(0, 1e6, n)
        bc = rng.standard_normal(n)
        g.set_matrix(dc, None)
        check_solve(g, l, u, dc, upper, bc, np.zeros(n), 1e-9, 0.01)
    g.close()


def test_singular_and_errors():
    s = cases.cantilever_system(8, 4, 4)
    l, u, diag, upper, b, n = s["l"], s["u"], s["diag"], s["upper"], s["b"], s["nCells"]
    g = pkg.GpuPCG(l, u, n)
    with pytest.raises(pkg.GpuPcgError) as e:
        g.solve(np.zeros(n), b)
    assert e.value.code == -5                       # solve before set_matrix
    # wApA underflows to 0 -> checkSingularity stops the loop before x is touched
    g.set_matrix(np.ones(n), np.zeros_like(upper))
    pg, pc = check_solve(g, l, u, np.ones(n), np.zeros_like(upper), np.full(n, 1e-301), np.zeros(n), 0.0, 0.0)
    assert pg["singular"] == 1 and pg["nIterations"] == 0
    with pytest.raises(pkg.GpuPcgError) as e:
        g.set_matrix(diag, upper, lower=upper*0.5)
    assert e.value.code == -7                       # asymmetric
    g.close()
    with pytest.raises(pkg.GpuPcgError) as e:
        pkg.GpuPCG(np.array([0, 1, 0], dtype=np.int32), np.array([1, 2, 2], dtype=np.int32), 3)
    assert e.value.code == -2


def test_tet_mesh_with_jitter():
    m = meshes.tet_box(10, 4, 6, jitter=0.15)
    l, u = m.lower_upper()
    diag, upper, b, x0 = synthetic_matrix(l, u, m.nCells, 21)
    g = check_kernels(l, u, m.nCells, diag, upper, b)
    check_solve(g, l, u, diag, upper, b, x0, 1e-9, 0.0)
    g.close()


def test_amul_with_interface_coefficients():
    """Processor-patch term of Amul with explicit neighbour values: Ax[fc[i]] -= bou[i]*xNbr[i] in patch order,
    including a cell that touches two patches and a patch that hits one cell twice."""
    s = cases.cantilever_system(6, 4, 4)
    l, u, diag, upper, n = s["l"], s["u"], s["diag"], s["upper"], s["nCells"]
    fcs = [np.array([0, 5, 5, 17], dtype=np.int32), np.array([5, 95, 3], dtype=np.int32)]
    rng = np.random.default_rng(9)
    bou = [rng.uniform(1, 2, fc.shape[0])*1e9 for fc in fcs]
    xn = rng.standard_normal(sum(fc.shape[0] for fc in fcs))
    g = pkg.GpuPCG(l, u, n, patchFaceCells=fcs, patchNbrRank=[1, 2])
    g.set_matrix(diag, upper, patchBouCoeffs=bou)
    xt = np.sin(0.37*np.arange(n))
    ref = oracle.amul(l, u, diag, upper, xt, fcs, bou, xn)
    assert np.array_equal(g.amul(xt, xn), ref)
    g.close()


def test_large_size_properties():
    """At the full 1 M-cell benchmark size: size-independent properties instead of the (slow) oracle solve --
    Amul linearity/symmetry, DIC exactness on the factor, and the true residual of the returned solution."""
    s = cases.cantilever_system(200, 50, 100)
    l, u, diag, upper, b, n = s["l"], s["u"], s["diag"], s["upper"], s["b"], s["nCells"]
    g = pkg.GpuPCG(l, u, n)
    g.set_matrix(diag, upper)
    rng = np.random.default_rng(1)
    v, w = rng.standard_normal(n), rng.standard_normal(n)
    Av, Aw = g.amul(v), g.amul(w)
    assert abs(np.dot(w, Av) - np.dot(v, Aw)) <= 1e-12*abs(np.dot(w, Av))          # symmetry
    assert np.array_equal(Av, oracle.amul(l, u, diag, upper, v))                      # still bit-exact (oracle Amul is fast)
    rD = g.dic_factor()
    assert np.array_equal(rD, oracle.dic_factor(l, u, diag, upper))
    assert np.array_equal(g.precondition(v), oracle.precondition(l, u, upper, rD, v))
    x = np.zeros(n)
    pg, hg = g.solve(x, b, 1e-9, 0.01)
    assert pg["converged"] == 1 and 0 < pg["nIterations"] < 1000
    r = b - g.amul(x)
    assert np.sum(np.abs(r))/pg["normFactor"] == pytest.approx(pg["finalResidual"], rel=1e-6)
    assert pg["finalResidual"] <= 0.01*pg["initialResidual"]
    g.close()


@pytest.mark.parametrize("ordered,p2p,mesh", [(1, None, "hex"), (1, "0", "hex"), (0, None, "hex"), (1, None, "tet"), (0, None, "tube")])
def test_two_gpu_parity_against_oracle_on_same_partition(ordered, p2p, mesh):
    """Needs >= 2 GPUs (gpurun --gpus 2): halo swaps + cross-rank reductions vs oracle.pcg_solve_multi.  ordered = 1: global
    sums in the reference's order -> history and solution bit-identical (mailboxes, and NCCL's all-reduce with p2p = "0");
    ordered = 0: the production tree reductions (bars in tools/multigpu_parity.py).  mesh: structured hex blocks (pencil
    sweeps), or UNSTRUCTURED sub-domains -- a jittered tet box / the shipped 3dTube solid mesh cut by `method simple`."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, ORDERED=str(ordered), MESH=mesh)
    if p2p is not None:
        env["GPUPCG_P2P_REDUCE"] = p2p
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tools", "multigpu_parity.py")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "MULTIGPU-PARITY-OK" in r.stdout, r.stdout[-3000:]


@pytest.mark.parametrize("kind", ["hex", "tet"])
def test_ordered_sums_mode_is_bit_identical_to_the_oracle(kind, monkeypatch):
    """GPUPCG_ORDERED_SUMS=1: gSumProd / gSumMag / gSum formed in the caller's cell order (csrc/kernels_ordered.cuh), every
    other kernel the production one -> residual history, iteration count and solution equal the oracle's bit for bit, for
    the scalar solve and for every component of the batched one.  What separates the production path from the reference is
    therefore the association of those sums alone."""
    monkeypatch.setenv("GPUPCG_ORDERED_SUMS", "1")
    if kind == "hex":
        s = cases.cantilever_system(40, 12, 16)
        l, u, diag, upper, b, n = s["l"], s["u"], s["diag"], s["upper"], s["b"], s["nCells"]
    else:
        from fluidstructureinteraction_b200 import meshes
        m = meshes.tet_box(8, 5, 6, jitter=0.15)
        l, u = m.lower_upper(); n = m.nCells
        diag, upper, b, _ = synthetic_matrix(l, u, n, 77)
    g = pkg.GpuPCG(l, u, n, device=0)
    g.set_matrix(diag, upper)
    for x0 in (np.zeros(n), 0.3*np.sin(0.11*np.arange(n))):        # xRef = 0 and xRef != 0 (normFactor's Amul of the xRef field)
        for relTol in (0.01, 0.0):
            xg, xc = x0.copy(), x0.copy()
            pg, hg = g.solve(xg, b, tolerance=1e-9, relTol=relTol)
            pc, hc = oracle.pcg_solve(l, u, diag, upper, xc, b, tolerance=1e-9, relTol=relTol)
            assert pg["nIterations"] == pc["nIterations"] and (pg["nIterations"] > 5 or x0.any())
            assert np.array_equal(hg, hc) and np.array_equal(xg, xc)
            assert pg["initialResidual"] == pc["initialResidual"] and pg["finalResidual"] == pc["finalResidual"]
            assert pg["normFactor"] == pc["normFactor"]
    dC = np.ascontiguousarray(np.stack([diag, diag*1.25, diag*1.5], axis=1))
    bC = np.ascontiguousarray(np.stack([b, -0.5*b, np.cos(0.11*np.arange(n))*np.abs(b).max()], axis=1))
    g.set_matrix_vector(dC, upper)
    xv = np.zeros((n, 3))
    pv, hv = g.solve_vector(xv, bC, tolerance=1e-9, relTol=0.0)
    for k in range(3):
        xk = np.zeros(n)
        pk, hk = oracle.pcg_solve(l, u, np.ascontiguousarray(dC[:, k]), upper, xk, np.ascontiguousarray(bC[:, k]), tolerance=1e-9, relTol=0.0)
        assert pv[k]["nIterations"] == pk["nIterations"]
        assert np.array_equal(hv[k], hk) and np.array_equal(xv[:, k], xk), k
    g.close()


def random_ldu(n, extra, seed):
    """A connected random graph in upper-triangular face order with some rows of high degree on both sides
    (exercises the > 4 faces-per-side tail of the tile kernels and non-mesh-like schedules)."""
    rng = np.random.default_rng(seed)
    pairs = {(i, i + 1) for i in range(n - 1)}
    a = rng.integers(0, n, size=extra); b = rng.integers(0, n, size=extra)
    for i, j in zip(a, b):
        if i != j:
            pairs.add((min(i, j), max(i, j)))
    hub = n//2
    for j in rng.integers(0, n, size=12):           # one hub row with many faces on both sides
        if j != hub:
            pairs.add((min(hub, j), max(hub, j)))
    arr = np.array(sorted(pairs), dtype=np.int32)
    return arr[:, 0].copy(), arr[:, 1].copy()


@pytest.mark.parametrize("n,extra", [(300, 900), (5000, 12000)])
def test_random_graph_with_high_degree_rows(n, extra):
    l, u = random_ldu(n, extra, 4)
    diag, upper, b, x0 = synthetic_matrix(l, u, n, 31)
    g = check_kernels(l, u, n, diag, upper, b)
    check_solve(g, l, u, diag, upper, b, x0, 1e-9, 0.0)
    g.close()


def _scalar_system(kind):
    """the other symmetric systems entering the same solver boundary."""
    from fluidstructureinteraction_b200 import cases, geometry
    m = meshes.hex_box(16, 6, 10)
    g = geometry.compute(m)
    if kind == "TEqn":          # thermalStressFoam TEqn.H:9-15: ddt(rhoC T) - laplacian(k T), hot and cold walls, rest adiabatic
        s = cases.scalar_laplacian_system(m, g, 45.0, fixedValue={"fixed": 300.0, "loaded": 500.0}, rate=7854.0*434.0/1000.0)
        s["source"] = s["source"] + (7854.0*434.0/1000.0)*g.V*300.0           # ddt source: rate*V*T_old
        x0 = np.full(m.nCells, 300.0)
    else:                        # consistentIcoFlow.C:213-224: laplacian(1/A, p) == div(phi), all Neumann + setReference
        s = cases.scalar_laplacian_system(m, g, 2.5e-3, refCell=0, refValue=0.0)
        rng = np.random.default_rng(17)
        dv = rng.standard_normal(m.nCells); dv -= dv.mean()                  # a divergence field that sums to zero
        s["source"] = s["source"] + dv*g.V
        x0 = np.zeros(m.nCells)
    # fvScalarMatrix::solve: addBoundaryDiag / addBoundarySource of the non-coupled patches
    diag = s["diag"].copy(); b = s["source"].copy()
    for fc, ic, bc in s["patches"]:
        np.add.at(diag, fc, ic); np.add.at(b, fc, bc)
    return m, s, diag, b, x0


@pytest.mark.parametrize("kind", ["TEqn", "pEqn"])
def test_other_scalar_systems_through_the_same_solver(kind):
    m, s, diag, b, x0 = _scalar_system(kind)
    g = check_kernels(s["l"], s["u"], m.nCells, diag, s["upper"], b)
    pg, pc = check_solve(g, s["l"], s["u"], diag, s["upper"], b, x0, 1e-9, 0.0)
    assert pg["converged"] and pg["nIterations"] >= 3
    check_solve(g, s["l"], s["u"], diag, s["upper"], b, x0, 1e-6, 0.05)      # the shipped fluid / thermal tolerances
    g.close()


@pytest.mark.parametrize("kind", ["pencil", "tiles"])
def test_spin_watchdog_returns_an_error_instead_of_hanging(kind, monkeypatch):
    """every device-side wait is rooted in a poll of global memory; one that makes no progress for
    GPUPCG_SPIN_TIMEOUT_MS gives up, the kernels run to their end and the call returns GPUPCG_ERR_STATE.  Forced here with a
    zero timeout checked at every failed poll (any hand-off that is not instantaneous trips it); a handle created afterwards
    with the default timeout solves normally (the fault word is per process and is cleared when it is reported)."""
    s = cases.cantilever_system(40, 12, 16)
    l, u, diag, upper, b, n = s["l"], s["u"], s["diag"], s["upper"], s["b"], s["nCells"]
    monkeypatch.setenv("GPUPCG_SPIN_TIMEOUT_MS", "0")
    monkeypatch.setenv("GPUPCG_SPIN_CHECK_MASK", "0")
    if kind == "tiles":
        monkeypatch.setenv("GPUPCG_TILE_MODE", "2")
    g = pkg.GpuPCG(l, u, n, device=0)
    g.set_matrix(diag, upper)
    with pytest.raises(pkg.GpuPcgError) as e:
        for _ in range(5):                       # a sweep whose every hand-off happened to be ready is possible, five in a row are not
            g.solve(np.zeros(n), b, tolerance=1e-12, relTol=0.0)
    assert "timed out" in str(e.value)
    g.close()
    monkeypatch.delenv("GPUPCG_SPIN_TIMEOUT_MS"); monkeypatch.delenv("GPUPCG_SPIN_CHECK_MASK")
    g = pkg.GpuPCG(l, u, n, device=0)
    g.set_matrix(diag, upper)
    x = np.zeros(n)
    pg, _ = g.solve(x, b, tolerance=1e-9, relTol=0.01)
    xc = np.zeros(n)
    pc, _ = oracle.pcg_solve(l, u, diag, upper, xc, b, tolerance=1e-9, relTol=0.01)
    assert pg["nIterations"] == pc["nIterations"] and np.max(np.abs(x - xc)) <= 1e-9*np.max(np.abs(xc))
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,ordered", [("hex", 0), ("hex", 1), ("tet", 0), ("tet", 1)])
def test_concurrent_component_solves(kind, ordered, monkeypatch):
    """GPUPCG_VECTOR_CONCURRENT=1: the three systems of a vector solve as three scalar solves in flight at once (own stream,
    host thread and vectors per component; plan tables and upper() shared -- csrc/gpupcg_vec.cuh vec_solve_concurrent).
    Each component must be the scalar solve of that system: same iteration counts and history as gpupcg_solve on a separate
    handle; with ordered sums bit-identical to the oracle.  Default-on only for structured plans >= 100 k cells, forced here
    (pencil sweeps on the hex block, tile sweeps on the tets)."""
    monkeypatch.setenv("GPUPCG_VECTOR_CONCURRENT", "1")
    if ordered:
        monkeypatch.setenv("GPUPCG_ORDERED_SUMS", "1")
    if kind == "hex":
        s = cases.cantilever_system(40, 12, 16)
        l, u, diag, upper, b, n = s["l"], s["u"], s["diag"], s["upper"], s["b"], s["nCells"]
    else:
        from fluidstructureinteraction_b200 import meshes
        m = meshes.tet_box(8, 5, 6, jitter=0.15)
        l, u = m.lower_upper(); n = m.nCells
        diag, upper, b, _ = synthetic_matrix(l, u, n, 77)
    dC = np.ascontiguousarray(np.stack([diag, diag*1.25, diag*1.5], axis=1))
    bC = np.ascontiguousarray(np.stack([b, -0.5*b, np.cos(0.11*np.arange(n))*np.abs(b).max()], axis=1))
    g = pkg.GpuPCG(l, u, n, device=0)
    for rep in range(2):                                   # second pass: new diagonal, upper() re-used (upper = None)
        if rep:
            dC = np.ascontiguousarray(dC*np.array([1.0, 1.1, 0.9]))
        g.set_matrix_vector(dC, upper if rep == 0 else None)
        xv = np.zeros((n, 3))
        pv, hv = g.solve_vector(xv, bC, tolerance=1e-9, relTol=0.0)
        g1 = pkg.GpuPCG(l, u, n, device=0)
        for k in range(3):
            dk, bk = np.ascontiguousarray(dC[:, k]), np.ascontiguousarray(bC[:, k])
            g1.set_matrix(dk, upper)
            xk = np.zeros(n)
            pk, hk = g1.solve(xk, bk, tolerance=1e-9, relTol=0.0)
            assert pv[k]["nIterations"] == pk["nIterations"] > 5
            assert np.allclose(hv[k], hk, rtol=1e-10, atol=0.0) and np.abs(xv[:, k] - xk).max() <= 1e-11*np.abs(xk).max()
            if ordered:
                xo = np.zeros(n)
                po, ho = oracle.pcg_solve(l, u, dk, upper, xo, bk, tolerance=1e-9, relTol=0.0)
                assert po["nIterations"] == pv[k]["nIterations"]
                assert np.array_equal(hv[k], ho) and np.array_equal(xv[:, k], xo), k
        g1.close()
    assert g.launch_count() > 0
    g.close()
