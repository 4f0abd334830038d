"""D-equation assembly: the oracle against closed forms on the orthogonal hex cantilever (CPU), and the GPU kernels
against the oracle bit for bit on hex and jittered-tet meshes (GPU)."""
import numpy as np
import pytest

import oracle
import fluidstructureinteraction_b200 as pkg
from fluidstructureinteraction_b200 import cases, geometry, meshes

PATCHES = {"fixed": "fixedDisplacement", "loaded": "tractionDisplacement", "free": "tractionDisplacement"}


def fields(mesh, seed, scale=1e-4):
    rng = np.random.default_rng(seed)
    nC, nF, nBF = mesh.nCells, mesh.nFaces, mesh.nFaces - mesh.nInternalFaces
    mu, lam = oracle.lame(cases.STEEL["E"], cases.STEEL["nu"])
    f = dict(mu=np.full(nC, mu)*rng.uniform(0.9, 1.1, nC), lam=np.full(nC, lam)*rng.uniform(0.9, 1.1, nC),
             rho=np.full(nC, cases.STEEL["rho"]), g=np.array([0.0, 0.0, -9.81]),
             D=scale*rng.standard_normal((nC, 3)), gradD=scale*rng.standard_normal((nC, 9)),
             gradDf=scale*rng.standard_normal((nF, 9)), traction=np.zeros((nBF, 3)), pressure=rng.uniform(0, 1e3, nBF),
             fixedValue=scale*rng.standard_normal((nBF, 3)))
    p = mesh.patch("loaded")
    f["traction"][p["start"] - mesh.nInternalFaces:p["start"] - mesh.nInternalFaces + p["size"]] = [0.0, 0.0, -1e4]
    return f


def test_geometry_of_uniform_hex():
    m = meshes.hex_box(6, 4, 5)
    g = geometry.compute(m)
    hx, hy, hz = 2.0/6, 0.5/4, 1.0/5
    assert np.allclose(g.V, hx*hy*hz, rtol=1e-13)
    assert np.allclose(g.weights, 0.5, rtol=1e-13)
    l, u = m.lower_upper()
    d = u.astype(np.int64) - l
    assert np.allclose(g.deltaCoeffs[:m.nInternalFaces][d == 1], 1/hx, rtol=1e-12)
    assert np.allclose(g.magSf[:m.nInternalFaces][d == 1], hy*hz, rtol=1e-12)
    p = m.patch("fixed")
    assert np.allclose(g.deltaCoeffs[p["start"]:p["start"] + p["size"]], 2/hx, rtol=1e-12)
    assert np.allclose(g.Sf[p["start"]:p["start"] + p["size"]], [-hy*hz, 0, 0], atol=1e-15)


def test_oracle_assembly_reduces_to_closed_form_on_orthogonal_hex():
    """With uniform steel, D = 0 and gradDf = 0 the assembled z-component system is cases.cantilever_system()."""
    nx, ny, nz = 6, 4, 5
    m = meshes.hex_box(nx, ny, nz)
    g = geometry.compute(m)
    f = fields(m, 1)
    mu, lam = oracle.lame(cases.STEEL["E"], cases.STEEL["nu"])
    f["mu"][:] = mu; f["lam"][:] = lam
    for k in ("D", "gradD", "gradDf", "fixedValue", "pressure"):
        f[k][...] = 0.0
    f["g"][:] = 0.0
    a = oracle.assemble_D(m, g, PATCHES, f["mu"], f["lam"], f["rho"], f["g"], f["D"], f["gradD"], f["gradDf"],
                          f["traction"], f["pressure"], f["fixedValue"], 1.0)
    s = cases.cantilever_system(nx, ny, nz)
    assert np.allclose(a["upper"], s["upper"], rtol=1e-12)
    diag = a["diag"].copy()
    src = a["source"][:, 2].copy()
    nIF = m.nInternalFaces
    np.add.at(diag, m.owner[nIF:], a["internalCoeffs"][:, 2])
    np.add.at(src, m.owner[nIF:], a["boundaryCoeffs"][:, 2])
    assert np.allclose(diag, s["diag"], rtol=1e-12)
    assert np.allclose(src, s["b"], rtol=1e-12, atol=1e-9)
    assert np.all(a["source"] == 0.0)
    # orthogonal mesh: the skew correction vanishes identically, so random D / gradD must not change the source
    f2 = fields(m, 2)
    a2 = oracle.assemble_D(m, g, PATCHES, f["mu"], f["lam"], f["rho"], f["g"], f2["D"], f2["gradD"], f["gradDf"],
                           f["traction"], f["pressure"], f["fixedValue"], 1.0)
    assert np.max(np.abs(a2["source"])) <= 1e-6*np.max(np.abs(a2["upper"]))*1e-4


def test_oracle_sigma_and_lame():
    mu, lam = oracle.lame(2e11, 0.3)
    assert mu == 2e11/(2.0*(1.0 + 0.3)) and lam == 0.3*2e11/((1.0 + 0.3)*(1.0 - 2.0*0.3))
    assert oracle.lame(2e11, 0.3, True)[1] == 0.3*2e11/((1.0 + 0.3)*(1.0 - 0.3))
    G = np.arange(9.0)[None, :]
    eps, sig = oracle.sigma(np.array([2.0]), np.array([3.0]), G)
    assert np.array_equal(eps[0], [0, 2, 4, 4, 6, 8])
    assert np.array_equal(sig[0], [0 + 36, 8, 16, 16 + 36, 24, 32 + 36])


PATCHES_SYM = {"fixed": "fixedDisplacement", "loaded": "tractionDisplacement", "free": "symmetryDisplacement"}


def test_oracle_symmetry_plane_coefficients():
    """symmetryDisplacement on an axis-aligned plane: the normal component is pinned (internalCoeffs = gamma|Sf|delta
    on it, zero on the tangential ones) and a displacement field that is already symmetric gives no boundary source."""
    m = meshes.hex_box(4, 3, 3)
    g = geometry.compute(m)
    f = fields(m, 2)
    f["gradD"][...] = 0.0
    a = oracle.assemble_D(m, g, PATCHES_SYM, f["mu"], f["lam"], f["rho"], f["g"], f["D"], f["gradD"], f["gradDf"],
                          f["traction"], f["pressure"], f["fixedValue"], 1.0)
    p = m.patch("free"); nIF = m.nInternalFaces
    sl = slice(p["start"] - nIF, p["start"] - nIF + p["size"])
    n = g.Sf[p["start"]:p["start"] + p["size"]]/g.magSf[p["start"]:p["start"] + p["size"], None]
    ic = a["internalCoeffs"][sl]
    assert np.all(ic[np.abs(n) < 0.5] == 0.0) and np.all(ic[np.abs(n) > 0.5] > 0.0)
    # boundaryCoeffs = gamma|Sf|*(snGrad - gIC*D_P): with gradD = 0, snGrad = -delta*n(n.D_P) and gIC*D_P = -delta*|n|*D_P
    # cancel on the normal component
    assert np.max(np.abs(a["boundaryCoeffs"][sl])) <= 1e-9*np.max(np.abs(ic))*np.max(np.abs(f["D"]))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["hex", "tet"])
def test_gpu_assembly_chunked_bit_exact(kind, monkeypatch):
    """The assembly runs chunk by chunk (per-face fluxes in an L2-sized ring; faces cut by a chunk boundary are
    recomputed by the later chunk): many small chunks must give the same bits as the oracle."""
    monkeypatch.setenv("GPUPCG_ASM_CHUNK_CELLS", "100")
    m = meshes.hex_box(12, 6, 8) if kind == "hex" else meshes.tet_box(6, 4, 5, jitter=0.15)
    g = geometry.compute(m)
    f = fields(m, 7)
    gm = pkg.GpuMesh(m, g, PATCHES)
    a = gm.assemble_D(f["mu"], f["lam"], f["rho"], f["g"], f["D"], f["gradD"], f["gradDf"], f["traction"],
                      f["pressure"], f["fixedValue"], 0.5)
    r = oracle.assemble_D(m, g, PATCHES, f["mu"], f["lam"], f["rho"], f["g"], f["D"], f["gradD"], f["gradDf"],
                          f["traction"], f["pressure"], f["fixedValue"], 0.5)
    for k in ("upper", "diag", "source", "internalCoeffs", "boundaryCoeffs"):
        assert np.array_equal(a[k], r[k]), k
    gm.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["hex", "tet"])
def test_gpu_assembly_bit_exact(kind):
    m = meshes.hex_box(12, 6, 8) if kind == "hex" else meshes.tet_box(6, 4, 5, jitter=0.15)
    g = geometry.compute(m)
    f = fields(m, 5)
    gms = pkg.GpuMesh(m, g, PATCHES_SYM)
    a = gms.assemble_D(f["mu"], f["lam"], f["rho"], f["g"], f["D"], f["gradD"], f["gradDf"], f["traction"],
                       f["pressure"], f["fixedValue"], 1.0)
    r = oracle.assemble_D(m, g, PATCHES_SYM, f["mu"], f["lam"], f["rho"], f["g"], f["D"], f["gradD"], f["gradDf"],
                          f["traction"], f["pressure"], f["fixedValue"], 1.0)
    for k in ("upper", "diag", "source", "internalCoeffs", "boundaryCoeffs"):
        assert np.array_equal(a[k], r[k]), "symmetry " + k
    gms.close()
    gm = pkg.GpuMesh(m, g, PATCHES)
    for limitCoeff in (1.0, 0.5):
        a = gm.assemble_D(f["mu"], f["lam"], f["rho"], f["g"], f["D"], f["gradD"], f["gradDf"], f["traction"],
                          f["pressure"], f["fixedValue"], limitCoeff)
        r = oracle.assemble_D(m, g, PATCHES, f["mu"], f["lam"], f["rho"], f["g"], f["D"], f["gradD"], f["gradDf"],
                              f["traction"], f["pressure"], f["fixedValue"], limitCoeff)
        for k in ("upper", "diag", "source", "internalCoeffs", "boundaryCoeffs"):
            assert np.array_equal(a[k], r[k]), k
    if kind == "tet":
        assert np.max(np.abs(r["source"])) > 0          # the skew correction is live on the jittered mesh
    eps, sig = gm.sigma(f["mu"], f["lam"], f["gradD"])
    e2, s2 = oracle.sigma(f["mu"], f["lam"], f["gradD"])
    assert np.array_equal(eps, e2) and np.array_equal(sig, s2)
    gm.close()


@pytest.mark.gpu
def test_assembled_system_solves_through_gpupcg():
    """Assembly -> addBoundaryDiag/addBoundarySource -> gpuPCG, against the same chain on the oracle."""
    m = meshes.hex_box(16, 6, 8)
    g = geometry.compute(m)
    f = fields(m, 9, scale=1e-6)
    gm = pkg.GpuMesh(m, g, PATCHES)
    a = gm.assemble_D(f["mu"], f["lam"], f["rho"], f["g"], f["D"], f["gradD"], f["gradDf"], f["traction"], f["pressure"],
                      f["fixedValue"], 1.0)
    l, u = m.lower_upper()
    nIF = m.nInternalFaces
    solver = pkg.GpuPCG(l, u, m.nCells)
    for c in range(3):
        diag = a["diag"].copy(); src = a["source"][:, c].copy()
        np.add.at(diag, m.owner[nIF:], a["internalCoeffs"][:, c])
        np.add.at(src, m.owner[nIF:], a["boundaryCoeffs"][:, c])
        solver.set_matrix(diag, a["upper"] if c == 0 else None)
        xg = np.zeros(m.nCells); xc = np.zeros(m.nCells)
        pg, hg = solver.solve(xg, src, 1e-9, 0.01)
        pc, hc = oracle.pcg_solve(l, u, diag, a["upper"], xc, src, 1e-9, 0.01)
        assert pg["nIterations"] == pc["nIterations"] and pg["nIterations"] > 3
        assert np.max(np.abs(xg - xc)) <= 1e-9*np.max(np.abs(xc))
    solver.close(); gm.close()


def linear_field(m, g):
    A = np.array([[1e-3, 2e-4, -3e-4], [5e-4, -2e-3, 1e-4], [7e-4, 3e-4, 4e-3]])      # D_j = sum_i x_i A_ij
    nIF = m.nInternalFaces
    return A, g.C@A, m.points@A, g.Cf[nIF:]@A, (g.Sf[nIF:]/g.magSf[nIF:, None])@A


def test_oracle_gradients_exact_for_linear_displacement():
    """fvc::grad(D,pointD) and fvc::fGrad(D,pointD) (fvcGradf.C) reproduce a linear field's gradient: everywhere on the
    orthogonal hex mesh; in the cells (any mesh) and on internal faces of the jittered tet mesh."""
    m = meshes.hex_box(5, 4, 3)
    g = geometry.compute(m)
    A, D, pD, fixed, pgrad = linear_field(m, g)
    G, Gb = oracle.grad(m, g, PATCHES, D, pD, np.tile(A.ravel(), (m.nCells, 1)), fixed, pgrad)
    Gf = oracle.fgrad(m, g, PATCHES, D, pD, G, fixed, pgrad, 1.0)
    for X in (G, Gb, Gf):
        assert np.max(np.abs(X - A.ravel())) < 1e-15
    m = meshes.tet_box(4, 3, 3, jitter=0.15)
    g = geometry.compute(m)
    A, D, pD, fixed, pgrad = linear_field(m, g)
    G, Gb = oracle.grad(m, g, PATCHES, D, pD, np.tile(A.ravel(), (m.nCells, 1)), fixed, pgrad)
    Gf = oracle.fgrad(m, g, PATCHES, D, pD, G, fixed, pgrad, 1.0)
    assert np.max(np.abs(G - A.ravel())) < 1e-15
    assert np.max(np.abs(Gf[:m.nInternalFaces] - A.ravel())) < 1e-14


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["hex", "tet"])
def test_gpu_gradients_bit_exact(kind):
    m = meshes.hex_box(12, 6, 8) if kind == "hex" else meshes.tet_box(6, 4, 5, jitter=0.15)
    g = geometry.compute(m)
    rng = np.random.default_rng(17)
    nBF = m.nFaces - m.nInternalFaces
    D = 1e-4*rng.standard_normal((m.nCells, 3)); pD = 1e-4*rng.standard_normal((m.nPoints, 3))
    Gold = 1e-4*rng.standard_normal((m.nCells, 9)); fixed = 1e-4*rng.standard_normal((nBF, 3))
    pgrad = 1e-4*rng.standard_normal((nBF, 3))
    for patches in (PATCHES, PATCHES_SYM):
        gm = pkg.GpuMesh(m, g, patches)
        for limitCoeff in (1.0, 0.5):
            G, Gb, Gf, ms = gm.gradients(D, pD, Gold, fixed, pgrad, limitCoeff)
            G2, Gb2 = oracle.grad(m, g, patches, D, pD, Gold, fixed, pgrad)
            Gf2 = oracle.fgrad(m, g, patches, D, pD, G2, fixed, pgrad, limitCoeff)
            assert np.array_equal(G, G2) and np.array_equal(Gb, Gb2) and np.array_equal(Gf, Gf2)
        gm.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["hex", "tet"])
def test_gpu_gradients_stay_resident_between_outer_iterations(kind):
    """Residency of evolve()'s outer loop (unsTotalLagrangianStress.C:846-926): gradients with NULL outputs leave
    grad(D) / the face gradient on the device, the next assembly (gradD = gradDf = NULL), the next gradients
    (gradDold = NULL: the device's grad(D) is the registered one) and sigma (gradD = NULL) read them there -- the same bits
    as the path through host arrays, and as the oracle."""
    m = meshes.hex_box(12, 6, 8) if kind == "hex" else meshes.tet_box(6, 4, 5, jitter=0.15)
    g = geometry.compute(m)
    f = fields(m, 7)
    rng = np.random.default_rng(23)
    pD1 = 1e-4*rng.standard_normal((m.nPoints, 3)); pD2 = 1e-4*rng.standard_normal((m.nPoints, 3))
    D1 = 1e-4*rng.standard_normal((m.nCells, 3)); D2 = 1e-4*rng.standard_normal((m.nCells, 3))
    gm = pkg.GpuMesh(m, g, PATCHES)
    with pytest.raises(pkg.GpuPcgError) as e:
        gm.assemble_D(f["mu"], f["lam"], f["rho"], f["g"], f["D"], None, None, f["traction"], f["pressure"], f["fixedValue"], 1.0)
    assert e.value.code == -5                       # nothing resident yet
    with pytest.raises(pkg.GpuPcgError):
        gm.fetch("gradD")
    # host path: assembly -> gradients -> assembly -> gradients, everything through host arrays
    a0 = gm.assemble_D(f["mu"], f["lam"], f["rho"], f["g"], f["D"], f["gradD"], f["gradDf"], f["traction"], f["pressure"], f["fixedValue"], 1.0)
    G1, Gb1, Gf1, _ = gm.gradients(D1, pD1, f["gradD"], f["fixedValue"], None, 1.0)
    a1 = gm.assemble_D(f["mu"], f["lam"], f["rho"], f["g"], D1, G1, Gf1, f["traction"], f["pressure"], f["fixedValue"], 1.0)
    G2, Gb2, Gf2, _ = gm.gradients(D2, pD2, G1, f["fixedValue"], None, 1.0)
    eps2, sig2 = gm.sigma(f["mu"], f["lam"], G2)
    gm.close()
    # resident path
    gm = pkg.GpuMesh(m, g, PATCHES)
    b0 = gm.assemble_D(f["mu"], f["lam"], f["rho"], f["g"], f["D"], f["gradD"], f["gradDf"], f["traction"], f["pressure"], f["fixedValue"], 1.0)
    gm.gradients_resident(D1, pD1, f["fixedValue"])
    b1 = gm.assemble_D(f["mu"], f["lam"], f["rho"], f["g"], D1, None, None, f["traction"], f["pressure"], f["fixedValue"], 1.0)
    gm.gradients_resident(D2, pD2, f["fixedValue"])
    eps3, sig3 = gm.sigma(f["mu"], f["lam"], None)
    for k in ("upper", "diag", "source", "internalCoeffs", "boundaryCoeffs"):
        assert np.array_equal(a0[k], b0[k]) and np.array_equal(a1[k], b1[k]), k
    assert np.array_equal(gm.fetch("gradD"), G2) and np.array_equal(gm.fetch("gradDf"), Gf2) and np.array_equal(gm.fetch("gradDb"), Gb2)
    assert np.array_equal(eps3, eps2) and np.array_equal(sig3, sig2)
    # and against the oracle
    r1 = oracle.assemble_D(m, g, PATCHES, f["mu"], f["lam"], f["rho"], f["g"], D1, G1, Gf1, f["traction"], f["pressure"], f["fixedValue"], 1.0)
    assert np.array_equal(b1["source"], r1["source"])
    gm.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["hex", "tet"])
def test_gpu_solve_on_the_resident_equation(kind):
    """gpupcg_solve_assembled_D: DEqn.solve() on the equation the assembly left on the device (addBoundaryDiag /
    addBoundarySource there, in fvMatrix's order; batched component solves) gives the same bits as assembly -> host ->
    fvMatrix<vector>::solve's diag / source per component -> gpupcg_set_matrix_vector + gpupcg_solve_vector."""
    m = meshes.hex_box(12, 6, 8) if kind == "hex" else meshes.tet_box(6, 4, 5, jitter=0.15)
    g = geometry.compute(m)
    f = fields(m, 9)
    nIF = m.nInternalFaces
    l, u = m.lower_upper()
    args = (f["mu"], f["lam"], f["rho"], f["g"], f["D"], f["gradD"], f["gradDf"], f["traction"], f["pressure"], f["fixedValue"], 1.0)
    gm = pkg.GpuMesh(m, g, PATCHES)
    solver = pkg.GpuPCG(l, u, m.nCells)
    with pytest.raises(pkg.GpuPcgError) as e:
        solver.solve_assembled_D(gm, np.zeros((m.nCells, 3)))
    assert e.value.code == -5                       # nothing assembled yet
    a = gm.assemble_D(*args)
    diagC = np.repeat(a["diag"][:, None], 3, axis=1)
    bC = a["source"].copy()
    own = np.asarray(m.owner[nIF:])
    for k in range(3):                              # fvMatrix order: boundary faces ascending
        np.add.at(diagC[:, k], own, a["internalCoeffs"][:, k])
        np.add.at(bC[:, k], own, a["boundaryCoeffs"][:, k])
    ref = pkg.GpuPCG(l, u, m.nCells)
    ref.set_matrix_vector(np.ascontiguousarray(diagC), a["upper"])
    D2 = np.ascontiguousarray(f["D"].copy())
    p2, _ = ref.solve_vector(D2, np.ascontiguousarray(bC), 1e-12, 0.0)
    gm.assemble_D(*args, copy_back=False)
    D1 = np.ascontiguousarray(f["D"].copy())
    p1 = solver.solve_assembled_D(gm, D1, 1e-12, 0.0)
    assert [p["nIterations"] for p in p1] == [p["nIterations"] for p in p2] and max(p["nIterations"] for p in p1) > 3
    assert np.array_equal(D1, D2)
    wrong = pkg.GpuPCG(l[:-1], u[:-1], m.nCells)
    with pytest.raises(pkg.GpuPcgError) as e:
        wrong.solve_assembled_D(gm, D1)
    assert e.value.code == -1                       # handle of another addressing
    for x in (solver, ref, wrong, gm):
        x.close()
