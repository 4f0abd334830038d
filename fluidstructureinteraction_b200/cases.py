"""Synthetic benchmark inputs at the lduMatrix::solver boundary.

cantilever_system() writes down, in closed form, the scalar LDU system that the first outer iteration
of unsTotalLagrangianStress::evolve() (unsTotalLagrangianStress.C:846-859) hands to the solver for the
loaded displacement component on the uniform orthogonal hex cantilever: -laplacian(2mu+lambda, D) with
fixedDisplacement (0 0 0) at x=0, tractionDisplacement (0 0 tz) at x=L, traction-free sides, D0 = 0,
gradDf = 0 (so every explicit term vanishes and deltaCoeffs = 1/h exactly).  It is an INPUT generator
for solver tests and the roofline runs; the general assembly (non-orthogonal, explicit terms) is the
GPU assembly path and its oracle.
"""
import numpy as np

from .meshes import hex_ldu

STEEL = dict(E=2.0e11, nu=0.3, rho=7854.0)      # run/stressFoam/plateHole/plateHole/constant/rheologyProperties:19-25


def lame(E, nu, planeStress=False):
    """constitutiveModel::mu / lambda (constitutiveModel.C:193-257)."""
    mu = E/(2.0*(1.0 + nu))
    if planeStress:
        lam = nu*E/((1.0 + nu)*(1.0 - nu))
    else:
        lam = nu*E/((1.0 + nu)*(1.0 - 2.0*nu))
    return mu, lam


def cantilever_system(nx, ny, nz, L=2.0, W=0.5, H=1.0, E=STEEL["E"], nu=STEEL["nu"], tz=-1.0e4):
    """Returns dict(l, u, diag, upper, b, nCells) for the z-displacement component."""
    mu, lam = lame(E, nu)
    gamma = 2.0*mu + lam
    hx, hy, hz = L/nx, W/ny, H/nz
    l, u = hex_ldu(nx, ny, nz)
    n = nx*ny*nz
    d = u.astype(np.int64) - l
    # gamma*|Sf|*deltaCoeffs per direction
    cx = gamma*(hy*hz)/hx
    cy = gamma*(hx*hz)/hy
    cz = gamma*(hx*hy)/hz
    mx = d == 1
    my = (d == nx) & ~mx & (ny > 1)
    coef = np.where(mx, cx, np.where(my, cy, cz))
    upper = -coef                                   # A = -laplacian  (A == B  ->  A - B)
    # negSumDiag (owner pass then neighbour pass; bincount sums each bin in input order)
    diag = np.bincount(l, weights=coef, minlength=n) + np.bincount(u, weights=coef, minlength=n)
    c = np.arange(n, dtype=np.int64)
    i = c % nx
    # fixedDisplacement at x = 0: internalCoeffs = gamma*|Sf|*deltaCoeffs_b, deltaCoeffs_b = 1/(hx/2)
    diag[i == 0] += gamma*(hy*hz)/(0.5*hx)
    b = np.zeros(n, dtype=np.float64)
    # tractionDisplacement at x = L: boundaryCoeffs = gamma*|Sf|*gradient, gradient = t/(2mu+lambda)
    b[i == nx - 1] += gamma*(hy*hz)*(tz/gamma)
    return dict(l=l, u=u, diag=diag, upper=upper, b=b, nCells=n, gamma=gamma)


def amul_bytes(nCells, nFaces):
    """Algorithmic bytes of one lduMatrix::Amul: 8(diag+x+Ax)N + (8+4+4)F."""
    return 24*nCells + 16*nFaces


def dic_bytes(nCells, nFaces):
    """DIC precondition, forward + backward: 8(rD+rA+wA)N + 2(8+4+4)F."""
    return 24*nCells + 32*nFaces


def pcg_iteration_bytes(nCells, nFaces):
    """One PCG iteration, reference data movement (unfused): Amul + DIC + 2 dots + pA update + x/r update."""
    return amul_bytes(nCells, nFaces) + dic_bytes(nCells, nFaces) + 16*nCells + 24*nCells + 16*nCells + 48*nCells


def cantilever_subdomain(nx, ny, nz, px, py, pz, rank, L=2.0, W=0.5, H=1.0, E=STEEL["E"], nu=STEEL["nu"], tz=-1.0e4):
    """The sub-domain `rank` of cantilever_system() decomposed `simple (px py pz)`, written down directly (a rank
    never builds the global 64 M-cell system).  Same layout as decompose.decompose_ldu: local cells in ascending
    global id, one processor patch per neighbouring rank (ascending), faces in global face order, bouCoeffs = -upper."""
    mu, lam = lame(E, nu)
    gamma = 2.0*mu + lam
    hx, hy, hz = L/nx, W/ny, H/nz
    coef = (gamma*(hy*hz)/hx, gamma*(hx*hz)/hy, gamma*(hx*hy)/hz)
    bi, bj, bk = rank % px, (rank // px) % py, rank // (px*py)

    def block_range(n, p, b):
        idx = np.arange(n)
        blk = np.minimum(idx*p // n, p - 1)
        sel = np.flatnonzero(blk == b)
        return int(sel[0]), int(sel[-1]) + 1

    i0, i1 = block_range(nx, px, bi)
    j0, j1 = block_range(ny, py, bj)
    k0, k1 = block_range(nz, pz, bk)
    lx, ly, lz = i1 - i0, j1 - j0, k1 - k0
    n = lx*ly*lz
    l, u = hex_ldu(lx, ly, lz)
    d = u.astype(np.int64) - l
    mx = d == 1
    my = (d == lx) & ~mx & (ly > 1)
    upper = -np.where(mx, coef[0], np.where(my, coef[1], coef[2]))
    c = np.arange(n, dtype=np.int64)
    li, lj, lk = c % lx, (c // lx) % ly, c // (lx*ly)
    gi, gj, gk = li + i0, lj + j0, lk + k0
    diag = (coef[0]*((gi > 0).astype(np.float64) + (gi < nx - 1)) + coef[1]*((gj > 0).astype(np.float64) + (gj < ny - 1))
            + coef[2]*((gk > 0).astype(np.float64) + (gk < nz - 1)))
    diag[gi == 0] += gamma*(hy*hz)/(0.5*hx)
    b = np.zeros(n)
    b[gi == nx - 1] += gamma*(hy*hz)*(tz/gamma)
    # processor patches, neighbours in ascending rank order: -z, -y, -x, +x, +y, +z
    cand = []
    if bk > 0: cand.append((rank - px*py, lk == 0, coef[2]))
    if bj > 0: cand.append((rank - px, lj == 0, coef[1]))
    if bi > 0: cand.append((rank - 1, li == 0, coef[0]))
    if bi < px - 1: cand.append((rank + 1, li == lx - 1, coef[0]))
    if bj < py - 1: cand.append((rank + px, lj == ly - 1, coef[1]))
    if bk < pz - 1: cand.append((rank + px*py, lk == lz - 1, coef[2]))
    cand.sort(key=lambda t: t[0])
    pfc = [np.flatnonzero(m).astype(np.int32) for _, m, _ in cand]
    pbc = [np.full(fc.shape[0], cf) for fc, (_, _, cf) in zip(pfc, cand)]
    nbr = [r for r, _, _ in cand]
    gid = (gi + nx*(gj + ny*gk))
    return dict(l=l, u=u, diag=diag, upper=upper, b=b, nCells=n, cells=gid, patchFaceCells=pfc, patchBouCoeffs=pbc,
                patchNbrDomain=nbr, gamma=gamma)


def scalar_laplacian_system(mesh, geom, gamma, fixedValue=None, refCell=None, refValue=0.0, rate=None):
    """The OTHER symmetric systems that reach the same lduMatrix::solver boundary, assembled the way
    [EXT] gaussLaplacianScheme::fvmLaplacianUncorrected + fixedValue / zeroGradient patches + fvMatrix::setReference do:
      * TEqn (src/solvers/thermalStressFoam/TEqn.H:9-15): fvm::ddt(rhoC, T) - fvm::laplacian(k, T) with fixedValue and
        zeroGradient patches -> fixedValue = {patch: value}, rate = rhoC/deltaT per cell (None: steady);
      * pEqn (flowModels/consistentIcoFlow/consistentIcoFlow.C:213-224): fvm::laplacian(1/A, p) == div(phi) with all-Neumann
        patches and a reference cell -> refCell, refValue.
    Returns dict(l, u, diag, upper, source, patches=[(faceCells, internalCoeffs, boundaryCoeffs)]) of laplacian(gamma, psi)
    (negative definite, as foam assembles it); `rate` adds +rate*V on the diagonal of MINUS that matrix, i.e. the system
    returned is  rate*V*psi - laplacian(gamma, psi) = source  when rate is given, else  laplacian(gamma, psi) = source."""
    nIF, nC = mesh.nInternalFaces, mesh.nCells
    l = np.ascontiguousarray(mesh.owner[:nIF], dtype=np.int32)
    u = np.ascontiguousarray(mesh.neighbour, dtype=np.int32)
    gam = gamma*np.ones(mesh.nFaces)
    upper = geom.deltaCoeffs[:nIF]*gam[:nIF]*geom.magSf[:nIF]
    diag = np.zeros(nC)
    np.subtract.at(diag, l, upper)              # negSumDiag: Diag[l[f]] -= Lower[f]; Diag[u[f]] -= Upper[f]
    np.subtract.at(diag, u, upper)
    source = np.zeros(nC)
    patches = []
    for p in mesh.patches:
        fs = slice(p["start"], p["start"] + p["size"])
        fc = np.ascontiguousarray(mesh.owner[fs], dtype=np.int64)
        if fixedValue and p["name"] in fixedValue:
            # fixedValue: gradientInternalCoeffs = -deltaCoeffs, gradientBoundaryCoeffs = deltaCoeffs*value;
            # [EXT gaussLaplacianScheme]: internalCoeffs = gammaMagSf*gradientInternalCoeffs,
            #                             boundaryCoeffs = -gammaMagSf*gradientBoundaryCoeffs
            gms = gam[fs]*geom.magSf[fs]
            ic = gms*(-geom.deltaCoeffs[fs])
            bc = -gms*(geom.deltaCoeffs[fs]*fixedValue[p["name"]])
        else:                                   # zeroGradient
            ic = np.zeros(p["size"]); bc = np.zeros(p["size"])
        patches.append((fc, ic, bc))
    sgn = 1.0
    if rate is not None:                        # rate*V*psi - laplacian: flip the laplacian's sign, add the ddt diagonal
        sgn = -1.0
        upper = -upper; diag = -diag + rate*geom.V
        patches = [(fc, -ic, -bc) for fc, ic, bc in patches]
    if refCell is not None:                     # fvMatrix::setReference: source[celli] += diag[celli]*value; diag[celli] += diag[celli]
        source[refCell] += diag[refCell]*refValue
        diag[refCell] += diag[refCell]
    return dict(l=l, u=u, diag=diag, upper=np.ascontiguousarray(upper), source=source, patches=patches, sign=sgn)


# ---------------------------------------------------------------------------------------------------------------------
# config 1: the shipped run/stressFoam/plateHole case and its analytical (Kirsch) solution
# ---------------------------------------------------------------------------------------------------------------------
PLATE_HOLE = dict(T=10000.0, a=0.5, E=2.0e11, nu=0.3, rho=7854.0, planeStress=True,
                  # constant/stressProperties:19-26, system/fvSolution:18-27, system/fvSchemes:49-56
                  nCorrectors=1000, convergenceTolerance=1e-8, tolerance=1e-9, relTol=0.01, limitCoeff=1.0,
                  # 0/D:19-62 (the tractions are overwritten by setPlateHoleBC::setBC at start-up)
                  patchTypes={"left": "symmetryDisplacement", "right": "tractionDisplacement", "down": "symmetryDisplacement",
                              "up": "tractionDisplacement", "hole": "tractionDisplacement", "frontAndBack": "empty"},
                  # 0/pointD:19-50
                  pointPatchTypes={"left": "symmetryPlane", "right": "calculated", "down": "symmetryPlane", "up": "calculated",
                                   "hole": "calculated", "frontAndBack": "empty"})


def kirsch_stress(C, T=10000.0, a=0.5):
    """setPlateHoleBC::plateHoleStress (run/stressFoam/plateHole/setPlateHoleBC/setPlateHoleBC.C:57-113): the stress of an
    infinite plate with a circular hole of radius a under uniaxial tension T along x.  C (n,3) -> (n,6) symmTensor
    (xx xy xz yy yz zz)."""
    C = np.atleast_2d(np.asarray(C, dtype=np.float64))
    r = np.sqrt(C[:, 0]**2 + C[:, 1]**2)
    th = np.arctan2(C[:, 1], C[:, 0])
    a2r2 = a**2/r**2
    a4r4 = 3.0*a**4/(2.0*r**4)
    s = np.zeros((C.shape[0], 6))
    s[:, 0] = T*(1.0 - a2r2*(3.0*np.cos(2*th)/2.0 + np.cos(4*th)) + a4r4*np.cos(4*th))
    s[:, 3] = T*(-a2r2*(np.cos(2*th)/2.0 - np.cos(4*th)) - a4r4*np.cos(4*th))
    s[:, 1] = T*(-a2r2*(np.sin(2*th)/2.0 + np.sin(4*th)) + a4r4*np.sin(4*th))
    return s


def kirsch_displacement(C, T=10000.0, a=0.5, E=2.0e11, nu=0.3):
    """setPlateHoleBC::plateHoleDisplacement (setPlateHoleBC.C:116-152), plane stress: kappa = (3 - nu)/(1 + nu)."""
    C = np.atleast_2d(np.asarray(C, dtype=np.float64))
    mu = E/(2.0*(1.0 + nu))
    kappa = (3.0 - nu)/(1.0 + nu)
    r = np.sqrt(C[:, 0]**2 + C[:, 1]**2)
    th = np.arctan2(C[:, 1], C[:, 0])
    d = np.zeros((C.shape[0], 3))
    d[:, 0] = (a*T/(8.0*mu))*((r/a)*(kappa + 1.0)*np.cos(th) + (2.0*a/r)*((1.0 + kappa)*np.cos(th) + np.cos(3*th))
                              - (2.0*a**3/r**3)*np.cos(3*th))
    d[:, 1] = (a*T/(8.0*mu))*((r/a)*(kappa - 3.0)*np.sin(th) + (2.0*a/r)*((1.0 - kappa)*np.sin(th) + np.sin(3*th))
                              - (2.0*a**3/r**3)*np.sin(3*th))
    return d


def plate_hole_tractions(mesh, geom, patchTypes):
    """setPlateHoleBC::setBC (setPlateHoleBC.C:202-247): on every tractionDisplacement patch traction = nf & sigma(Cf) with
    the z of Cf dropped; on `hole` the stress is evaluated at the face centre projected onto the circle r = 0.5."""
    nIF = mesh.nInternalFaces
    tr = np.zeros((mesh.nFaces - nIF, 3))
    for p in mesh.patches:
        if patchTypes[p["name"]] != "tractionDisplacement":
            continue
        fs = slice(p["start"], p["start"] + p["size"])
        nf = geom.Sf[fs]/geom.magSf[fs, None]
        C = geom.Cf[fs].copy(); C[:, 2] = 0.0
        if p["name"] == "hole":
            C = C/np.sqrt(np.sum(C*C, axis=1))[:, None]
            C = C*0.5
        s = kirsch_stress(C)
        tr[p["start"] - nIF:p["start"] - nIF + p["size"]] = np.stack(
            [nf[:, 0]*s[:, 0] + nf[:, 1]*s[:, 1] + nf[:, 2]*s[:, 2],
             nf[:, 0]*s[:, 1] + nf[:, 1]*s[:, 3] + nf[:, 2]*s[:, 4],
             nf[:, 0]*s[:, 2] + nf[:, 1]*s[:, 4] + nf[:, 2]*s[:, 5]], axis=1)
    return tr


def plate_hole_errors(mesh, geom, D, pointD, sigma):
    """setPlateHoleBC::calcError (setPlateHoleBC.C:250-388): max |D - Da| over cells, max |pointD - Da| over points, max
    |sigma_ij - sigma_a,ij| over cells for xx, xy, yy (sigma (n,6) symmTensor)."""
    C = geom.C.copy(); C[:, 2] = 0.0
    Da = kirsch_displacement(C)
    P = mesh.points.copy(); P[:, 2] = 0.0
    Pa = kirsch_displacement(P)
    sa = kirsch_stress(C)
    return dict(DError=float(np.sqrt(np.sum((D - Da)**2, axis=1)).max()),
                pointDError=float(np.sqrt(np.sum((pointD - Pa)**2, axis=1)).max()),
                sigmaXXErr=float(np.abs(sigma[:, 0] - sa[:, 0]).max()), sigmaXYErr=float(np.abs(sigma[:, 1] - sa[:, 1]).max()),
                sigmaYYErr=float(np.abs(sigma[:, 3] - sa[:, 3]).max()),
                DMax=float(np.sqrt(np.sum(Da**2, axis=1)).max()), sigmaXXMax=float(np.abs(sa[:, 0]).max()))


# ---------------------------------------------------------------------------------------------------------------------
# shipped solid cases: the settings unsTotalLagrangianStress reads from the case dictionaries
# ---------------------------------------------------------------------------------------------------------------------
def solid_case_settings(d):
    """d = dict(D, pointD, rheologyProperties, stressProperties, fvSolution, fvSchemes, controlDict) as foamdict.parse
    returns them.  -> the arguments of a solid case (GpuSolidCase / the oracle's SolidCase) and of its evolve():
    unsTotalLagrangianStress.C:59-183 (constructor: rheology, stressProperties), :769-845 (evolve: nCorrectors,
    convergenceTolerance, relConvergenceTolerance, nonLinear), fvSolution solvers.D / relaxationFactors.D, fvSchemes
    d2dt2(D) / ddt(D) / snGrad(D)."""
    from . import foamdict as fd
    rh = d["rheologyProperties"]
    coeffs = d["stressProperties"][str(d["stressProperties"]["stressModel"]) + "Coeffs"]
    sol = d["fvSolution"]["solvers"]["D"]
    if str(sol["solver"]) not in ("PCG", "gpuPCG") or str(sol.get("preconditioner", "DIC")) != "DIC":
        raise ValueError("solvers.D is not PCG/DIC: %r" % (sol,))
    sch = d["fvSchemes"]

    def scheme(group, key):
        g = sch.get(group, {})
        v = g.get(key, g.get("default", "none"))
        return v if isinstance(v, list) else [v]

    sn = scheme("snGradSchemes", "snGrad(D)")
    lap = scheme("laplacianSchemes", "laplacian(DD,D)")
    if lap[:2] != ["Gauss", "linear"] or lap[2:] != sn:
        raise ValueError("laplacian(DD,D) %r does not match snGrad(D) %r" % (lap, sn))
    if sn[0] == "skewCorrected":
        limitCoeff = float(sn[1])
    elif sn[0] == "corrected":
        # [EXT] correctedSnGrad: not restated; it and skewCorrected 1 both vanish (to rounding) on an orthogonal mesh such
        # as the beamInCrossFlow solid -- the caller checks the mesh (oracle/ASSUMPTIONS.md #11)
        limitCoeff = 1.0
    else:
        raise ValueError("unsupported snGrad(D) scheme %r" % (sn,))
    out = dict(E=fd.dimensioned(rh["rheology"]["E"]), nu=fd.dimensioned(rh["rheology"]["nu"]), rho=fd.dimensioned(rh["rheology"]["rho"]),
               planeStress=fd.switch(rh["planeStress"]), nCorrectors=int(coeffs["nCorrectors"]),
               convergenceTolerance=float(coeffs["convergenceTolerance"]),
               relConvergenceTolerance=float(coeffs.get("relConvergenceTolerance", 0.0)),
               nonLinear=fd.switch(coeffs.get("nonLinear", "no")), tolerance=float(sol.get("tolerance", 1e-6)),
               relTol=float(sol.get("relTol", 0.0)), maxIter=int(sol.get("maxIter", 1000)),
               relaxD=(float(d["fvSolution"].get("relaxationFactors", {})["D"]) if "D" in d["fvSolution"].get("relaxationFactors", {}) else None),
               d2dt2=str(scheme("d2dt2Schemes", "d2dt2(D)")[0]), ddt=str(scheme("ddtSchemes", "ddt(D)")[0]),
               snGradScheme=str(sn[0]), limitCoeff=limitCoeff, deltaT=float(d["controlDict"]["deltaT"]),
               patchTypes={}, pointPatchTypes={}, traction={}, pressure={}, fixedValue={})
    for name, b in d["D"]["boundaryField"].items():
        t = str(b["type"])
        out["patchTypes"][name] = t
        if t == "tractionDisplacement":
            out["traction"][name] = fd.uniform(b["traction"], 1, 3)[0]
            out["pressure"][name] = fd.uniform(b["pressure"], 1, 1)[0][0]
        elif t in ("fixedDisplacement", "fixedValue"):
            out["fixedValue"][name] = fd.uniform(b["value"], 1, 3)[0]
    for name, b in d["pointD"]["boundaryField"].items():
        out["pointPatchTypes"][name] = str(b["type"])
    return out


def boundary_fields(mesh, settings, traction=None, pressure=None):
    """Per-boundary-face traction (nBF,3), pressure (nBF,), fixedValue (nBF,3) from the patch-wise `uniform` entries of
    0/D; `traction` / `pressure` = {patch: value} override the dictionary (the load a fluid or a function object sets)."""
    nIF = mesh.nInternalFaces
    nBF = mesh.nFaces - nIF
    tr, pr, fv = np.zeros((nBF, 3)), np.zeros(nBF), np.zeros((nBF, 3))
    for p in mesh.patches:
        s = slice(p["start"] - nIF, p["start"] - nIF + p["size"])
        n = p["name"]
        if n in settings["traction"]:
            tr[s] = settings["traction"][n]; pr[s] = settings["pressure"][n]
        if traction and n in traction:
            tr[s] = traction[n]
        if pressure and n in pressure:
            pr[s] = pressure[n]
        if n in settings["fixedValue"]:
            fv[s] = settings["fixedValue"][n]
    return tr, pr, fv


def three_k(E, nu, rho, planeStress=False):
    """unsTotalLagrangianStress::threeK() (unsTotalLagrangianStress.C:1306-1318) = constitutiveModel::threeK()*rho
    (constitutiveModel.C:260-303): E/(rho*(1 - 2 nu)) [plane stress: E/(rho*(1 - nu))] times rho."""
    return (E/(rho*(1.0 - nu)))*rho if planeStress else (E/(rho*(1.0 - 2.0*nu)))*rho


# run/thermalStressFoam/flange: constant/thermalProperties:17-24, constant/rheologyProperties, 0/T, system/controlDict (config 4)
FLANGE = dict(E=2.0e11, nu=0.3, rho=7800.0, C=500.0, k=50.0, alpha=1.0e-6, T0=0.0, Tcold=273.0, Thot=573.0, deltaT=0.01,
              nCorrectors=1000, convergenceTolerance=1e-5, relaxD=0.5)


def flange_case(polymeshNpz):
    """The mesh of config 4 (run/thermalStressFoam/flange: flange.ans through ansys.ans_to_polymesh) with what the case records of
    its boundary.  `ansys.boundary_surfaces` recovers patch2 (348 faces, the flat face y = -7.5, T = 273) and patch4 (384 faces, the
    central bore, T = 573) by their sizes; T is zeroGradient on the rest (patch1, patch3), so the T equation can run with its
    SHIPPED boundary conditions.  For D the shipped split of the remaining faces into patch1 (traction-free) and patch3 (96
    faces, fixedDisplacement) is not recorded: the two bolt holes (2 x 144 faces) stand in for the fixed patch.
    Returns (mesh with patches cold / hot / fixed / free, geometry, T boundary kind, T boundary value, D patch types)."""
    import numpy as np
    from . import ansys, geometry, meshes
    m = meshes.PolyMesh.load_npz(polymeshNpz)
    g = geometry.compute(m)
    nIF = m.nInternalFaces
    lab, sizes = ansys.boundary_surfaces(m, g)
    sizes = np.asarray(sizes)
    cold, hot = int(np.flatnonzero(sizes == 348)[0]), int(np.flatnonzero(sizes == 384)[0])
    holes = [int(s) for s in np.flatnonzero(sizes == 144) if abs(g.Cf[nIF:][lab == s][:, 0].mean()) > 10.0]
    if len(holes) != 2:
        raise ValueError("flange: bolt holes not found")
    label = np.full(lab.shape, 3)
    label[lab == cold] = 0; label[lab == hot] = 1; label[np.isin(lab, holes)] = 2
    m2, order = m.repatch(label, ["cold", "hot", "fixed", "free"])
    g2 = geometry.compute(m2)
    kind = np.where(label[order] <= 1, 1, 0).astype(np.int32)
    val = np.where(label[order] == 1, FLANGE["Thot"], FLANGE["Tcold"]).astype(np.float64)
    pt = {"cold": "tractionDisplacement", "hot": "tractionDisplacement", "fixed": "fixedDisplacement", "free": "tractionDisplacement"}
    return m2, g2, kind, val, pt
