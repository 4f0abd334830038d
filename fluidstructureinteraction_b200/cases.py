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
