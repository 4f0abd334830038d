"""fvMesh geometry as foam-extend-3.1 computes it (INPUT generator for the assembly kernels and their oracle).

In a foam-extend run these fields come from the mesh object (mesh.Sf(), mesh.magSf(), mesh.Cf(), mesh.C(), mesh.V(),
mesh.weights(), mesh.deltaCoeffs()) and are uploaded once per mesh; here they are restated from the published
algorithms (SURVEY.md Appendix A.8 -- [EXT] primitiveMeshFaceCentresAndAreas.C, primitiveMeshCellCentresAndVols.C,
surfaceInterpolation::makeWeights / makeDeltaCoeffs).  Both the oracle and the GPU kernels take them as given, so a
detail that differed from the real 3.1 source would move config-1 numbers but not oracle <-> GPU parity.
"""
import numpy as np

VSMALL = 1.0e-300


class FvGeometry(object):
    """Sf (nF,3), magSf (nF,), Cf (nF,3), C (nC,3), V (nC,), weights (nIF,), deltaCoeffs (nF,) incl. boundary faces."""
    pass


def _face_geometry(points, faceStart, facePoints):
    nF = faceStart.shape[0] - 1
    sizes = np.diff(faceStart)
    Cf = np.zeros((nF, 3))
    Sf = np.zeros((nF, 3))
    for n in np.unique(sizes):
        sel = np.flatnonzero(sizes == n)
        idx = faceStart[sel][:, None] + np.arange(n)[None, :]
        P = points[facePoints[idx]]                                     # (m, n, 3)
        if n == 3:
            Cf[sel] = (1.0/3.0)*(P[:, 0] + P[:, 1] + P[:, 2])
            Sf[sel] = 0.5*np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0])
        else:
            fCentre = P[:, 0].copy()
            for k in range(1, n):
                fCentre += P[:, k]
            fCentre /= n
            sumN = np.zeros((sel.shape[0], 3))
            sumA = np.zeros(sel.shape[0])
            sumAc = np.zeros((sel.shape[0], 3))
            for k in range(n):
                nxt = P[:, (k + 1) % n]
                c = P[:, k] + nxt + fCentre
                nvec = np.cross(nxt - P[:, k], fCentre - P[:, k])
                a = np.sqrt(np.sum(nvec*nvec, axis=1))
                sumN += nvec
                sumA += a
                sumAc += a[:, None]*c
            Cf[sel] = (1.0/3.0)*sumAc/(sumA + VSMALL)[:, None]
            Sf[sel] = 0.5*sumN
    return Cf, Sf


def compute(mesh):
    g = FvGeometry()
    own, nei = mesh.owner.astype(np.int64), mesh.neighbour.astype(np.int64)
    nC, nF, nIF = mesh.nCells, mesh.nFaces, mesh.nInternalFaces
    Cf, Sf = _face_geometry(mesh.points, mesh.faceStart.astype(np.int64), mesh.facePoints.astype(np.int64))
    # cell centres and volumes by pyramid decomposition about the face-centre average
    cEst = np.zeros((nC, 3))
    nCellFaces = np.zeros(nC)
    np.add.at(cEst, own, Cf)
    np.add.at(nCellFaces, own, 1.0)
    np.add.at(cEst, nei, Cf[:nIF])
    np.add.at(nCellFaces, nei, 1.0)
    cEst /= nCellFaces[:, None]
    ctr = np.zeros((nC, 3))
    vol = np.zeros(nC)
    p3 = np.maximum(np.sum(Sf*(Cf - cEst[own]), axis=1), VSMALL)
    pc = 0.75*Cf + 0.25*cEst[own]
    np.add.at(ctr, own, p3[:, None]*pc)
    np.add.at(vol, own, p3)
    p3n = np.maximum(np.sum(Sf[:nIF]*(cEst[nei] - Cf[:nIF]), axis=1), VSMALL)
    pcn = 0.75*Cf[:nIF] + 0.25*cEst[nei]
    np.add.at(ctr, nei, p3n[:, None]*pcn)
    np.add.at(vol, nei, p3n)
    g.C = ctr/vol[:, None]
    g.V = vol*(1.0/3.0)
    g.Sf, g.Cf = Sf, Cf
    g.magSf = np.sqrt(np.sum(Sf*Sf, axis=1))
    # linear interpolation weights and delta coefficients
    SfdOwn = np.abs(np.sum(Sf[:nIF]*(Cf[:nIF] - g.C[own[:nIF]]), axis=1))
    SfdNei = np.abs(np.sum(Sf[:nIF]*(g.C[nei] - Cf[:nIF]), axis=1))
    g.weights = SfdNei/(SfdOwn + SfdNei)
    delta = g.C[nei] - g.C[own[:nIF]]
    magDelta = np.sqrt(np.sum(delta*delta, axis=1))
    nHat = Sf[:nIF]/g.magSf[:nIF, None]
    dc = np.empty(nF)
    dc[:nIF] = 1.0/np.maximum(np.sum(nHat*delta, axis=1), 0.05*magDelta)
    db = Cf[nIF:] - g.C[own[nIF:]]
    dc[nIF:] = 1.0/np.sqrt(np.sum(db*db, axis=1))
    g.deltaCoeffs = dc
    return g


def least_squares_vectors(mesh, g, emptyPatches=()):
    """[EXT] leastSquaresVectors::makeLeastSquaresVectors -- mesh data like weights / deltaCoeffs, an INPUT of the scalar
    `skewCorrected` correction (its gradient scheme `snGradCorr(T) leastSquares`, run/thermalStressFoam/flange/system/fvSchemes):
        dd[cell] += (1/magSqr(d)) sqr(d) over the cell's faces, d = C[nei] - C[own] (internal) or Cf - C[own] (boundary);
        lsP[f] = (1/magSqr(d)) (inv(dd[own]) & d),  lsN[f] = -(1/magSqr(d)) (inv(dd[nei]) & d),  lsB likewise for boundary faces.
    Faces of `empty` patches take no part (3-D meshes: dd is invertible).  Returns lsP (nIF,3), lsN (nIF,3), lsB (nBF,3)."""
    own, nei = mesh.owner.astype(np.int64), mesh.neighbour.astype(np.int64)
    nC, nF, nIF = mesh.nCells, mesh.nFaces, mesh.nInternalFaces
    keep = np.ones(nF - nIF, dtype=bool)
    for p in mesh.patches:
        if p["name"] in emptyPatches:
            keep[p["start"] - nIF:p["start"] - nIF + p["size"]] = False

    def sqr_over_magsqr(d):
        w = 1.0/np.sum(d*d, axis=1)
        return w, np.stack([d[:, 0]*d[:, 0], d[:, 0]*d[:, 1], d[:, 0]*d[:, 2], d[:, 1]*d[:, 1], d[:, 1]*d[:, 2], d[:, 2]*d[:, 2]], axis=1)*w[:, None]

    dI = g.C[nei] - g.C[own[:nIF]]
    wI, sI = sqr_over_magsqr(dI)
    dB = g.Cf[nIF:] - g.C[own[nIF:]]
    wB, sB = sqr_over_magsqr(dB)
    dd = np.zeros((nC, 6))
    np.add.at(dd, own[:nIF], sI)
    np.add.at(dd, nei, sI)
    np.add.at(dd, own[nIF:][keep], sB[keep])
    xx, xy, xz, yy, yz, zz = (dd[:, k] for k in range(6))
    det = xx*yy*zz + xy*yz*xz + xz*xy*yz - xx*yz*yz - xy*xy*zz - xz*yy*xz
    if np.any(np.abs(det) < 1e-30):
        raise ValueError("singular least-squares tensor (2-D mesh?): not handled")
    inv = np.stack([yy*zz - yz*yz, xz*yz - xy*zz, xy*yz - xz*yy, xx*zz - xz*xz, xy*xz - xx*yz, xx*yy - xy*xy], axis=1)/det[:, None]

    def sym_dot(S, v):
        return np.stack([S[:, 0]*v[:, 0] + S[:, 1]*v[:, 1] + S[:, 2]*v[:, 2],
                         S[:, 1]*v[:, 0] + S[:, 3]*v[:, 1] + S[:, 4]*v[:, 2],
                         S[:, 2]*v[:, 0] + S[:, 4]*v[:, 1] + S[:, 5]*v[:, 2]], axis=1)

    lsP = wI[:, None]*sym_dot(inv[own[:nIF]], dI)
    lsN = -wI[:, None]*sym_dot(inv[nei], dI)
    lsB = wB[:, None]*sym_dot(inv[own[nIF:]], dB)
    lsB[~keep] = 0.0
    return np.ascontiguousarray(lsP), np.ascontiguousarray(lsN), np.ascontiguousarray(lsB)
